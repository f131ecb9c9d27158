"""X1 (b): three iterations of the REFERENCE'S OWN optimisation loop `optimzie()`
(diff_render/diftet_6_subdiv/6_optim/optim_with_mask_subdiv_from_gridmov.py:105-408, unmodified, imported from the unpacked
copy) on a res-8 grid and a synthetic 6-view 32x32 data set, with deftet_b200 behind it.  Run as a subprocess with
cwd = <copy>/diff_render/diftet_6_subdiv/6_optim (the script uses relative paths) by tests/test_gpu_reference_callers.py.

argv: <mode> <out.json>
  mode = leaf   the reference's own 3_model/deftet.py::Deftet and 5_rendereq/deftetrneder.py::rendermeshcolor (unmodified) run on
                the drop-in LEAF modules (prepare_for_wz, utils_tetsv, cameraop, vertex2face) and the Kaolin shim
                (kal.render.mesh.deftet_sparse_render -> csrc/render.cu);
  mode = fused  `from deftet import Deftet` / `from deftetrneder import ...` resolve to the drop-in fused model and renderer."""
import argparse
import json
import os
import shutil
import sys
import tempfile

mode, out_path = sys.argv[1], sys.argv[2]
REPO = os.environ["DEFTET_B200_REPO"]
DROPIN_DR = os.path.join(REPO, "deftet_b200", "dropin", "diff_render")
work = tempfile.mkdtemp(prefix="x1b_")
if mode == "leaf":
    leaf = os.path.join(work, "leaf")
    os.makedirs(leaf)
    for f in ("prepare_for_wz.py", "utils_tetsv.py", "cameraop.py", "vertex2face.py"):
        shutil.copy(os.path.join(DROPIN_DR, f), leaf)
    sys.path.insert(0, leaf)
else:
    sys.path.insert(0, DROPIN_DR)
sys.path.insert(0, os.getcwd())

import numpy as np
import torch

import optim_with_mask_subdiv_from_gridmov as S          # the reference script, unmodified (its __main__ block does not run)

ref_root = os.path.realpath(os.environ["DEFTET_REFERENCE_ROOT"])
assert os.path.realpath(S.__file__).startswith(ref_root), S.__file__
model_file = os.path.realpath(sys.modules[S.Deftet.__module__].__file__)
render_file = os.path.realpath(sys.modules[S.rendermeshcolor.__module__].__file__)
if mode == "leaf":
    assert model_file.startswith(ref_root) and render_file.startswith(ref_root), (model_file, render_file)
    import prepare_for_wz
    assert os.path.realpath(prepare_for_wz.__file__).startswith(os.path.realpath(work)), prepare_for_wz.__file__
else:
    assert not model_file.startswith(ref_root) and not render_file.startswith(ref_root), (model_file, render_file)

sys.path.insert(0, REPO)
from deftet_b200.grid import acute_lattice_grid, write_tet_file
from load_blender import pose_spherical                     # reference 2_data/load_blender.py

torch.manual_seed(0)
np.random.seed(0)
res, H, W, N = 8, 32, 32, 6
g = acute_lattice_grid(res)
tetfolder = os.path.join(work, "data")
os.makedirs(tetfolder)
write_tet_file(os.path.join(tetfolder, "cube_%d_tet.tet" % res), g.vertices, g.tets)
model = S.Deftet(basefolder=tetfolder, res=res, coef=2.5, feature_dim=4, feature_raw=True, feature_fixed_dim=0, feature_fixed_init=None)

# synthetic data set: a coloured disc with alpha, N views on a sphere of radius 4 (NeRF-synthetic convention)
yy, xx = np.meshgrid(np.linspace(-1, 1, H), np.linspace(-1, 1, W), indexing="ij")
alpha = ((xx ** 2 + yy ** 2) < 0.35).astype(np.float32)
images = np.zeros((N, H, W, 4), dtype=np.float32)
for i in range(N):
    images[i, :, :, 0] = alpha * (0.3 + 0.1 * i)
    images[i, :, :, 1] = alpha * 0.5
    images[i, :, :, 2] = alpha * (0.9 - 0.1 * i)
    images[i, :, :, 3] = alpha
poses = np.stack([pose_spherical(a, -30.0, 4.0).numpy() for a in np.linspace(-180, 180, N + 1)[:-1]]).astype(np.float32)
render_poses = poses[:2]
focal = 0.5 * W / np.tan(0.5 * 0.6911)
cameras = [poses, render_poses, [H, W, focal], [np.arange(0, 4), np.arange(4, 5), np.arange(5, 6)]]

S.args = argparse.Namespace(deletenum=2, deletethres=0.001, pixelsampling=0.5, i_img=1000)
S.localtrain = False
lossweights = {"weights_im_loss": 1.0, "weights_mask_loss": 1.0, "weights_mask_reg": 0.01, "weights_point_mov": 0.1, "weights_tetvariance": 10.0,
               "weights_vector": torch.tensor([0.01, 0.01, 0.01, 0.01], device="cuda"),
               "weights_vector_with_gridmov": torch.tensor([0.01, 0.01, 0.01, 0.01, 0.1, 0.1, 0.1], device="cuda")}
before = [p.detach().clone().cpu() for p in model.parameters()]
sv = os.path.join(work, "out")
os.makedirs(sv)
n_tet_before = int(model.tet_list_tx4.shape[0]) if hasattr(model, "tet_list_tx4") else -1
S.optimzie(images, cameras, model, lr=1e-2, lr2=1e-4, bs=1, optnum=3, svfolder=sv, dev="cuda", gridmov=True, loadpath=None, sublevel=0,
           lossweights=lossweights)
after = [p.detach().clone().cpu() for p in model.parameters()]
finite = all(bool(torch.isfinite(p).all()) for p in after)
changed = any(a.shape != b.shape or float((a - b).abs().max()) > 0 for a, b in zip(after, before))
rec = {"mode": mode, "iterations": 3, "finite": finite, "params_changed": changed, "model_file": model_file, "render_file": render_file,
       "saved": sorted(os.listdir(sv))[:8]}
json.dump(rec, open(out_path, "w"))
print("X1 diffrender ok", json.dumps(rec))
shutil.rmtree(work, ignore_errors=True)
