"""Generate tests/golden/devox.npz from the REFERENCE'S OWN `trilinear_devoxelize` (run here, where /root/reference is mounted):
    python tests/golden/make_golden_devox.py

Loads, unmodified, /root/reference/layers/pv_module/functional/devoxelization.py (its `backend` import, which JIT-builds PVCNN's CUDA
extension, is stubbed: the live `trilinear_devoxelize` at :47-53 only needs torch) and runs it on CPU tensors, forward and
autograd, for the three encoder levels of layers/pc_model.py:50 at reduced channel counts plus edge cases (coordinates on and
outside the border, exact voxel centres, a permuted coords view, the `sample_f` prelude of pc_model.py:182-194).
The fixture pins oracle/devox.py (tests/test_golden.py) and the CUDA kernels of csrc/devox.cu (tests/test_gpu_devox.py); it
travels to the GPU box, the reference does not.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_FILE = "/root/reference/layers/pv_module/functional/devoxelization.py"


def import_reference():
    for name in ("layers", "layers.pv_module", "layers.pv_module.functional", "layers.pv_module.functional.backend"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["layers.pv_module.functional.backend"]._backend = None
    spec = importlib.util.spec_from_file_location("ref_devoxelization", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def sample_f_reference(ref, point_pos, c_list):
    """layers/pc_model.py:182-194 (point-cloud branch) verbatim in behaviour: the method needs a whole DefTetModel, so its six lines are
    re-stated here around the reference's own trilinear_devoxelize."""
    point_pos = point_pos + 0.5
    point_pos = point_pos.permute(0, 2, 1)
    outs = []
    for c in c_list:
        r = c.shape[-1]
        norm_coords = torch.clamp(point_pos * r, 0, r - 1)
        outs.append(ref.trilinear_devoxelize(c, norm_coords, r, True))
    return torch.cat(outs, dim=1)


def main():
    ref = import_reference()
    g = torch.Generator().manual_seed(20260517)
    out = {}
    cases = [("r32", 1, 2, 32, 600), ("r16", 2, 8, 16, 500), ("r8", 2, 12, 8, 400), ("r5_odd", 1, 3, 5, 200), ("r40_big", 1, 1, 40, 300)]
    for name, B, C, R, N in cases:
        feat = torch.randint(-128, 128, (B, C, R, R, R), generator=g).float() / 32      # few mantissa bits: the fixture compresses
        coords = torch.rand(B, 3, N, generator=g) * (R + 1.0) - 1.0           # some outside [0, R-1] on both sides
        coords[:, :, :16] = torch.randint(0, R, (B, 3, 16), generator=g).float()      # exact voxel centres
        coords[:, :, 16:24] = float(R - 1)
        coords[:, :, 24:32] = 0.0
        coords[:, 0, 32:40] = R - 1.0 - 1e-4
        feat.requires_grad_(True)
        coords.requires_grad_(True)
        o = ref.trilinear_devoxelize(feat, coords, R, True)
        go = torch.randn(o.shape, generator=g)
        gf, gc = torch.autograd.grad(o, (feat, coords), go)
        for k, v in (("feat", feat), ("coords", coords), ("out", o), ("grad_out", go), ("grad_feat", gf), ("grad_coords", gc)):
            out["%s_%s" % (name, k)] = v.detach().numpy().astype(np.float32)

    # sample_f over three levels, positions in and slightly outside the unit cube
    B, N = 2, 700
    levels = [(1, 32), (8, 16), (12, 8)]
    c_list = [(torch.randint(-128, 128, (B, C, R, R, R), generator=g).float() / 32).requires_grad_(True) for C, R in levels]
    pos = ((torch.rand(B, N, 3, generator=g) - 0.5) * 1.06).requires_grad_(True)
    o = sample_f_reference(ref, pos, c_list)
    go = torch.randn(o.shape, generator=g)
    grads = torch.autograd.grad(o, [pos] + c_list, go)
    out["sf_pos"] = pos.detach().numpy()
    out["sf_out"] = o.detach().numpy()
    out["sf_grad_out"] = go.numpy()
    out["sf_grad_pos"] = grads[0].numpy()
    for i, c in enumerate(c_list):
        out["sf_feat%d" % i] = c.detach().numpy()
        out["sf_grad_feat%d" % i] = grads[1 + i].numpy()
    path = os.path.join(ROOT, "tests", "golden", "devox.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
