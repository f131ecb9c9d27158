"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run here, where /root/reference is mounted).

  python tests/golden/make_golden.py

* energies_res8.npz   -- reference `DefTet` class (layers/DefTet/deftet.py) imported under sys.modules stubs for
                         kaolin / the CUDA-extension wrappers, autograd on CPU: energies + gradient on a seeded
                         deformation of this repo's res-8 grid.
* builders_res8.npz   -- reference `utils/tet_utils.py::tet_to_face` (pure Python) and the reference's compiled
                         ctypes builders (oracle/_ref/*/run.so through oracle/ref_driver) on the same grid.
* cube40_known.npz    -- the known answers of BASELINE.md section 3 re-derived from the shipped cube_40 grid.
The fixtures travel to the GPU box (the reference does not); tests/test_golden.py pins the oracle and the CUDA
kernels to them.
"""
import os
import subprocess
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference_deftet():
    for name in ["kaolin", "utils.tet_utils", "utils.mesh_utils", "layers.DefTet.check_condition_tetrahedron_base.utils"]:
        m = types.ModuleType(name)
        m.check_condition_f_base = None
        sys.modules[name] = m
    sys.path.insert(0, REF)
    import importlib
    return importlib.import_module("layers.DefTet.deftet")


def main():
    from deftet_b200.grid import read_tet_file
    from oracle import native
    from tests.util import deformed_grid
    ref_deftet = import_reference_deftet()
    D = ref_deftet.DefTet()
    # ---- energies on the res-8 grid -------------------------------------------------------------------------
    g, pos, tet = deformed_grid(8, 2, seed=42)
    inv = D.tet_inverse_v(torch.from_numpy(g.centred()), tet)
    p = pos.clone().requires_grad_(True)
    soup = torch.gather(p.unsqueeze(2).expand(-1, -1, 4, -1), 1, tet.unsqueeze(0).expand(2, -1, -1).unsqueeze(-1).expand(-1, -1, -1, 3))
    am, ed, vv = D.amips_energy(soup, inv), D.edge_length(soup, pow=4), D.volume_variance(soup, pow=4)
    (0.7 * am + 1.3 * ed + 1e9 * vv).sum().backward()
    np.savez(os.path.join(OUT, "energies_res8.npz"), pos=pos.numpy(), inverse_v=inv.numpy(), amips=am.detach().numpy(), edge=ed.detach().numpy(),
             volvar=vv.detach().numpy(), grad=p.grad.numpy(), weights=np.array([0.7, 1.3, 1e9]), seed=42)
    print("energies", am.tolist(), ed.tolist(), vv.tolist())
    # ---- builders on the res-8 grid ---------------------------------------------------------------------------
    with tempfile.TemporaryDirectory() as d:
        for n in ("tet_point_adj", "tet_adj_share", "tet_face_adj", "colaps_v"):
            os.makedirs(os.path.join(d, "utils/lib", n))
            os.symlink(os.path.join(ROOT, "oracle/_ref", n, "run.so"), os.path.join(d, "utils/lib", n, "run.so"))
        code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.modules['tqdm']=type(sys)('tqdm'); sys.modules['tqdm'].tqdm=lambda x:x;"
                "from utils import tet_utils; t=np.load(%r); r=tet_utils.tet_to_face(int(t['n']), t['tets']); np.savez(%r, f3=r[0], ft2=r[1], fs2=r[2], bnd=r[3])")
        np.savez(os.path.join(d, "in.npz"), tets=g.tets, n=g.n_vert)
        subprocess.run([sys.executable, "-c", code % (REF, os.path.join(d, "in.npz"), os.path.join(d, "out.npz"))], cwd=d, check=True,
                       stdout=subprocess.DEVNULL)
        t2f = dict(np.load(os.path.join(d, "out.npz")))
    share, n_share = native.ref_run_tet_builder("tet_adj_share", g.tets, g.n_vert, g.n_tet * 8, 3)
    fadj, n_fadj = native.ref_run_tet_builder("tet_face_adj", g.tets, g.n_vert, g.n_tet * 200, 2)
    padj, n_padj = native.ref_run_tet_builder("tet_point_adj", g.tets, g.n_vert, g.n_tet * 12, 2)
    padj = padj[:n_padj]
    padj = padj[np.lexsort((padj[:, 1], padj[:, 0]))]
    soup_pts = g.centred()[g.tets.reshape(-1)]
    cmap, cinv = native.ref_colaps_v(soup_pts)
    np.savez_compressed(os.path.join(OUT, "builders_res8.npz"), f3=t2f["f3"], ft2=t2f["ft2"], fs2=t2f["fs2"], bnd=t2f["bnd"],
                        share=share[:2 * n_share], face_adj=fadj[:n_fadj], point_adj_sorted=padj, colaps_map=cmap, colaps_inv=cinv)
    print("builders: faces", t2f["f3"].shape, "share", n_share, "face_adj", n_fadj, "edges", n_padj, "collapse", soup_pts.shape[0], "->", len(cinv))
    # ---- known answers on the shipped cube_40 -----------------------------------------------------------------------
    v40, t40 = read_tet_file(os.path.join(REF, "diff_render/diftet_6_subdiv/data/cube_40_tet.tet"))
    pos40 = torch.from_numpy((v40 - 0.5).astype(np.float32))
    tet40 = torch.from_numpy(t40)
    inv40 = D.tet_inverse_v(pos40, tet40)
    soup40 = pos40[tet40.reshape(-1)].reshape(1, -1, 4, 3)
    known = dict(amips=float(D.amips_energy(soup40, inv40)), edge=float(D.edge_length(soup40, pow=4)), volvar=float(D.volume_variance(soup40, pow=4)),
                 n_vert=v40.shape[0], n_tet=t40.shape[0])
    _, n_share40 = native.ref_run_tet_builder("tet_adj_share", t40, v40.shape[0], t40.shape[0] * 8, 3)
    _, n_edge40 = native.ref_run_tet_builder("tet_point_adj", t40, v40.shape[0], t40.shape[0] * 12, 2)
    known.update(shared_faces=n_share40, directed_edges=n_edge40)
    np.savez(os.path.join(OUT, "cube40_known.npz"), **known)
    print("cube_40 known answers", known)


if __name__ == "__main__":
    main()
