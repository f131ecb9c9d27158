"""CUDA path vs the golden vectors produced by the reference itself (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from deftet_b200.grid import acute_lattice_grid
from tests.util import deformed_grid, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_energies_vs_reference_golden():
    from deftet_b200 import energies as E
    z = np.load(os.path.join(GOLD, "energies_res8.npz"))
    g, pos, tet = deformed_grid(8, 2, seed=int(z["seed"]))
    dtet = tet.cuda().to(torch.int32)
    inv = E.tet_inverse_v(torch.from_numpy(g.centred()).cuda(), dtet)
    assert rel_err(inv, z["inverse_v"]) < 1e-5
    p = pos.cuda().requires_grad_(True)
    am, ed, vv = E.tet_energies(p, dtet, inv)
    w = z["weights"]
    (float(w[0]) * am + float(w[1]) * ed + float(w[2]) * vv).sum().backward()
    assert rel_err(am, z["amips"]) < 1e-5 and rel_err(ed, z["edge"]) < 1e-5 and rel_err(vv, z["volvar"]) < 2e-5
    assert rel_err(p.grad, z["grad"]) < 1e-5


def test_builders_vs_reference_golden():
    from deftet_b200 import builders
    z = np.load(os.path.join(GOLD, "builders_res8.npz"))
    g = acute_lattice_grid(8)
    tet = torch.from_numpy(g.tets).cuda()
    f3, ft2, fs2, bnd = builders.tet_to_face(g.n_vert, tet)
    assert np.array_equal(f3.cpu().numpy(), z["f3"]) and np.array_equal(ft2.cpu().numpy(), z["ft2"])
    assert np.array_equal(fs2.cpu().numpy(), z["fs2"]) and np.array_equal(bnd.cpu().numpy(), z["bnd"])
    assert np.array_equal(builders.tet_adj_share(tet, g.n_vert).cpu().numpy(), z["share"])
    assert np.array_equal(builders.tet_face_adj(tet, g.n_vert).cpu().numpy(), z["face_adj"])
    assert np.array_equal(builders.tet_point_adj(tet, g.n_vert).cpu().numpy(), z["point_adj_sorted"])
    m, inv = builders.collapse_vertices(torch.from_numpy(g.centred()[g.tets.reshape(-1)]).cuda())
    assert np.array_equal(m.cpu().numpy(), z["colaps_map"]) and np.array_equal(inv.cpu().numpy(), z["colaps_inv"])


def test_known_answer_counts_at_scale():
    """Size-independent properties at bench scale (res 70): Euler-type counts and symmetric adjacency."""
    from deftet_b200 import builders
    g = acute_lattice_grid(70)
    tet = torch.from_numpy(g.tets).cuda()
    f3, ft2, fs2, bnd = builders.tet_to_face(g.n_vert, tet)
    assert 2 * f3.shape[0] + bnd.shape[0] == 4 * g.n_tet
    share = builders.tet_adj_share(tet, g.n_vert)
    assert share.shape[0] == 2 * f3.shape[0]
    a = share[:, 0].long() * g.n_tet + share[:, 1].long()
    b = share[:, 1].long() * g.n_tet + share[:, 0].long()
    assert torch.equal(torch.sort(a)[0], torch.sort(b)[0])                    # (t0,t1) present <=> (t1,t0) present
    edges = builders.tet_point_adj(tet, g.n_vert)
    e = edges[:, 0].long() * g.n_vert + edges[:, 1].long()
    er = edges[:, 1].long() * g.n_vert + edges[:, 0].long()
    assert torch.equal(e, torch.sort(e)[0]) and torch.equal(torch.sort(er)[0], e)   # sorted, symmetric
    m, inv = builders.collapse_vertices(torch.from_numpy(g.centred()[g.tets.reshape(-1)]).cuda())
    assert inv.shape[0] == g.n_vert and int(m.max()) == g.n_vert - 1          # tet soup collapses back to the V vertices
    assert torch.equal(m[inv.long()].cpu(), torch.arange(g.n_vert, dtype=torch.int32))
