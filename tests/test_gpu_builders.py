"""A10-A14 parity: GPU builders vs the reference's Python twins (restated in oracle/builders.py) and, when
oracle/_ref was built from /root/reference, vs the reference's own compiled run.so -- bit-exact."""

import numpy as np
import pytest
import torch

from deftet_b200.grid import acute_lattice_grid
from oracle import builders as orc
from oracle import native

pytestmark = pytest.mark.gpu
HAVE_REF = native.ref_lib("tet_adj_share") is not None


@pytest.fixture(scope="module", params=[8, 12])
def grid(request):
    return acute_lattice_grid(request.param)


def test_tet_to_face(grid):
    from deftet_b200 import builders
    f3, ft2, fs2, bnd = orc.tet_to_face(grid.n_vert, grid.tets)
    o3, ot2, os2, obnd = builders.tet_to_face(grid.n_vert, torch.from_numpy(grid.tets).cuda())
    assert np.array_equal(o3.cpu().numpy(), f3) and np.array_equal(ot2.cpu().numpy(), ft2)
    assert np.array_equal(os2.cpu().numpy(), fs2) and np.array_equal(obnd.cpu().numpy(), bnd)


def test_tet_point_adj(grid):
    from deftet_b200 import builders
    ref = orc.tet_to_adj_edges(grid.tets)
    edges, w = builders.tet_point_adj(torch.from_numpy(grid.tets).cuda(), grid.n_vert, normalize=True)
    assert np.array_equal(edges.cpu().numpy(), ref)
    deg = np.bincount(ref[:, 0], minlength=grid.n_vert)
    assert np.allclose(w.cpu().numpy(), 1.0 / deg[ref[:, 0]], rtol=1e-7)
    if HAVE_REF:
        out, n = native.ref_run_tet_builder("tet_point_adj", grid.tets, grid.n_vert, grid.n_tet * 12, 2)
        ref_c = out[:n]
        ref_c = ref_c[np.lexsort((ref_c[:, 1], ref_c[:, 0]))]          # reference order is hash order
        assert np.array_equal(edges.cpu().numpy(), ref_c)
    # host ABI (numpy buffers, like utils/lib/tet_point_adj/interface.py)
    out, n = builders.host_run("tet_point_adj", grid.tets, grid.n_vert, grid.n_tet * 12, 2)
    assert n == ref.shape[0] and np.array_equal(out[:n], ref)
    sp = builders.tet_to_adj_sparse(grid.n_vert, torch.from_numpy(grid.tets).cuda(), normalize=True)
    rowsum = torch.sparse.sum(sp, dim=1).to_dense()
    assert torch.allclose(rowsum[torch.from_numpy(deg > 0).cuda()], torch.ones(int((deg > 0).sum())).cuda(), atol=1e-5)


def test_tet_adj_share(grid):
    from deftet_b200 import builders
    ref = orc.tet_adj_share(grid.n_vert, grid.tets)
    out = builders.tet_adj_share(torch.from_numpy(grid.tets).cuda(), grid.n_vert)
    assert np.array_equal(out.cpu().numpy(), ref)
    if HAVE_REF:
        o, n = native.ref_run_tet_builder("tet_adj_share", grid.tets, grid.n_vert, grid.n_tet * 8, 3)
        assert np.array_equal(out.cpu().numpy(), o[:2 * n])
    o, n = builders.host_run("tet_adj_share", grid.tets, grid.n_vert, grid.n_tet * 8, 3)
    assert np.array_equal(o[:2 * n], ref)


def test_tet_face_adj(grid):
    from deftet_b200 import builders
    ref = orc.tet_face_adj(grid.n_vert, grid.tets)
    out = builders.tet_face_adj(torch.from_numpy(grid.tets).cuda(), grid.n_vert)
    assert np.array_equal(out.cpu().numpy(), ref)
    if HAVE_REF:
        o, n = native.ref_run_tet_builder("tet_face_adj", grid.tets, grid.n_vert, grid.n_tet * 200, 2)
        assert np.array_equal(out.cpu().numpy(), o[:n])
    o, n = builders.host_run("tet_face_adj", grid.tets, grid.n_vert, grid.n_tet * 200, 2)
    assert np.array_equal(o[:n], ref)


def test_tet_face_adj_int32_key_overflow():
    """n_point > 46340 makes the reference's int edge key wrap (run.cpp:39); the wrapped order is reproduced."""
    from deftet_b200 import builders
    g = acute_lattice_grid(8)
    off = 60000                                  # shift vertex ids so that a*n+b overflows int32
    tets = g.tets + off
    n_point = g.n_vert + off
    ref = orc.tet_face_adj(n_point, tets)
    out = builders.tet_face_adj(torch.from_numpy(tets).cuda(), n_point)
    assert np.array_equal(out.cpu().numpy(), ref)
    if HAVE_REF:
        o, n = native.ref_run_tet_builder("tet_face_adj", tets, n_point, g.n_tet * 200, 2)
        assert np.array_equal(out.cpu().numpy(), o[:n])


def test_collapse_vertices(grid):
    from deftet_b200 import builders
    soup = grid.centred()[grid.tets.reshape(-1)]                       # 4T x 3 tet soup
    rng = np.random.default_rng(0)
    extra = np.array([[0.0, -0.0, 1e-7], [-0.0, 0.0, -1e-7], [0.123455, 0.123465, 0.5], [0.1234549, 0.1234651, 0.5],
                      [2.5e-6, -2.5e-6, 7.5e-6], [1.5e-5, 0.000005, -0.000005], [12345678.0, -3.0e9, 8388608.0]], dtype=np.float32)
    pts = np.concatenate([soup, extra, (rng.random((500, 3)) - 0.5).astype(np.float32), extra]).astype(np.float32)
    m_ref, inv_ref = orc.colaps_v(pts)
    m, inv = builders.collapse_vertices(torch.from_numpy(pts).cuda())
    assert np.array_equal(m.cpu().numpy(), m_ref) and np.array_equal(inv.cpu().numpy(), inv_ref)
    if HAVE_REF:
        m_c, inv_c = native.ref_colaps_v(pts)
        assert np.array_equal(m.cpu().numpy(), m_c) and np.array_equal(inv.cpu().numpy(), inv_c)
    mh, invh = builders.host_colaps_v(pts)
    assert np.array_equal(mh, m_ref) and np.array_equal(invh, inv_ref)
