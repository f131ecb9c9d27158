"""A1 / A2 parity: binned search kernels vs the brute-force C oracle (bit-exact indices)."""
import numpy as np
import pytest
import torch

from oracle import bary as orc_bary
from oracle import energies as orc_e
from oracle import native as orc
from tests.util import deformed_grid, rel_err

pytestmark = pytest.mark.gpu


def _query_points(B, P, seed):
    gen = torch.Generator().manual_seed(seed)
    return (torch.rand(B, P, 3, generator=gen) - 0.5) * 1.05     # dataloader.py:108 style


@pytest.mark.parametrize("res,B,P,amp", [(8, 2, 4000, 0.25), (12, 1, 6000, 0.25), (8, 1, 3000, 0.0), (16, 2, 5000, 0.4)])
def test_point_in_tet_soup_bit_exact(res, B, P, amp):
    from deftet_b200 import search
    g, pos, tet = deformed_grid(res, B, seed=res, amp=amp)
    soup = orc_e.gather_tets(pos, tet)
    pts = _query_points(B, P, 7)
    if amp == 0.0:      # undeformed grid: put queries exactly on lattice vertices / face planes too
        pts[0, :100] = pos[0, :100]
        pts[0, 200:400] = torch.round(pts[0, 200:400] * res) / res
    ref = orc.point_in_tet(soup.numpy(), pts.numpy())
    out = search.point_in_tet_soup(soup.cuda(), pts.cuda())
    assert out.shape == (B, P, 1) and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), ref)
    for G in (3, 17, 40):   # result must not depend on the binning resolution
        out2 = search.point_in_tet_soup(soup.cuda(), pts.cuda(), grid_res=G)
        assert np.array_equal(out2.cpu().numpy(), ref)


def test_point_in_tet_indexed_and_bary_backward():
    from deftet_b200 import search
    res, B, P = 10, 2, 5000
    g, pos, tet = deformed_grid(res, B, seed=3)
    pts = _query_points(B, P, 11)
    soup = orc_e.gather_tets(pos, tet)
    ref = orc.point_in_tet(soup.numpy(), pts.numpy())
    dpos = pos.cuda().requires_grad_(True)
    dpts = pts.cuda().requires_grad_(True)
    cond, bary = search.point_in_tet(dpos, tet.cuda(), dpts)
    assert np.array_equal(cond.cpu().numpy(), ref)
    gen = torch.Generator().manual_seed(1)
    g_w = torch.randn(B, P, 4, generator=gen)
    (bary * g_w.cuda()).sum().backward()
    cond_t = torch.from_numpy(ref[..., 0])
    w64, gp64, gq64 = orc_bary.weights_with_grad(pos, tet, pts, cond_t, g_w, dtype=torch.float64)
    inside = cond_t >= 0
    # weights of found points are a partition of unity in [0,1] up to rounding
    assert rel_err(bary.detach().cpu()[inside], w64[inside], floor=1.0) < 1e-5
    assert float(bary.detach().cpu()[~inside].abs().max()) == 0.0
    assert rel_err(dpos.grad, gp64) < 1e-5
    assert rel_err(dpts.grad, gq64) < 1e-5
    # and the fp32 autograd of the reference expression agrees with the fp64 one to the same tolerance
    w32, gp32, _ = orc_bary.weights_with_grad(pos, tet, pts, cond_t, g_w, dtype=torch.float32)
    assert rel_err(gp32, gp64) < 1e-4


def test_point_in_tet_empty_and_outside():
    from deftet_b200 import search
    g, pos, tet = deformed_grid(8, 1, seed=0)
    soup = orc_e.gather_tets(pos, tet).cuda()
    far = torch.full((1, 64, 3), 7.0).cuda()
    assert float(search.point_in_tet_soup(soup, far).max()) == -1.0
    assert search.point_in_tet_soup(soup, torch.zeros(1, 0, 3).cuda()).shape == (1, 0, 1)
    with pytest.raises(RuntimeError):
        search.point_in_tet_soup(soup[:, :, :3], far)


@pytest.mark.parametrize("B,Q,M,seed", [(1, 3000, 5000, 0), (3, 2000, 1500, 1), (2, 500, 7, 2), (1, 10, 1, 3)])
def test_nearest_neighbor_bit_exact(B, Q, M, seed):
    from deftet_b200 import search
    gen = torch.Generator().manual_seed(seed)
    pts = torch.rand(B, M, 3, generator=gen) - 0.5
    # surface-like targets + exact duplicates (lowest index must win) + far-away queries
    pts = pts / pts.norm(dim=-1, keepdim=True).clamp(min=1e-3) * 0.35
    if M > 10:
        pts[:, M // 2:M // 2 + 5] = pts[:, 0:5]
    q = (torch.rand(B, Q, 3, generator=gen) - 0.5) * 1.2
    q[:, :5] = pts[:, :5]
    q[:, 5:8] = 9.0
    ref = orc.nearest_neighbor(q.numpy(), pts.numpy())
    out = search.nearest_neighbor_index(q.cuda(), pts.cuda())
    assert out.dtype == torch.int32
    assert np.array_equal(out.cpu().numpy().astype(np.int64), ref)
    for G in (4, 33, 70, 128):
        out = search.nearest_neighbor_index(q.cuda(), pts.cuda(), grid_res=G)
        assert np.array_equal(out.cpu().numpy().astype(np.int64), ref)
    nn = search.NearestNeighbor()
    assert nn(q.cuda(), pts.cuda()).dtype == torch.int64


def test_nearest_neighbor_quantised_ties():
    """Many exactly-equal distances (lattice points): strict '<' in index order means lowest index wins."""
    from deftet_b200 import search
    gen = torch.Generator().manual_seed(5)
    pts = torch.randint(0, 6, (2, 4000, 3), generator=gen).float() / 8
    q = torch.randint(0, 12, (2, 3000, 3), generator=gen).float() / 16
    ref = orc.nearest_neighbor(q.numpy(), pts.numpy())
    out = search.nearest_neighbor_index(q.cuda(), pts.cuda())
    assert np.array_equal(out.cpu().numpy().astype(np.int64), ref)


def test_tet_interpolate_matches_torch_gather():
    from deftet_b200 import search
    res, B, P, C = 8, 2, 3000, 3
    g, pos, tet = deformed_grid(res, B, seed=2)
    pts = _query_points(B, P, 4)
    gen = torch.Generator().manual_seed(0)
    field = torch.randn(B, g.n_vert, C, generator=gen)
    dpos = pos.cuda()
    dfield = field.cuda().requires_grad_(True)
    cond, bary = search.point_in_tet(dpos, tet.cuda(), pts.cuda())
    bary = bary.detach().requires_grad_(True)
    out = search.tet_interpolate(dfield, tet.cuda(), cond, bary)
    gw = torch.randn(B, P, C, generator=gen)
    (out * gw.cuda()).sum().backward()
    # torch reference
    rfield = field.clone().requires_grad_(True)
    rb = bary.detach().cpu().clone().requires_grad_(True)
    c = cond.cpu().squeeze(-1)
    vid = tet.long()[c.clamp(min=0).long()]
    phi = torch.gather(rfield.unsqueeze(2).expand(-1, -1, 4, -1), 1, vid.unsqueeze(-1).expand(-1, -1, -1, C))
    ref = (rb.unsqueeze(-1) * phi).sum(dim=2) * (c >= 0).float().unsqueeze(-1)
    (ref * gw).sum().backward()
    assert rel_err(out, ref.detach()) < 1e-6
    assert rel_err(dfield.grad, rfield.grad) < 1e-5
    assert rel_err(bary.grad, rb.grad * (c >= 0).float().unsqueeze(-1)) < 1e-6
    # paste_occ drop-in: misses are clamped to tet 0 in place
    occ = torch.rand(B, g.n_tet, generator=gen).cuda()
    cc = cond.clone()
    pasted = search.paste_occ(occ, cc)
    assert float(cc.min()) >= 0 and pasted.shape == (B, P)


def test_located_mse():
    from deftet_b200 import search
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(3, 1000, generator=gen)
    t = torch.randn(3, 1000, generator=gen)
    cond = torch.randint(-1, 5, (3, 1000, 1), generator=gen).float()
    cond[2] = -1.0
    dx = x.cuda().requires_grad_(True)
    loss = search.located_mse(dx, t.cuda(), cond.cuda())
    (loss * torch.tensor([1.0, 2.0, 3.0]).cuda()).sum().backward()
    rx = x.double().requires_grad_(True)
    m = (cond.squeeze(-1) >= 0).double()
    ref = (((rx - t.double()) ** 2) * m).sum(-1) / m.sum(-1).clamp(min=1.0)
    (ref * torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64)).sum().backward()
    assert rel_err(loss, ref.detach()) < 1e-6 and rel_err(dx.grad, rx.grad) < 1e-6


@pytest.mark.parametrize("kernel", ["brick", "thread"])
def test_nearest_neighbor_both_kernels_surface_workload(kernel, monkeypatch):
    """The warp-cooperative brick kernel (default) and the per-thread walk of round 1 (DTB_NN_KERNEL=thread) on the shape of the
    chamfer workload: queries within a fraction of a cell of a densely sampled surface, plus queries far from it (per-lane
    fallback), duplicates and an empty-home-cell band."""
    from deftet_b200 import search
    if kernel == "thread":
        monkeypatch.setenv("DTB_NN_KERNEL", "thread")
    else:
        monkeypatch.delenv("DTB_NN_KERNEL", raising=False)
    gen = torch.Generator().manual_seed(11)
    B, M, Q = 2, 20000, 12000
    d = torch.randn(B, M, 3, generator=gen)
    pts = d / d.norm(dim=-1, keepdim=True) * torch.tensor([0.3, 0.22]).reshape(B, 1, 1)
    pts[:, 100:110] = pts[:, 0:10]
    dq = torch.randn(B, Q, 3, generator=gen)
    q = dq / dq.norm(dim=-1, keepdim=True) * torch.tensor([0.3, 0.22]).reshape(B, 1, 1)
    q[:, : Q // 2] += 0.004 * torch.randn(B, Q // 2, 3, generator=gen)            # near the surface
    q[:, Q // 2: 3 * Q // 4] *= 1.15                                               # a shell a few cells away
    q[:, 3 * Q // 4:] = (torch.rand(B, Q - 3 * Q // 4, 3, generator=gen) - 0.5) * 2   # anywhere, also outside the bbox
    q[:, :10] = pts[:, :10]
    ref = orc.nearest_neighbor(q.numpy(), pts.numpy())
    for G in (0, 16, 48, 96):
        out = search.nearest_neighbor_index(q.cuda(), pts.cuda(), grid_res=G)
        assert np.array_equal(out.cpu().numpy().astype(np.int64), ref), "kernel %s G=%d" % (kernel, G)
