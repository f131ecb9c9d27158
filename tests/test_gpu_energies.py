"""A6-A8 parity: fused energy kernels vs the torch/autograd oracle (<= 1e-5 relative, the north-star bar)."""
import pytest
import torch

from oracle import energies as orc
from tests.util import deformed_grid, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.mark.parametrize("res,B", [(8, 1), (8, 3), (16, 2), (20, 2)])
def test_energies_forward_backward_indexed(res, B):
    from deftet_b200 import energies as E
    g, pos, tet = deformed_grid(res, B, seed=res + B)
    inv_ref = orc.tet_inverse_v(torch.from_numpy(g.centred()), tet)
    w = (0.7, 1.3, 1e9)      # volume variance is ~1e-12: weight it up so its gradient is visible
    ref = orc.energies_with_grad(pos, tet, inv_ref, w)

    dpos = pos.cuda().requires_grad_(True)
    dtet = tet.cuda().to(torch.int32)
    inv = E.tet_inverse_v(torch.from_numpy(g.centred()).cuda(), dtet)
    assert rel_err(inv, inv_ref) < RTOL
    am, ed, vv = E.tet_energies(dpos, dtet, inv)
    assert rel_err(am, ref["amips"]) < RTOL
    assert rel_err(ed, ref["edge"]) < RTOL
    assert rel_err(vv, ref["volvar"]) < 2e-5          # 4th centred moment of ~1e-4-sized numbers in fp32
    (w[0] * am + w[1] * ed + w[2] * vv).sum().backward()
    assert rel_err(dpos.grad, ref["grad"]) < RTOL


def test_rest_pose_known_answers():
    """AMIPS of the undeformed grid is exactly 3."""
    from deftet_b200 import energies as E
    g, pos, tet = deformed_grid(16, 2, amp=0.0)
    dtet = tet.cuda().to(torch.int32)
    inv = E.tet_inverse_v(torch.from_numpy(g.centred()).cuda(), dtet)
    am, ed, vv = E.tet_energies(pos.cuda(), dtet, inv)
    assert torch.allclose(am.cpu(), torch.full((2,), 3.0), atol=2e-6)
    ref_ed = orc.edge_length(orc.gather_tets(pos, tet))
    assert rel_err(ed, ref_ed) < RTOL


@pytest.mark.parametrize("which", ["amips", "edge", "volume"])
def test_energies_soup_dropin(which):
    """The drop-in forms take tet_bxfx4x3 like DefTet.amips_energy / edge_length / volume_variance."""
    from deftet_b200 import energies as E
    g, pos, tet = deformed_grid(12, 2, seed=5)
    inv_ref = orc.tet_inverse_v(torch.from_numpy(g.centred()), tet)
    soup = orc.gather_tets(pos, tet).clone().requires_grad_(True)
    if which == "amips":
        ref = orc.amips_energy(soup, inv_ref)
    elif which == "edge":
        ref = orc.edge_length(soup)
    else:
        ref = orc.volume_variance(soup) * 1e9
    gw = torch.tensor([0.3, 1.7])
    (ref * gw).sum().backward()

    dsoup = soup.detach().cuda().requires_grad_(True)
    if which == "amips":
        out = E.amips_energy_soup(dsoup, inv_ref.cuda())
    elif which == "edge":
        out = E.edge_length_soup(dsoup)
    else:
        out = E.volume_variance_soup(dsoup) * 1e9
    (out * gw.cuda()).sum().backward()
    tol = 2e-5 if which == "volume" else RTOL
    assert rel_err(out, ref.detach()) < tol
    assert rel_err(dsoup.grad, soup.grad) < tol


def _both_paths(monkeypatch, fn):
    """Run fn() through the opt-in tile-local kernels and through the direct-gather kernels (default)."""
    monkeypatch.setenv("DTB_ENERGY_PATH", "tiled")
    tiled = fn()
    monkeypatch.delenv("DTB_ENERGY_PATH", raising=False)
    direct = fn()
    return tiled, direct


@pytest.mark.parametrize("res,B,group", [(12, 5, "1"), (12, 5, "2"), (12, 5, "4"), (20, 3, "8"), (16, 9, None)])
def test_tiled_energies_match_oracle_and_direct_kernels(res, B, group, monkeypatch):
    """energies_tiled.cu: ragged last tile, batch not a multiple of the samples-per-CTA group, every group size."""
    from deftet_b200 import energies as E
    if group is not None:
        monkeypatch.setenv("DTB_ENERGY_GROUP", group)
    g, pos, tet = deformed_grid(res, B, seed=7 * res + B)
    inv_ref = orc.tet_inverse_v(torch.from_numpy(g.centred()), tet)
    w = (0.7, 1.3, 1e9)
    ref = orc.energies_with_grad(pos, tet, inv_ref, w)
    dtet = tet.cuda().to(torch.int32)
    inv = E.tet_inverse_v(torch.from_numpy(g.centred()).cuda(), dtet)

    def run():
        p = pos.cuda().requires_grad_(True)
        am, ed, vv = E.tet_energies(p, dtet, inv)
        (w[0] * am + w[1] * ed + w[2] * vv).sum().backward()
        return am.detach(), ed.detach(), vv.detach(), p.grad

    tiled, direct = _both_paths(monkeypatch, run)
    for out in (tiled, direct):
        assert rel_err(out[0], ref["amips"]) < RTOL and rel_err(out[1], ref["edge"]) < RTOL
        assert rel_err(out[2], ref["volvar"]) < 2e-5
        assert rel_err(out[3], ref["grad"]) < RTOL
    assert rel_err(tiled[3], direct[3]) < RTOL


def test_tiled_energies_shuffled_topology_and_flag_subsets(monkeypatch):
    monkeypatch.setenv("DTB_ENERGY_PATH", "tiled")
    """A tet order without any locality (every tile touches ~4x256 distinct vertices -> the staging buffers grow to their
    worst-case size) and each energy on its own (NULL gradients for the others)."""
    from deftet_b200 import energies as E
    g, pos, tet = deformed_grid(16, 2, seed=3)
    perm = torch.randperm(tet.shape[0], generator=torch.Generator().manual_seed(1))
    tet = tet[perm].contiguous()
    inv_ref = orc.tet_inverse_v(torch.from_numpy(g.centred()), tet)
    dtet = tet.cuda().to(torch.int32)
    inv = E.tet_inverse_v(torch.from_numpy(g.centred()).cuda(), dtet)
    tiles = E.TetTiles(dtet, g.n_vert)
    assert 256 < tiles.nloc_max <= 1024
    soup = orc.gather_tets(pos, tet)
    for flag, fn, k in ((E.AMIPS, lambda s: orc.amips_energy(s, inv_ref), 0), (E.EDGE, orc.edge_length, 1),
                        (E.VOLUME, lambda s: orc.volume_variance(s) * 1e9, 2)):
        rp = pos.clone().requires_grad_(True)
        ref = fn(orc.gather_tets(rp, tet))
        ref.sum().backward()
        p = pos.cuda().requires_grad_(True)
        out = E.tet_energies(p, dtet, inv if flag == E.AMIPS else None, flags=flag, tiles=tiles)[k]
        if flag == E.VOLUME:
            out = out * 1e9
        out.sum().backward()
        tol = 2e-5 if flag == E.VOLUME else RTOL
        assert rel_err(out, ref.detach()) < tol
        assert rel_err(p.grad, rp.grad) < tol
    # an inverse-matrix tensor that is not 16-byte aligned takes the plain-load staging path
    inv_off = torch.empty(inv.numel() + 1, device="cuda")[1:].view_as(inv).copy_(inv)
    assert inv_off.data_ptr() % 16 != 0
    a1 = E.tet_energies(pos.cuda(), dtet, inv, tiles=tiles)[0]
    a2 = E.tet_energies(pos.cuda(), dtet, inv_off, tiles=tiles)[0]
    assert torch.allclose(a1, a2, rtol=1e-6)
