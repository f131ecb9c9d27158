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
