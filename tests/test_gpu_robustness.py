"""Exactness of the pruned search kernels must not depend on where the data sits or how large it is: translated / scaled
copies of the same scene (fp32 coordinates far from the origin, tiny and huge extents) against the brute-force oracle."""
import numpy as np
import pytest
import torch

from oracle import energies as orc_e
from oracle import native as orc
from tests.util import deformed_grid

pytestmark = pytest.mark.gpu

TRANSFORMS = [(1.0, 0.0), (1.0, 37.5), (100.0, -250.0), (0.01, 0.003), (1.0, 4096.0)]


@pytest.mark.parametrize("scale,shift", TRANSFORMS)
def test_nearest_neighbor_translated_scaled(scale, shift):
    from deftet_b200 import search
    gen = torch.Generator().manual_seed(0)
    d = torch.randn(2, 6000, 3, generator=gen)
    pts = (d / d.norm(dim=-1, keepdim=True) * 0.3) * scale + shift
    q = ((torch.rand(2, 4000, 3, generator=gen) - 0.5) * 0.9) * scale + shift
    q[:, :50] = pts[:, :50]
    ref = orc.nearest_neighbor(q.numpy(), pts.numpy())
    out = search.nearest_neighbor_index(q.cuda(), pts.cuda())
    assert np.array_equal(out.cpu().numpy().astype(np.int64), ref)


@pytest.mark.parametrize("scale,shift", TRANSFORMS)
def test_point_in_tet_translated_scaled(scale, shift):
    from deftet_b200 import search
    g, pos, tet = deformed_grid(8, 2, seed=4)
    pos = pos * scale + shift
    gen = torch.Generator().manual_seed(1)
    pts = ((torch.rand(2, 5000, 3, generator=gen) - 0.5) * 1.05) * scale + shift
    soup = orc_e.gather_tets(pos, tet)
    ref = orc.point_in_tet(soup.numpy(), pts.numpy())
    cond, _ = search.point_in_tet(pos.cuda(), tet.cuda(), pts.cuda())
    assert np.array_equal(cond.cpu().numpy(), ref)


@pytest.mark.parametrize("scale,shift", TRANSFORMS)
def test_point_face_distance_translated_scaled(scale, shift):
    from deftet_b200 import surface
    gen = torch.Generator().manual_seed(2)
    F, S = 600, 3000
    c = torch.randn(1, F, 1, 3, generator=gen)
    c = c / c.norm(dim=-1, keepdim=True) * 0.3
    faces = (c + 0.03 * (torch.rand(1, F, 3, 3, generator=gen) - 0.5)) * scale + shift
    pts = ((torch.rand(1, S, 3, generator=gen) - 0.5) * 0.9) * scale + shift
    d_ref, f_ref = orc.point_face_distance(pts.numpy(), faces.numpy())
    d, f = surface.tet_analytic_distance_f_batch(pts.cuda(), faces.cuda(), torch.tensor([float(F)]).cuda())
    assert np.array_equal(f.cpu().numpy(), f_ref)
    assert np.array_equal(d.cpu().numpy(), d_ref)


def test_degenerate_inputs_do_not_hang_or_crash():
    from deftet_b200 import search, surface
    # all targets identical, all queries identical, NaN / inf coordinates
    pts = torch.zeros(1, 500, 3)
    q = torch.zeros(1, 300, 3)
    assert int(search.nearest_neighbor_index(q.cuda(), pts.cuda()).max()) == 0
    q[0, 0] = float("nan")
    q[0, 1] = float("inf")
    out = search.nearest_neighbor_index(q.cuda(), pts.cuda())
    ref = orc.nearest_neighbor(q.numpy(), pts.numpy())
    assert np.array_equal(out.cpu().numpy().astype(np.int64)[0, 2:], ref[0, 2:])
    faces = torch.zeros(1, 40, 3, 3)                                  # zero-area faces: invisible to the reference (k3 == 0)
    d, f = surface.tet_analytic_distance_f_batch(torch.rand(1, 100, 3).cuda(), faces.cuda(), torch.tensor([40.0]).cuda())
    assert float(f.max()) == -1.0 and float(d.min()) == 10000.0
    g, pos, tet = deformed_grid(8, 1, seed=0)
    flat = pos.clone()
    flat[..., 2] = 0.0                                                # every tet degenerate (zero volume)
    pts = torch.rand(1, 500, 3) - 0.5
    ref = orc.point_in_tet(orc_e.gather_tets(flat, tet).numpy(), pts.numpy())
    cond, _ = search.point_in_tet(flat.cuda(), tet.cuda(), pts.cuda())
    # with zero-volume tets the reference predicates accept points arbitrarily far from the tet (both sign tests false,
    # check_condition_tet_for.cu:105-121): such tets are not pruned by their bounding box but tested against every point
    assert np.array_equal(cond.cpu().numpy(), ref)


def test_point_in_tet_collapsed_nan_and_sliver_tets_follow_the_reference():
    """ADVICE r1 / VERDICT r1: a collapsed tet (identical vertices) accepts EVERY point in the reference and wins if its id comes
    first; a tet with a NaN vertex likewise; non-finite query points are 'contained' in the first consistently oriented tet."""
    from deftet_b200 import search
    g, pos, tet = deformed_grid(12, 2, seed=5)
    T = tet.shape[0]
    ta, tb, tc = T // 3, T // 2, (2 * T) // 3
    pos = pos.clone()
    tet = tet.clone()
    gen = torch.Generator().manual_seed(2)
    pts = (torch.rand(2, 3000, 3, generator=gen) - 0.5) * 1.1
    # sample 0: tet ta collapses to a point (all four ids the same vertex), tet tb becomes exactly coplanar
    tet[ta] = tet[ta, 0]
    # sample-independent topology, so poison coordinates per sample instead: a far-away copy of one vertex for the sliver
    v = tet[tb]
    pos[0, v[3]] = (pos[0, v[0]] + pos[0, v[1]] + pos[0, v[2]]) / 3.0
    pos[1, tet[tc, 2]] = float("nan")                                          # sample 1: NaN vertex (poisons every tet around it)
    pts[0, 5] = float("nan"); pts[0, 6, 1] = float("inf"); pts[1, 7, 2] = float("-inf")
    soup = orc_e.gather_tets(pos, tet).numpy()
    ref = orc.point_in_tet(soup, pts.numpy())
    cond, _ = search.point_in_tet(pos.cuda(), tet.cuda(), pts.cuda())
    got = cond.cpu().numpy()
    assert np.array_equal(got, ref), (np.argwhere(got != ref)[:10], got[got != ref][:10], ref[got != ref][:10])
    assert (ref[0] == ta).sum() > 100                # the collapsed tet really does capture points of sample 0
    cond2 = search.point_in_tet_soup(torch.from_numpy(soup).cuda(), pts.cuda())
    assert np.array_equal(cond2.cpu().numpy(), ref)
