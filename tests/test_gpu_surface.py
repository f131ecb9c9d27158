"""A9 / A3 / A4 / A5 parity: boundary extraction, chamfer, point->surface distance, face adjacency + normal loss."""
import numpy as np
import pytest
import torch

from oracle import builders as orc_b
from oracle import native as orc
from oracle import surface as orc_s
from tests.util import deformed_grid, rel_err, sphere_occupancy, sphere_points

pytestmark = pytest.mark.gpu


def _scene(res, B, seed, amp=0.25):
    g, pos, tet = deformed_grid(res, B, seed=seed, amp=amp)
    centres = [[0.05 * (b - 1), 0.02 * b, -0.03 * b] for b in range(B)]
    radii = [0.22 + 0.06 * b for b in range(B)]
    occ = sphere_occupancy(pos, tet, centres, radii)
    f3, ft2, _, _ = orc_b.tet_to_face(g.n_vert, g.tets)
    gt = sphere_points(B, 3000, centres, radii, seed=seed)
    return g, pos, tet, occ, torch.from_numpy(f3), torch.from_numpy(ft2), gt


def test_boundary_faces_match_reference_lists():
    from deftet_b200 import surface
    g, pos, tet, occ, f3, ft2, gt = _scene(10, 3, 1)
    occ[2] = 0.0                                              # one sample with an empty surface
    ref = orc_s.get_boundary_index(f3, ft2, occ)
    out = surface.get_boundary_index(f3.cuda(), ft2.cuda(), occ.cuda())
    assert len(out) == 3 and out[2].shape == (0, 3)
    for r, o in zip(ref, out):
        assert o.dtype == torch.int64 and torch.equal(o.cpu(), r)
    # padded-ragged form + overflow flag
    table = surface.FaceTable(f3.cuda(), ft2.cuda())
    faces, counts, overflow = surface.boundary_faces(table, occ.cuda(), 16)
    assert int(overflow.item()) == 1 and counts.tolist() == [16, 16, 0]
    assert torch.equal(faces[0].cpu().long(), ref[0][:16])


@pytest.mark.parametrize("res,amp", [(8, 0.25), (10, 0.0), (12, 0.3)])
def test_analytic_distance_dropin(res, amp):
    """closest_f and closest_d are bit-identical to the brute-force restatement; the undeformed grid has
    faces with k3 == 0 (invisible to the reference) and exactly vertical faces."""
    from deftet_b200 import surface
    g, pos, tet, occ, f3, ft2, gt = _scene(res, 2, res, amp=amp)
    bnd = orc_s.get_boundary_index(f3, ft2, occ)
    for b in range(2):
        faces = orc_s.gather_faces(pos[b:b + 1], bnd[b])
        pts = gt[b:b + 1]
        d_ref, f_ref = orc.point_face_distance(pts.numpy(), faces.numpy())
        n_face = torch.tensor([float(faces.shape[1])])
        dfaces = faces.cuda().requires_grad_(True)
        d, f = surface.tet_analytic_distance_f_batch(pts.cuda(), dfaces, n_face.cuda())
        assert d.shape == (1, pts.shape[1], 1)
        assert np.array_equal(f.cpu().numpy(), f_ref)
        assert np.array_equal(d.detach().cpu().numpy(), d_ref)
        gen = torch.Generator().manual_seed(b)
        gd = torch.rand(1, pts.shape[1], 1, generator=gen)
        (d * gd.cuda()).sum().backward()
        g_ref = orc.point_face_distance_bwd(pts.numpy(), faces.numpy(), f_ref, gd.numpy())
        assert rel_err(dfaces.grad, g_ref) < 1e-5
    # n_face_b smaller than the padded face count: trailing faces are ignored
    faces = orc_s.gather_faces(pos[0:1], bnd[0])
    half = faces.shape[1] // 2
    d_ref, f_ref = orc.point_face_distance(gt[0:1].numpy(), faces.numpy(), np.array([half], dtype=np.float32))
    d, f = surface.tet_analytic_distance_f_batch(gt[0:1].cuda(), faces.cuda(), torch.tensor([float(half)]).cuda())
    assert np.array_equal(f.cpu().numpy(), f_ref) and np.array_equal(d.cpu().numpy(), d_ref)


def test_analytic_distance_vertical_and_degenerate_faces():
    """Random soup including near-vertical / exactly vertical / zero-area triangles and far points."""
    from deftet_b200 import surface
    gen = torch.Generator().manual_seed(3)
    F, S = 400, 2500
    c = torch.rand(1, F, 1, 3, generator=gen) - 0.5
    faces = c + 0.05 * (torch.rand(1, F, 3, 3, generator=gen) - 0.5)
    faces[0, :40, :, 0] = faces[0, :40, 0:1, 0]                       # exactly vertical (normal has no z): k3 == 0
    faces[0, 40:80, 2, :2] = faces[0, 40:80, 0, :2] + 1e-6            # almost vertical
    faces[0, 80:90, 1] = faces[0, 80:90, 0]                           # zero area
    pts = (torch.rand(1, S, 3, generator=gen) - 0.5) * 1.4
    d_ref, f_ref = orc.point_face_distance(pts.numpy(), faces.numpy())
    d, f = surface.tet_analytic_distance_f_batch(pts.cuda(), faces.cuda(), torch.tensor([float(F)]).cuda())
    assert np.array_equal(f.cpu().numpy(), f_ref)
    assert np.array_equal(d.cpu().numpy(), d_ref)


@pytest.mark.parametrize("res,amp", [(8, 0.25), (10, 0.0)])
def test_face_adjacency_dropin(res, amp):
    from deftet_b200 import surface
    g, pos, tet, occ, f3, ft2, gt = _scene(res, 2, 4, amp=amp)
    bnd = orc_s.get_boundary_index(f3, ft2, occ)
    for b in range(2):
        face = orc_s.gather_faces(pos[b:b + 1], bnd[b])[0]
        adj_ref, pairs_ref = orc.face_adjacency(face.numpy())
        out = surface.tet_face_adj_m_f_idx(face.cuda())
        assert out.dtype == torch.int64 and out.shape[0] == 2
        assert np.array_equal(out.cpu().numpy(), pairs_ref)
    assert surface.tet_face_adj_m_f_idx(torch.zeros(0, 3, 3).cuda()).shape == (0,)


def test_face_adjacency_truncates_at_30_in_ascending_order():
    from deftet_b200 import surface
    # a fan of 40 triangles around one shared edge: every face has 39 neighbours, only the 30 smallest survive
    gen = torch.Generator().manual_seed(0)
    apex = torch.rand(40, 3, generator=gen)
    face = torch.zeros(40, 3, 3)
    face[:, 0] = torch.tensor([0.0, 0.0, 0.0])
    face[:, 1] = torch.tensor([1.0, 0.0, -0.0])
    face[:, 2] = apex
    adj_ref, pairs_ref = orc.face_adjacency(face.numpy())
    out = surface.tet_face_adj_m_f_idx(face.cuda())
    assert np.array_equal(out.cpu().numpy(), pairs_ref)


@pytest.mark.parametrize("res,B", [(8, 2), (12, 3)])
def test_engine_surface_losses_and_gradients(res, B):
    """Batched chamfer / surface distance / normal loss + gradients vs the per-sample reference loop."""
    from deftet_b200 import surface
    g, pos, tet, occ, f3, ft2, gt = _scene(res, B, 7)
    if B == 3:
        occ[1] = 0.0
    bnd = orc_s.get_boundary_index(f3, ft2, occ)
    Fmax = max(int(b.shape[0]) for b in bnd) + 3
    S = 20
    gen = torch.Generator().manual_seed(5)
    u = torch.sqrt(torch.rand(B, Fmax, S, 1, generator=gen))
    v = torch.rand(B, Fmax, S, 1, generator=gen)
    w = torch.tensor([[1.0, 0.7, 1.3][:B], [0.5, 1.1, 0.9][:B], [0.8, 1.2, 0.6][:B]])

    rpos = pos.clone().requires_grad_(True)
    ch, an, nl = orc_s.surface_losses(rpos, bnd, gt, [u[b:b + 1, :bnd[b].shape[0]] for b in range(B)],
                                      [v[b:b + 1, :bnd[b].shape[0]] for b in range(B)])
    (w[0] * ch + w[1] * an + w[2] * nl).sum().backward()

    table = surface.FaceTable(f3.cuda(), ft2.cuda())
    faces, counts, overflow = surface.boundary_faces(table, occ.cuda(), Fmax)
    assert int(overflow.item()) == 0
    dpos = pos.cuda().requires_grad_(True)
    dch = surface.surface_chamfer(dpos, faces, counts, u[..., 0].cuda(), v[..., 0].cuda(), gt.cuda())
    dan = surface.surface_distance(dpos, faces, counts, gt.cuda())
    dnl = surface.surface_normal_loss(dpos, faces, counts)
    wc = w.cuda()
    (wc[0] * dch + wc[1] * dan + wc[2] * dnl).sum().backward()
    assert rel_err(dch, ch.detach()) < 1e-5
    assert rel_err(dan, an.detach()) < 1e-5
    assert rel_err(dnl, nl.detach()) < 1e-5
    assert rel_err(dpos.grad, rpos.grad) < 1e-5


def test_engine_concurrent_streams_equal_serial():
    """GeometryEngine.losses on side streams (fork/join) gives the same losses and gradient as the serial order."""
    from deftet_b200.engine import GeometryEngine
    g, pos, tet, occ, f3, ft2, gt = _scene(10, 2, 9)
    eng = GeometryEngine(g.centred(), g.tets, max_boundary_faces=512, device="cuda")
    assert torch.equal(eng.tet_face_fx3.cpu().long(), f3) and torch.equal(eng.tet_face_tetidx_fx2.cpu().long(), ft2)
    gen = torch.Generator().manual_seed(1)
    u = torch.sqrt(torch.rand(2, 512, 20, generator=gen)).cuda()
    v = torch.rand(2, 512, 20, generator=gen).cuda()
    pts = ((torch.rand(2, 3000, 3, generator=gen) - 0.5) * 1.05).cuda()
    res = []
    for conc in (False, True):
        p = pos.cuda().requires_grad_(True)
        out = eng.losses(p, occ.cuda(), gt.cuda(), u, v, pts, concurrent=conc)
        total = (out["amips"] + out["edge"] + out["chamfer"] + out["distance"] + out["normal"] + (out["barycentric"] ** 2).sum(dim=(1, 2))).sum()
        total.backward()
        torch.cuda.synchronize()
        res.append((float(total), p.grad.clone(), out["condition"].clone()))
    assert abs(res[0][0] - res[1][0]) < 1e-5 * abs(res[0][0])
    assert torch.equal(res[0][2], res[1][2])
    assert rel_err(res[1][1], res[0][1]) < 1e-5


@pytest.mark.parametrize("S", [20, 40, 70])
def test_grouped_nearest_neighbour_kernels_agree_with_oracle(S, monkeypatch):
    """The 1-NN of the sampled surface points through the three query kernels -- grouped (default for the chamfer path: one
    warp per face's samples, no query binning), brick (queries sorted by cell) and the per-thread walk of round 1 -- must all be
    the brute-force answer of the oracle, for group sizes below, above and far above one warp."""
    from deftet_b200 import surface
    g, pos, tet, occ, f3, ft2, gt = _scene(10, 2, 13)
    table = surface.FaceTable(f3.cuda(), ft2.cuda())
    Fmax = 700
    faces, counts, overflow = surface.boundary_faces(table, occ.cuda(), Fmax)
    assert int(overflow.item()) == 0
    gen = torch.Generator().manual_seed(S)
    u = torch.sqrt(torch.rand(2, Fmax, S, generator=gen)).cuda()
    v = torch.rand(2, Fmax, S, generator=gen).cuda()
    gtc = gt.cuda()
    gtc[:, 50:60] = gtc[:, 0:10]                      # duplicated targets: the lowest index must win
    outs = {}
    for kern in ("group", "brick", "thread"):
        if kern == "group":
            monkeypatch.delenv("DTB_NN_KERNEL", raising=False)
        else:
            monkeypatch.setenv("DTB_NN_KERNEL", kern)
        for G in (0, 16, 64):
            q, nn = surface.sample_and_match(pos.cuda(), faces, counts, u, v, gtc, G)
            outs[(kern, G)] = nn.clone()
    monkeypatch.delenv("DTB_NN_KERNEL", raising=False)
    cnt = counts.tolist()
    for b in range(2):
        n = cnt[b] * S
        ref = orc.nearest_neighbor(q[b:b + 1, :n].cpu().numpy(), gtc[b:b + 1].cpu().numpy())
        for key, nn in outs.items():
            assert np.array_equal(nn[b, :n].cpu().numpy().astype(np.int64), ref[0]), key


@pytest.mark.parametrize("res,G", [(8, 64), (8, 128), (12, 96), (16, 16)])
def test_analytic_distance_faces_larger_than_a_brick(res, G):
    """Regression (found by tests/test_gpu_scale_parity.py at res 40): when the faces are large against the search grid
    (bounding radius + 1.25 cells > one 4-cell brick) the staged 3x3x3 neighbourhood does not contain the whole reach region;
    the region must be clamped before a query is certified.  Coarse tet grid + fine search grid + points near and far."""
    from deftet_b200 import surface
    g, pos, tet, occ, f3, ft2, gt = _scene(res, 2, 31 + res)
    table = surface.FaceTable(f3.cuda(), ft2.cuda())
    faces, counts, overflow = surface.boundary_faces(table, occ.cuda(), 2048)
    assert int(overflow.item()) == 0
    gen = torch.Generator().manual_seed(G)
    pts = torch.cat([gt, gt * (1.0 + 0.25 * torch.rand(2, gt.shape[1], 1, generator=gen)), (torch.rand(2, 2000, 3, generator=gen) - 0.5)], dim=1)
    soup, cd, cf = surface.closest_faces(pos.cuda(), faces, counts, pts.cuda(), G)
    cnt = counts.tolist()
    for b in range(2):
        d_ref, f_ref = orc.point_face_distance(pts[b:b + 1].numpy(), soup[b:b + 1, :cnt[b]].cpu().numpy())
        assert np.array_equal(cf[b].cpu().numpy(), f_ref.reshape(-1))
        assert np.array_equal(cd[b].cpu().numpy(), d_ref.reshape(-1))


def test_engine_reports_boundary_capacity_overflow():
    """with more boundary faces than `max_boundary_faces` the surface losses are computed on a truncated surface; the engine
    must say so (lazily, without a host synchronisation inside the step): the NEXT call raises."""
    from deftet_b200.engine import GeometryEngine
    g, pos, tet, occ, f3, ft2, gt = _scene(10, 2, 9)
    eng = GeometryEngine(g.centred(), g.tets, max_boundary_faces=16, device="cuda")          # far too small on purpose
    u = torch.rand(2, 16, 20).cuda()
    v = torch.rand(2, 16, 20).cuda()
    out = eng.losses(pos.cuda(), occ.cuda(), gt.cuda(), u, v, None, want=("chamfer",))
    assert int(out["boundary_overflow"].item()) == 1
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="max_boundary_faces"):
        eng.losses(pos.cuda(), occ.cuda(), gt.cuda(), u, v, None, want=("chamfer",))
