"""deftet_b200 vs the REFERENCE'S OWN CUDA KERNELS executed on the same GPU (oracle/_ref/kernels_cuda, built from the
reference's unmodified __global__ kernels by oracle/build_ref_kernels.sh; see oracle/ref_cuda.py).

The reference's device build contracts multiplies and adds into FMAs (nvcc default, as in its torch cpp_extension build), so on
EXACT ties its choice of index is compiler dependent.  The contract checked here:
  * every index that differs must be a tie: the two candidates are equally good up to fp32 rounding (checked in fp64 / by the
    non-contracted oracle on exactly the differing elements);
  * float outputs agree within 1e-5 relative (the north-star tolerance);
  * adjacency (equality predicates only, no rounding) is bit-identical.
"""
import numpy as np
import pytest
import torch

from oracle import builders as orc_b
from oracle import energies as orc_e
from oracle import native as orc
from oracle import ref_cuda
from oracle import surface as orc_s
from tests.util import deformed_grid, rel_err, sphere_occupancy, sphere_points

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref/kernels_cuda not built (needs /root/reference at build time)")]


def _bary64(soup_bt43, pts_bp3, ids_bp):
    """fp64 barycentric weights of each point in the tet `ids` (B,P) -> (B,P,4)."""
    B, P = ids_bp.shape
    t = torch.gather(soup_bt43.double(), 1, ids_bp.clamp(min=0).long().reshape(B, P, 1, 1).expand(-1, -1, 4, 3))
    e = (t[:, :, :3] - t[:, :, 3:4]).transpose(-1, -2)                       # columns v_i - v_3
    w3 = torch.linalg.solve(e, (pts_bp3.double() - t[:, :, 3]).unsqueeze(-1)).squeeze(-1)
    return torch.cat([w3, 1 - w3.sum(-1, keepdim=True)], dim=-1)


@pytest.mark.parametrize("res,B,P,amp", [(12, 2, 20000, 0.25), (16, 1, 30000, 0.4)])
def test_point_in_tet_matches_reference_device_kernel(res, B, P, amp):
    from deftet_b200 import search
    g, pos, tet = deformed_grid(res, B, seed=res + 1, amp=amp)
    soup = orc_e.gather_tets(pos, tet).cuda()
    gen = torch.Generator().manual_seed(res)
    pts = ((torch.rand(B, P, 3, generator=gen) - 0.5) * 1.05).cuda()
    ref = ref_cuda.point_in_tet(soup, pts)
    out = search.point_in_tet_soup(soup, pts)
    diff = (ref != out).squeeze(-1)
    n_diff = int(diff.sum())
    print("A1 vs reference CUDA: %d of %d ids differ" % (n_diff, B * P))
    if n_diff:
        # a differing id is acceptable only if BOTH tets contain the point up to rounding (point on a shared face / edge)
        assert bool(((ref >= 0) == (out >= 0))[diff.unsqueeze(-1)].all())
        for ids in (ref, out):
            w = _bary64(soup, pts, ids.squeeze(-1))
            assert float(w[diff].min()) > -1e-5
        assert n_diff <= max(2, int(1e-4 * B * P))
    # indexed path (the one the engine uses) gives the same ids as the soup drop-in
    cond, _ = search.point_in_tet(pos.cuda(), tet.cuda().int(), pts)
    assert torch.equal(cond, out)


@pytest.mark.parametrize("B,Q,M,seed", [(2, 20000, 30000, 0), (1, 50000, 4000, 1)])
def test_nearest_neighbor_matches_reference_device_kernel(B, Q, M, seed):
    from deftet_b200 import search
    gen = torch.Generator().manual_seed(seed)
    p = torch.rand(B, M, 3, generator=gen) - 0.5
    p = (p / p.norm(dim=-1, keepdim=True).clamp(min=1e-3) * 0.35).cuda()
    q = ((torch.rand(B, Q, 3, generator=gen) - 0.5) * 1.2).cuda()
    ref = ref_cuda.nearest_neighbor(q, p).long()
    out = search.nearest_neighbor_index(q, p).long()
    diff = ref != out
    n_diff = int(diff.sum())
    print("A2 vs reference CUDA: %d of %d indices differ" % (n_diff, B * Q))
    if n_diff:
        d_ref = (q.double() - torch.gather(p, 1, ref.unsqueeze(-1).expand(-1, -1, 3)).double()).pow(2).sum(-1)
        d_out = (q.double() - torch.gather(p, 1, out.unsqueeze(-1).expand(-1, -1, 3)).double()).pow(2).sum(-1)
        assert float(((d_ref - d_out).abs() / d_ref.clamp(min=1e-30))[diff].max()) < 1e-6
        assert n_diff <= max(2, int(1e-4 * B * Q))


def _surface_scene(res, B, seed, S):
    g, pos, tet = deformed_grid(res, B, seed=seed, amp=0.25)
    centres = [[0.05 * (b - 1), 0.02 * b, -0.03 * b] for b in range(B)]
    radii = [0.22 + 0.06 * b for b in range(B)]
    occ = sphere_occupancy(pos, tet, centres, radii)
    f3, ft2, _, _ = orc_b.tet_to_face(g.n_vert, g.tets)
    bnd = orc_s.get_boundary_index(torch.from_numpy(f3), torch.from_numpy(ft2), occ)
    gt = sphere_points(B, S, centres, radii, seed=seed)
    return pos, bnd, gt


@pytest.mark.parametrize("res,S", [(12, 20000), (20, 30000)])
def test_point_face_distance_matches_reference_device_kernel(res, S):
    from deftet_b200 import surface
    pos, bnd, gt = _surface_scene(res, 2, res, S)
    for b in range(2):
        faces = orc_s.gather_faces(pos[b:b + 1], bnd[b]).cuda()                       # (1,F,3,3)
        F = faces.shape[1]
        pts = gt[b:b + 1].cuda()
        d_ref, f_ref = ref_cuda.point_face_distance(pts, faces)
        dfaces = faces.clone().requires_grad_(True)
        d, f = surface.tet_analytic_distance_f_batch(pts, dfaces, torch.tensor([float(F)]).cuda())
        assert rel_err(d.detach(), d_ref, floor=1e-3) < 1e-5
        diff = (f != f_ref).reshape(-1)
        n_diff = int(diff.sum())
        print("A4 vs reference CUDA (F=%d): %d of %d closest faces differ" % (F, n_diff, S))
        if n_diff:
            # ties (closest feature is an edge / vertex shared by both faces): the reference's face must be exactly as close
            # under the non-contracted evaluation, up to rounding
            idx = torch.nonzero(diff).reshape(-1).cpu()
            p1 = pts[0].cpu()[idx].reshape(-1, 1, 3)
            fr = faces[0].cpu()[f_ref.reshape(-1).cpu().long()[idx]].reshape(-1, 1, 3, 3)
            d_alt, _ = orc.point_face_distance(p1.numpy(), fr.numpy())
            d_our = d.detach().reshape(-1).cpu()[idx].numpy()
            assert float(np.max(np.abs(d_alt.reshape(-1) - d_our) / np.maximum(d_our, 1e-3))) < 1e-5
        # backward on the SAME closest faces (ours) through both implementations
        gen = torch.Generator().manual_seed(b)
        gd = torch.rand(1, S, 1, generator=gen).cuda()
        (d * gd).sum().backward()
        g_ref = ref_cuda.point_face_distance_bwd(pts, faces, f.detach(), gd)
        assert rel_err(dfaces.grad, g_ref) < 1e-5


@pytest.mark.parametrize("res", [10, 16])
def test_face_adjacency_matches_reference_device_kernel(res):
    from deftet_b200 import surface
    pos, bnd, _ = _surface_scene(res, 2, 4, 10)
    for b in range(2):
        face = orc_s.gather_faces(pos[b:b + 1], bnd[b])[0].cuda()
        adj_ref = ref_cuda.face_adjacency(face).cpu().numpy()
        adj_orc, pairs_orc = orc.face_adjacency(face.cpu().numpy())
        assert np.array_equal(adj_ref, adj_orc)                    # reference device kernel == C restatement, bit for bit
        out = surface.tet_face_adj_m_f_idx(face)
        assert np.array_equal(out.cpu().numpy(), pairs_orc)
