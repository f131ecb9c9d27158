"""Parity of the engine's batched GPU path against the REFERENCE'S OWN DEVICE KERNELS at BASELINE.json sizes.

TEST / BENCH-CHECKER INFRASTRUCTURE (imports oracle/): used by tests/test_gpu_scale_parity.py and by `bench.py --verify`,
never by the product path.  For one synthetic scene (deftet_b200.synthetic.analytic_scene -- the bench workload) it runs
every index-valued kernel of the step through deftet_b200 AND through oracle/_ref/kernels_cuda (the reference's unmodified
__global__ kernels built for sm_100a, brute force), and the float-valued energies through the torch restatement of the
reference's DefTet methods on CUDA tensors, and reports how far apart they are:

  A1 point-in-tet ids           check_condition_tet_for.cu:124-189       n_diff, every differing id must be a tie
  A2 nearest-neighbour ids      nearest_neighbor_cuda.cu:17-55           n_diff, ties = equal distance in fp64
  A4 closest face + distance    tet_analytic_distance_for.cu:257-307     max rel err of d, n_diff faces (ties re-evaluated); points where the
                                                                         reference's FMA-contracted device build disagrees must equal the
                                                                         non-contracted source (brute-force oracle) BITWISE
  A4 backward                   tet_analytic_distance_back.cu:592-686    max rel err of the gradient vs the non-contracted source (<= 1e-5);
                                                                         the device build is reported beside it
  A5 adjacency table            tet_face_adj_m_for.cu:72-108             bit-identical
  A6-A8 energies + gradient     layers/DefTet/deftet.py:239-338          max rel err

The tie contract is the one of tests/test_gpu_reference_cuda.py: the reference's device build contracts FMAs, so on exact
ties its choice of index is compiler dependent; a differing index is accepted only if both candidates are equally good.
"""
from __future__ import annotations

import numpy as np
import torch


def _bary64(pos_b, tet, pts, ids):
    """fp64 barycentric weights of pts (n,3) in tets `ids` (n,) of one sample."""
    t = pos_b.double()[tet.long()[ids.long()]]                                  # (n,4,3)
    e = (t[:, :3] - t[:, 3:4]).transpose(-1, -2)
    w3 = torch.linalg.solve(e, (pts.double() - t[:, 3]).unsqueeze(-1)).squeeze(-1)
    return torch.cat([w3, 1 - w3.sum(-1, keepdim=True)], dim=-1)


def _rel(a, b, floor=0.0):
    a, b = a.double(), b.double()
    return float((a - b).abs().max()) / max(float(b.abs().max()), floor, 1e-300)


def verify_scene(eng, scene, u, v, samples=None, check_energies=True, strict=True):
    """-> dict of parity figures; raises AssertionError (strict) when a difference is not a tie / exceeds 1e-5."""
    from oracle import energies as orc_e
    from oracle import native as orc
    from oracle import ref_cuda
    from deftet_b200 import energies, search, surface

    if not ref_cuda.available():
        return {"unavailable": "oracle/_ref/kernels_cuda not built (needs /root/reference at build time)"}
    pos, occ, gt, pts = scene["pos"], scene["occ"], scene["gt"], scene["pts"]
    B, V, _ = pos.shape
    T, P, S = eng.n_tet, pts.shape[1], gt.shape[1]
    tet = eng.tet
    rep = {"B": B, "T": T, "P": P, "S": S}
    bad = []

    def need(ok, msg):
        if not ok:
            bad.append(msg)

    # ---- A1: indexed binned kernel vs the reference's O(P*T) scan over the materialised soup ----------------------------
    cond, _ = search.point_in_tet(pos, tet, pts)
    soup = pos[:, tet.long().reshape(-1)].reshape(B, T, 4, 3).contiguous()
    ref = ref_cuda.point_in_tet(soup, pts)
    del soup
    diff = (ref != cond).squeeze(-1)
    n_diff = int(diff.sum())
    rep["A1_ids"] = {"n": B * P, "n_diff": n_diff, "inside_frac": float((cond >= 0).float().mean())}
    if n_diff:
        need(bool(((ref >= 0) == (cond >= 0)).squeeze(-1)[diff].all()), "A1: inside/outside disagreement")
        wmin = 0.0
        for b in range(B):
            m = diff[b]
            if bool(m.any()):
                for ids in (ref[b, m, 0], cond[b, m, 0]):
                    keep = ids >= 0
                    if bool(keep.any()):
                        wmin = min(wmin, float(_bary64(pos[b], tet, pts[b, m][keep], ids[keep]).min()))
        rep["A1_ids"]["tie_min_weight"] = wmin
        need(wmin > -1e-5, "A1: a differing id is not a tie (min fp64 weight %g)" % wmin)
        need(n_diff <= max(2, int(1e-4 * B * P)), "A1: too many ties (%d)" % n_diff)

    # ---- surface stage on the engine's ragged layout --------------------------------------------------------------------
    faces, counts, ovf = surface.boundary_faces(eng.face_table, occ, eng.max_boundary_faces)
    need(int(ovf.item()) == 0, "A9: boundary capacity exceeded")
    cnt = counts.tolist()
    rep["F_b"] = cnt
    Sf = u.shape[2]
    q, nn = surface.sample_and_match(pos, faces, counts, u, v, gt)
    soupf, cd, cf = surface.closest_faces(pos, faces, counts, gt)
    gen = torch.Generator(device=pos.device).manual_seed(11)
    gd = torch.rand(B, S, device=pos.device, generator=gen)
    # the engine's fused op (face soup -> distance -> sqrt-mean, backward scattered to vertices) differentiated by autograd
    pe = pos.detach().clone().requires_grad_(True)
    surface.surface_distance(pe, faces, counts, gt).sum().backward()
    g_engine = pe.grad
    a2 = {"n": 0, "n_diff": 0, "max_tie_rel": 0.0}
    a4 = {"n": 0, "faces": 0, "n_diff_face": 0, "d_max_rel": 0.0, "n_fma_sensitive": 0, "n_fma_sensitive_checked": 0, "n_not_oracle": 0,
          "bwd_max_rel": 0.0, "engine_bwd_max_rel": 0.0, "bwd_vs_ref_kernel_max_rel": 0.0, "bwd_vs_ref_kernel_faces_off": 0}
    a5 = {"faces": 0, "identical": True}
    for b in (range(B) if samples is None else samples):
        nb = int(cnt[b])
        if nb == 0:
            continue
        # A2
        nq = nb * Sf
        qb, gb = q[b:b + 1, :nq].contiguous(), gt[b:b + 1].contiguous()
        r = ref_cuda.nearest_neighbor(qb, gb).long()
        o = nn[b:b + 1, :nq].long()
        d = r != o
        a2["n"] += nq
        nd = int(d.sum())
        a2["n_diff"] += nd
        if nd:
            d_r = (qb.double() - torch.gather(gb, 1, r.unsqueeze(-1).expand(-1, -1, 3)).double()).pow(2).sum(-1)
            d_o = (qb.double() - torch.gather(gb, 1, o.unsqueeze(-1).expand(-1, -1, 3)).double()).pow(2).sum(-1)
            a2["max_tie_rel"] = max(a2["max_tie_rel"], float(((d_r - d_o).abs() / d_r.clamp(min=1e-30))[d].max()))
        # A4 forward.  Contract: bit-identical to the reference SOURCE evaluated without FMA contraction (oracle/deftet_oracle.c, pinned
        # bit-for-bit against the reference's own functions compiled for the host).  The reference's DEVICE build contracts FMAs, which
        # changes its answer where the xy-projected inside test is ill-conditioned (faces with |n_z| ~ 0: k3 ~ 0) -- every point where
        # the two GPU results are not the same up to a tie is therefore re-run through the brute-force oracle and must equal OURS bitwise.
        fb = soupf[b:b + 1, :nb].contiguous()
        d_ref, f_ref = ref_cuda.point_face_distance(gb, fb)
        d_our, f_our = cd[b].reshape(1, S, 1), cf[b].reshape(1, S, 1)
        a4["n"] += S
        scale = max(float(d_ref.max()), 1e-3)
        off = ((d_our - d_ref).abs() > 1e-5 * scale).reshape(-1)
        df = (f_our != f_ref).reshape(-1)
        ndf = int(df.sum())
        a4["n_diff_face"] += ndf
        if ndf:
            idx = torch.nonzero(df).reshape(-1)
            p1 = gb[0, idx].reshape(-1, 1, 3).cpu().numpy()
            fr = fb[0, f_ref.reshape(-1)[idx].long()].reshape(-1, 1, 3, 3).cpu().numpy()
            d_alt, _ = orc.point_face_distance(p1, fr)                      # the reference's face under the non-contracted source
            d_o = d_our.reshape(-1)[idx].cpu().numpy()
            not_tie = torch.from_numpy(np.abs(d_alt.reshape(-1) - d_o) / np.maximum(d_o, 1e-3) > 1e-5).to(off.device)
            off[idx[not_tie]] = True
        n_off = int(off.sum())
        a4["n_fma_sensitive"] += n_off
        if n_off:
            idx = torch.nonzero(off).reshape(-1)[:4096]
            d_or, f_or = orc.point_face_distance(gb[0, idx].reshape(1, -1, 3).cpu().numpy(), fb.cpu().numpy())
            same = (d_or.reshape(-1) == d_our.reshape(-1)[idx].cpu().numpy()) & (f_or.reshape(-1) == f_our.reshape(-1)[idx].cpu().numpy())
            a4["n_fma_sensitive_checked"] += int(idx.numel())
            a4["n_not_oracle"] += int((~same).sum())
        a4["d_max_rel"] = max(a4["d_max_rel"], float(((d_our - d_ref).abs().reshape(-1)[~off].max() / scale)) if n_off < S else 0.0)
        # A4 backward: every implementation differentiates the SAME (our) closest faces.  Ours must equal the non-contracted source
        # (oracle) to 1e-5; the reference's device build is reported beside it (it deviates on the faces whose barycentric weights are
        # ill-conditioned under contraction, see above).
        g1 = gd[b].reshape(1, S, 1).contiguous()
        dfaces = fb.clone().requires_grad_(True)
        d2, _ = surface.tet_analytic_distance_f_batch(gb, dfaces, torch.tensor([float(nb)], device=pos.device))
        (d2 * g1).sum().backward()
        gb_c, fb_c, fo_c = gb.cpu().numpy(), fb.cpu().numpy(), f_our.contiguous().cpu().numpy()
        g_orc = torch.from_numpy(orc.point_face_distance_bwd(gb_c, fb_c, fo_c, g1.cpu().numpy())).to(pos.device)
        a4["bwd_max_rel"] = max(a4["bwd_max_rel"], _rel(dfaces.grad, g_orc))
        g_ref = ref_cuda.point_face_distance_bwd(gb, fb, f_our.contiguous(), g1)                   # (1,nb,3,3)
        e_ref = (dfaces.grad - g_ref).abs().reshape(nb, -1).max(dim=1).values / float(g_ref.abs().max())
        a4["bwd_vs_ref_kernel_max_rel"] = max(a4["bwd_vs_ref_kernel_max_rel"], float(e_ref.max()))
        a4["bwd_vs_ref_kernel_faces_off"] += int((e_ref > 1e-5).sum())
        a4["faces"] += nb
        # engine backward: d mean_i sqrt(d_i + 1e-10) / d vertex = the reference's face-corner gradient (its backward source,
        # upstream 1 / (2 S sqrt(d + 1e-10)) as mesh_utils.py:368-374 + .mean give it) scattered to the vertices in fp64
        up = (0.5 / (S * torch.sqrt(d_our.double() + 1e-10))).float().contiguous()
        gf = torch.from_numpy(orc.point_face_distance_bwd(gb_c, fb_c, fo_c, up.cpu().numpy())).to(pos.device)[0].double()   # (nb,3,3)
        gv = torch.zeros(V, 3, device=pos.device, dtype=torch.float64)
        gv.index_add_(0, faces[b, :nb].long().reshape(-1), gf.reshape(-1, 3))
        a4["engine_bwd_max_rel"] = max(a4["engine_bwd_max_rel"], _rel(g_engine[b], gv))
        # A5
        adj_ref = ref_cuda.face_adjacency(fb[0])
        adj_our = surface.face_adjacency_table(faces=faces[b:b + 1, :nb].contiguous(), counts=counts[b:b + 1].contiguous(), n_vert=V,
                                               want="i32")[0]
        a5["faces"] += nb
        a5["identical"] = a5["identical"] and bool(torch.equal(adj_ref.int(), adj_our))
    rep["A2_ids"], rep["A4"], rep["A5"] = a2, a4, a5
    need(a2["max_tie_rel"] < 1e-6, "A2: differing index is not a tie (%g)" % a2["max_tie_rel"])
    need(a2["n_diff"] <= max(2, int(1e-4 * max(a2["n"], 1))), "A2: too many ties (%d)" % a2["n_diff"])
    need(a4["d_max_rel"] < 1e-5, "A4: distance off by %g" % a4["d_max_rel"])
    need(a4["n_not_oracle"] == 0, "A4: %d points differ from the reference kernel AND from the non-contracted oracle" % a4["n_not_oracle"])
    need(a4["n_fma_sensitive"] <= max(4, int(2e-4 * max(a4["n"], 1))), "A4: too many contraction-sensitive points (%d)" % a4["n_fma_sensitive"])
    need(a4["bwd_vs_ref_kernel_faces_off"] <= max(4, int(0.03 * max(a4["faces"], 1))),
         "A4 backward: too many faces differ from the reference's device build (%d)" % a4["bwd_vs_ref_kernel_faces_off"])
    need(a4["bwd_max_rel"] < 1e-5, "A4 backward off by %g" % a4["bwd_max_rel"])
    need(a4["engine_bwd_max_rel"] < 1e-5, "A4 engine backward off by %g" % a4["engine_bwd_max_rel"])
    need(a5["identical"], "A5: adjacency table differs from the reference kernel")

    # ---- A6-A8: fused kernels vs the reference's torch expression + autograd on the GPU --------------------------------
    if check_energies:
        w = (1.0, 1.0, 1e6)
        refe = orc_e.energies_with_grad(pos, tet.long(), eng.inverse_v, w)
        p = pos.detach().clone().requires_grad_(True)
        am, ed, vv = energies.tet_energies(p, tet, eng.inverse_v)
        (w[0] * am + w[1] * ed + w[2] * vv).sum().backward()
        e = {"amips_rel": _rel(am, refe["amips"]), "edge_rel": _rel(ed, refe["edge"]),
             "volvar_rel": _rel(vv, refe["volvar"]), "grad_rel": _rel(p.grad, refe["grad"])}
        rep["A6_A8"] = e
        need(e["amips_rel"] < 1e-5 and e["edge_rel"] < 1e-5, "A6/A8 energy off (%g, %g)" % (e["amips_rel"], e["edge_rel"]))
        need(e["volvar_rel"] < 2e-5, "A7 off by %g" % e["volvar_rel"])
        need(e["grad_rel"] < 1e-5, "A6-A8 gradient off by %g" % e["grad_rel"])
    rep["ok"] = not bad
    rep["failures"] = bad
    if strict and bad:
        raise AssertionError("parity at scale failed: " + "; ".join(bad) + " -- " + str(rep))
    return rep
