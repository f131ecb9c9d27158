"""Per-row measurements of the SURVEY.md section 8 scope table at the BASELINE.json configs (one JSON line per row).

  python tools/bench_components.py [--quick] > profiles/components_rNN.jsonl     (on the B200 box)

GPU numbers: CUDA events around the public call, median of repeated runs, L2 flushed between runs.  CPU numbers:
the reference's own compiled builders (oracle/_ref, kind "reference") where they exist, else the oracle port."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from deftet_b200 import builders, energies, render, search
from deftet_b200.grid import acute_lattice_grid
from tools.quick_time import timeit


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="comma list of sections: core (A1-A16 rows), n3 (topology/regulariser kernels), diffrender (config-5 step)")
    a = ap.parse_args()
    only = set(x for x in a.only.split(",") if x)
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    from oracle import native
    if not only or "n3" in only:
        section_n3(dev, flush, peak, a.quick)
    if not only or "diffrender" in only:
        section_diffrender(dev, flush, a.quick)
    if not only or "n2" in only:
        section_n2(dev, flush, peak, a.quick)
    if not only or "devox" in only:
        section_devox(dev, flush, peak, a.quick)
    if only and "core" not in only:
        return

    # ---- config 2: res 40, batch 8, energies fwd+bwd (A6-A8) ------------------------------------------------------
    for res, B in ((40, 8), (70, 8), (100, 4)):
        g = acute_lattice_grid(res)
        V, T = g.n_vert, g.n_tet
        pos = torch.from_numpy(g.centred()).to(dev).unsqueeze(0).repeat(B, 1, 1)
        pos = (pos + 0.1 / res * (torch.rand_like(pos) - 0.5)).requires_grad_(True)
        tet = torch.from_numpy(g.tets).to(dev).to(torch.int32)
        inv = energies.tet_inverse_v(torch.from_numpy(g.centred()).to(dev), tet)

        def fb():
            pos.grad = None
            am, ed, vv = energies.tet_energies(pos, tet, inv)
            (am + ed + vv).sum().backward()
        med, _ = timeit(fb, 20, 3, flush)
        by = 2 * 52 * T + 36 * B * V + 8 * B * T
        emit(row="A6-A8 energies fwd+bwd", res=res, batch=B, V=V, T=T, ms=med, tets_per_ms=B * T / med, algorithmic_bytes=by,
             hbm_frac=by / (med * 1e-3) / 1e9 / peak)

        # ---- builders A10-A14 (config 4 = res 100, rebuilt every step) ---------------------------------------------
        if res in (40, 100) or not a.quick:
            rows = {"A10 tet_point_adj": lambda: builders.tet_point_adj(tet, V), "A11 tet_to_face": lambda: builders.tet_to_face(V, tet),
                    "A12 tet_adj_share": lambda: builders.tet_adj_share(tet, V), "A13 tet_face_adj": lambda: builders.tet_face_adj(tet, V)}
            soup = torch.from_numpy(g.centred()[g.tets.reshape(-1)]).to(dev)
            rows["A14 collapse_vertices"] = lambda: builders.collapse_vertices(soup)
            for name, fn in rows.items():
                med, _ = timeit(fn, 5, 2, flush)
                cpu = None
                refname = {"A10": "tet_point_adj", "A12": "tet_adj_share", "A13": "tet_face_adj"}.get(name[:3])
                try:
                    if refname and native.ref_lib(refname) and (res <= 70 or name[:3] != "A13"):
                        cols = {"tet_point_adj": (12, 2), "tet_adj_share": (8, 3), "tet_face_adj": (200, 2)}[refname]
                        t0 = time.perf_counter()
                        native.ref_run_tet_builder(refname, g.tets, V, T * cols[0], cols[1])
                        cpu = {"ms": (time.perf_counter() - t0) * 1e3, "kind": "reference (utils/lib/%s/run.cpp, 1 thread, incl. file I/O of the out-of-process driver)" % refname}
                    elif name[:3] == "A14" and native.ref_lib("colaps_v"):
                        t0 = time.perf_counter()
                        native.ref_colaps_v(soup.cpu().numpy())
                        cpu = {"ms": (time.perf_counter() - t0) * 1e3, "kind": "reference (utils/lib/colaps_v/run.cpp, 1 thread)"}
                except Exception as e:  # pragma: no cover
                    cpu = {"error": str(e)[:100]}
                emit(row=name, res=res, V=V, T=T, ms=med, tets_per_ms=T / med, cpu=cpu)

    # ---- A1 alone at res 70 ---------------------------------------------------------------------------------------------
    g = acute_lattice_grid(70)
    B, P = 8, 100000
    pos = torch.from_numpy(g.centred()).to(dev).unsqueeze(0).repeat(B, 1, 1)
    tet = torch.from_numpy(g.tets).to(dev).to(torch.int32)
    pts = (torch.rand(B, P, 3, device=dev) - 0.5) * 1.05
    med, _ = timeit(lambda: search.point_in_tet(pos, tet, pts), 10, 3, flush)
    by = B * 32 * P + 12 * B * g.n_vert + 16 * g.n_tet
    emit(row="A1 point_in_tet fwd (+bary)", res=70, batch=B, points=P, ms=med, tets_per_ms=B * g.n_tet / med, algorithmic_bytes=by,
         hbm_frac=by / (med * 1e-3) / 1e9 / peak)
    soup_t = pos[:, tet.long().reshape(-1)].reshape(B, -1, 4, 3).contiguous()
    med, _ = timeit(lambda: search.point_in_tet_soup(soup_t, pts), 10, 3, flush)
    emit(row="A1 check_condition_f_base drop-in (materialised tet_bxfx4x3)", res=70, batch=B, points=P, ms=med, tets_per_ms=B * g.n_tet / med)

    # ---- A15 rasterizer: config 5 scale (res-40 grid faces, 800x800 pixels, K=300) ---------------------------------------
    g40 = acute_lattice_grid(40)
    f3, ft2, fs2, bnd = builders.tet_to_face(g40.n_vert, torch.from_numpy(g40.tets).to(dev))
    faces = torch.cat([f3, bnd]).long()
    F = faces.shape[0]
    vpos = torch.from_numpy(g40.centred()).to(dev) * 2.5
    W = 400 if a.quick else 800
    focal = 0.5 * W / np.tan(0.5 * 0.6911)
    cam = vpos + torch.tensor([0.0, 0.0, -4.0], device=dev)
    xy = cam[:, :2] / (-cam[:, 2:3]) * focal / (0.5 * W)
    fz = cam[faces][..., 2].unsqueeze(0).contiguous()
    fxy = (xy[faces] * 1000).unsqueeze(0).contiguous()
    feat = torch.rand(1, F, 3, 4, device=dev)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, W, device=dev), torch.linspace(-1, 1, W, device=dev), indexing="ij")
    pix = (torch.stack([xs, ys], -1).reshape(1, -1, 2) * 1000).contiguous()
    rng = torch.tensor([-1000.0, 0.0], device=dev).reshape(1, 1, 2).expand(1, pix.shape[1], 2).contiguous()
    for K in ((64,) if a.quick else (64, 300)):
        out, idx = render.deftet_sparse_render(pix, rng, fz, fxy, feat, knum=K)
        hits = int((idx >= 0).sum())
        med, _ = timeit(lambda: render.deftet_sparse_render(pix, rng, fz, fxy, feat, knum=K), 3, 1, flush)
        by = pix.shape[1] * K * (16 + 8) + F * (12 + 24 + 48)
        emit(row="A15 deftet_sparse_render fwd", pixels=pix.shape[1], faces=F, K=K, hits_per_pixel=hits / pix.shape[1], ms=med, algorithmic_bytes=by,
             hbm_frac=by / (med * 1e-3) / 1e9 / peak, max_slots_used=int((idx >= 0).sum(-1).max()))
        del out, idx
    # fused render + composite (never writes the (P,K,D) tensor), forward and forward+backward
    fxy_g = fxy.clone().requires_grad_(True)
    feat_g = feat.clone().requires_grad_(True)
    for K in ((64,) if a.quick else (64, 300)):
        with torch.no_grad():
            med, _ = timeit(lambda: render.render_composite(pix, rng, fz, fxy, feat, knum=K), 3, 1, flush)
        by = pix.shape[1] * (8 + 8 + 16) + F * (12 + 24 + 48)
        emit(row="A15 fused render_composite fwd", pixels=pix.shape[1], faces=F, K=K, ms=med, algorithmic_bytes=by,
             hbm_frac=by / (med * 1e-3) / 1e9 / peak)

        def fb():
            fxy_g.grad = None; feat_g.grad = None
            c, m = render.render_composite(pix, rng, fz, fxy_g, feat_g, knum=K)
            (c.sum() + m.sum()).backward()
        med, _ = timeit(fb, 3, 1, flush)
        emit(row="A15 fused render_composite fwd+bwd", pixels=pix.shape[1], faces=F, K=K, ms=med)

        def fb_unfused():
            fxy_g.grad = None; feat_g.grad = None
            ims, _ = render.deftet_sparse_render(pix, rng, fz, fxy_g, feat_g, knum=K)
            m = torch.clamp(ims[..., :1], 1e-10, 1 - 1e-10)
            vis = m * torch.cumprod(torch.nn.functional.pad(1 - m[:, :, :-1], (0, 0, 1, 0), value=1.0), dim=2)
            ((ims[..., 1:] * vis).sum(2).sum() + vis.sum()).backward()
        if K <= 64:
            med, _ = timeit(fb_unfused, 3, 1, flush)
            emit(row="A15 drop-in deftet_sparse_render + torch peel2mask fwd+bwd", pixels=pix.shape[1], faces=F, K=K, ms=med)
    # ---- A16 check_sign: sphere mesh, T centroids -----------------------------------------------------------------------------
    from tests.test_gpu_render import _icosphere
    v, f = _icosphere(5)
    verts = torch.from_numpy(v * 0.3).to(dev).unsqueeze(0)
    cen = pos[:1, tet.long().reshape(-1)].reshape(1, -1, 4, 3).mean(2).contiguous()
    med, _ = timeit(lambda: render.check_sign(verts, torch.from_numpy(f).to(dev), cen), 5, 2, flush)
    emit(row="A16 check_sign", mesh_faces=int(f.shape[0]), points=int(cen.shape[1]), ms=med, tets_per_ms=cen.shape[1] / med)


def section_n3(dev, flush, peak, quick):
    """N3 rows: topology editing + regularisers + fused projection at the shipped-grid scales (res 40 / 60)."""
    from deftet_b200 import topology
    for res in ((40,) if quick else (40, 60)):
        g = acute_lattice_grid(res)
        V, T = g.n_vert, g.n_tet
        tet = torch.from_numpy(g.tets).to(dev).to(torch.int32)
        pts = torch.from_numpy(g.centred()).to(dev)
        feat = torch.rand(V, 7, device=dev)
        med, _ = timeit(lambda: topology.tet_edges(tet, V), 5, 2, flush)
        emit(row="N3 tet_edges (generate_edge + generate_tet_edge_idx)", res=res, V=V, T=T, ms=med, tets_per_ms=T / med)
        med, _ = timeit(lambda: topology.generate_subdivision(tet, pts, feat, None), 5, 2, flush)
        p2, f2, t2 = topology.generate_subdivision(tet, pts, feat, None)
        emit(row="N3 generate_subdivision (all tets 1->8)", res=res, V=V, T=T, V_out=int(p2.shape[0]), T_out=int(t2.shape[0]), ms=med, tets_per_ms=T / med)

        def geometry():
            topology.tet_to_face_idx(V, tet, True); topology.tet_neighbours(tet, V); topology.generate_point_adj_idx(V, tet)
        med, _ = timeit(geometry, 5, 2, flush)
        emit(row="N3 updategeometry (tet_to_face_idx + tet_neighbours + point_adj_idx)", res=res, V=V, T=T, ms=med, tets_per_ms=T / med)
        nbr = topology.tet_neighbours(tet, V)
        w = torch.rand(V, 1, device=dev) ** 4
        med, _ = timeit(lambda: topology.delete_tet_by_weight(tet, w, nbr, 0.3, 3), 5, 2, flush)
        emit(row="N3 deletetet (3 neighbour levels)", res=res, V=V, T=T, ms=med, tets_per_ms=T / med)
        table, adjsum = topology.generate_point_adj_idx(V, tet)
        x = feat.clone().requires_grad_(True)

        def lap():
            x.grad = None
            topology.featlap(x, table, adjsum + 1e-10).sum().backward()
        med, _ = timeit(lap, 10, 3, flush)
        by = 2 * (V * table.shape[1] * 4 + 3 * V * 7 * 4)
        emit(row="N3 get_featlap fwd+bwd (7 channels)", res=res, V=V, M=int(table.shape[1]), ms=med, algorithmic_bytes=by, hbm_frac=by / (med * 1e-3) / 1e9 / peak)
        pp = pts.clone().requires_grad_(True)

        def vol():
            pp.grad = None
            (topology.volume_deviation(pp, tet) ** 2).sum().backward()
        med, _ = timeit(vol, 10, 3, flush)
        by = 2 * (16 * T + 12 * V) + 8 * T + 12 * V
        emit(row="N3 get_volume_variance fwd+bwd", res=res, V=V, T=T, ms=med, tets_per_ms=T / med, algorithmic_bytes=by, hbm_frac=by / (med * 1e-3) / 1e9 / peak)
        faces = topology.tet_to_face_idx(V, tet, True)[0]
        F = faces.shape[0]
        rot = torch.eye(3, device=dev).unsqueeze(0)
        cpos = torch.tensor([[0.0, 0.0, 4.0]], device=dev)
        proj = torch.tensor([2.0, 2.0, -1.0], device=dev)
        ff4 = torch.rand(V, 4, device=dev, requires_grad=True)

        def prj():
            pp.grad = None; ff4.grad = None
            fz, fxy, ff = topology.project_faces(pp, ff4, faces, rot, cpos, proj, 1000.0, True)
            (fxy.sum() + ff.sum()).backward()
        med, _ = timeit(prj, 10, 3, flush)
        by = 2 * (12 * F + F * 3 * (4 + 8 + 16)) + 28 * V * 2
        emit(row="N3 project_faces fwd+bwd (1 view, 4 features)", res=res, V=V, faces=int(F), ms=med, algorithmic_bytes=by, hbm_frac=by / (med * 1e-3) / 1e9 / peak)


def section_n2(dev, flush, peak, quick):
    """N2: GraphConv neighbourhood product A x on the vertex adjacency, 256-wide features (layers/gcn_decoder.py:44-56), next to
    the reference expression (torch.sparse.mm on the transposed/reshaped operand, utils/matrix_utils.py:22-33) on the same GPU."""
    from deftet_b200 import graph
    for res, B in ((70, 8),) if quick else ((40, 8), (70, 8), (100, 4)):
        g = acute_lattice_grid(res)
        V = g.n_vert
        tet = torch.from_numpy(g.tets).to(dev)
        adj = builders.tet_to_adj_sparse(V, tet, normalize=True).coalesce()
        csr = graph.csr_of(adj)
        p = 256
        x = torch.randn(B, V, p, device=dev)
        by = 2 * 4 * B * V * p + 8 * csr.nnz + 4 * V
        med, mn = timeit(lambda: graph.sparse_batch_matmul(adj, x), 10, 3, flush)
        xg = x.clone().requires_grad_(True)

        def fb():
            xg.grad = None
            graph.sparse_batch_matmul(adj, xg).backward(x)
        med_fb, _ = timeit(fb, 10, 3, flush)

        def ref():
            d = x.transpose(0, 1).reshape(V, B * p)
            return torch.sparse.mm(adj, d).reshape(V, B, p).transpose(0, 1)
        med_ref, _ = timeit(ref, 5, 2, flush)
        emit(row="N2 sparse_batch_matmul fwd (A x, 256 features)", res=res, batch=B, V=V, nnz=csr.nnz, ms=med, ms_min=mn, algorithmic_bytes=by,
             hbm_frac=by / (med * 1e-3) / 1e9 / peak, achieved_gbs=by / (med * 1e-3) / 1e9, fwd_bwd_ms=med_fb,
             reference_torch_sparse_mm_ms=med_ref, speedup_vs_torch_sparse=med_ref / med)


def section_devox(dev, flush, peak, quick):
    """N4 (second half): sample_f = trilinear_devoxelize over the three encoder levels of layers/pc_model.py:50 at the grid's vertex
    count (train: decode_pos) and at the tet centroids (inference: decode_occ without the 10 000-centroid mask), next to the
    reference expression itself -- torch F.grid_sample per level + torch.cat -- on the same GPU."""
    from deftet_b200 import devox
    levels = [(64, 32), (128, 16), (512, 8)]
    ctot = sum(c for c, _ in levels)
    for res, B, what in ((70, 8, "vertices"),) if quick else ((40, 8, "vertices"), (70, 8, "vertices"), (70, 2, "centroids")):
        g = acute_lattice_grid(res)
        pts = g.centred() if what == "vertices" else g.centred()[g.tets.reshape(-1)].reshape(-1, 4, 3).mean(axis=1)
        N = pts.shape[0]
        pos = torch.from_numpy(pts.astype(np.float32)).to(dev).unsqueeze(0).repeat(B, 1, 1)
        pos = (pos + 0.1 / res * (torch.rand_like(pos) - 0.5)).contiguous()
        vols = [torch.randn(B, c, r, r, r, device=dev) for c, r in levels]
        by_f = 4 * B * ctot * N + sum(4 * B * c * r ** 3 for c, r in levels) + 12 * B * N
        med, mn = timeit(lambda: devox.sample_f(pos, vols), 10, 3, flush)
        pg = pos.clone().requires_grad_(True)
        vg = [v.clone().requires_grad_(True) for v in vols]
        go = torch.randn(B, ctot, N, device=dev)

        def fb():
            pg.grad = None
            for v in vg:
                v.grad = None
            devox.sample_f(pg, vg).backward(go)
        med_fb, _ = timeit(fb, 10, 3, flush)
        med_f_feat, _ = timeit(lambda: torch.autograd.grad(devox.sample_f(pos, vg), vg, go), 10, 3, flush)      # fwd + d/d volume only
        med_f_pos, _ = timeit(lambda: torch.autograd.grad(devox.sample_f(pg, vols), pg, go), 10, 3, flush)      # fwd + d/d positions only
        med_simple, _ = timeit(lambda: devox.sample_f(pos, vols, devox.SIMPLE), 10, 3, flush)

        def fb_simple():
            pg.grad = None
            for v in vg:
                v.grad = None
            devox.sample_f(pg, vg, devox.SIMPLE).backward(go)
        med_fb_simple, _ = timeit(fb_simple, 5, 2, flush)

        def ref_expr(p, vs):
            pt = (p + 0.5).permute(0, 2, 1)
            outs = []
            for v in vs:
                r = v.shape[-1]
                c = torch.clamp(pt * r, 0, r - 1)
                c = (c * 2 + 1.0) / r - 1.0
                grid = torch.flip(c.permute(0, 2, 1).reshape(B, 1, 1, -1, 3), dims=[-1])
                outs.append(torch.nn.functional.grid_sample(v, grid, padding_mode='border', align_corners=False).squeeze(2).squeeze(2))
            return torch.cat(outs, dim=1)
        med_ref, _ = timeit(lambda: ref_expr(pos, vols), 5, 2, flush)

        def ref_fb():
            pg.grad = None
            for v in vg:
                v.grad = None
            ref_expr(pg, vg).backward(go)
        med_ref_fb, _ = timeit(ref_fb, 5, 2, flush)
        # volume gradient alone, per level and per kernel choice (default, channel owner where it applies, shared atomics, sorted reduction); random order = worst case for runs
        per_level = {}
        for order in ("lattice", "shuffled"):
            p_use = pos if order == "lattice" else pos[:, torch.randperm(N, device=dev)].contiguous()
            for c, r in levels:
                v = torch.randn(B, c, r, r, r, device=dev, requires_grad=True)
                gl = torch.randn(B, c, N, device=dev)
                for fl, tag in ((0, "default"), (devox.NO_SORT, "owner"), (devox.NO_SORT | devox.NO_OWNER, "shared_atomics"), (devox.FORCE_SORT, "sorted")):
                    o = devox.sample_f(p_use, [v], fl)
                    t_, _ = timeit(lambda: torch.autograd.grad(o, v, gl, retain_graph=True), 10, 3, flush)
                    per_level["%s_R%d_C%d_%s" % (order, r, c, tag)] = round(t_, 4)
        emit(row="N4 sample_f volume-gradient kernels at %s" % what, res=res, batch=B, N=N, grad_volume_ms=per_level)
        emit(row="N4 sample_f fwd (trilinear_devoxelize x3 levels, 704 ch) at %s" % what, res=res, batch=B, N=N, ms=med, ms_min=mn,
             algorithmic_bytes=by_f, hbm_frac=by_f / (med * 1e-3) / 1e9 / peak, achieved_gbs=by_f / (med * 1e-3) / 1e9, fwd_bwd_ms=med_fb,
             bwd_algorithmic_bytes=by_f + 12 * B * N, reference_torch_grid_sample_ms=med_ref, reference_torch_grid_sample_fwd_bwd_ms=med_ref_fb,
             speedup_fwd=med_ref / med, speedup_fwd_bwd=med_ref_fb / med_fb, fwd_plus_grad_volume_ms=med_f_feat, fwd_plus_grad_positions_ms=med_f_pos,
             one_point_per_thread_kernels_ms=med_simple, one_point_per_thread_kernels_fwd_bwd_ms=med_fb_simple)


def section_diffrender(dev, flush, quick):
    """Config 5 (BASELINE.json configs[4]): one optimisation step of the diff_render loop through the Deftet model mirror -- res-40
    grid x tetcoef 2.5, one 800x800 view, ALL pixels, K=300, L1 image + mask loss, occupancy / Laplacian / volume regularisers,
    backward, Adam step (6_optim/optim_with_mask_subdiv_from_gridmov.py:186-283)."""
    import tempfile
    from deftet_b200 import diffrender
    W = 400 if quick else 800
    K = 300
    with tempfile.TemporaryDirectory() as d:
        model = diffrender.Deftet(d, res=40, coef=2.5, feature_dim=4, seed=0, device=dev)
    model.sethw(W, W, 1000)
    focal = 0.5 * W / np.tan(0.5 * 0.6911)
    proj = torch.tensor([focal / (0.5 * W), focal / (0.5 * W), -1.0], device=dev).reshape(3, 1)
    n_view = 8
    th = np.linspace(0, 2 * np.pi, n_view, endpoint=False)
    rots = torch.from_numpy(np.stack([np.array([[np.cos(t), 0, np.sin(t)], [0, 1, 0], [-np.sin(t), 0, np.cos(t)]]) for t in th]).astype(np.float32)).to(dev)
    poss = torch.stack([rots[b].t() @ torch.tensor([0.0, 0.0, 4.0], device=dev) for b in range(n_view)])
    sample = torch.ones(W, W, dtype=torch.bool, device=dev)
    gt_im = torch.rand(1, W * W, 3, device=dev)
    gt_mask = (torch.rand(1, W * W, 1, device=dev) > 0.5).float()
    opt_g = torch.optim.Adam(list(model.parameters())[1:], lr=1e-2)
    opt_d = torch.optim.Adam(list(model.parameters())[:1], lr=1e-4)
    wvec = torch.tensor([1.0, 1.0, 1.0, 1.0, 10.0, 10.0, 10.0], device=dev)
    state = {"k": 0}

    def step():
        k = state["k"] % n_view
        state["k"] += 1
        opt_g.zero_grad(set_to_none=True); opt_d.zero_grad(set_to_none=True)
        col, mask = model(sample, rots[k:k + 1], poss[k:k + 1], proj, diffrender.rendermeshcolor, knum=K)
        loss = torch.nn.functional.l1_loss(col, gt_im) + torch.nn.functional.l1_loss(mask, gt_mask)
        w, c = diffrender.preprocess_save(None, model.get_feat())
        mov = model.get_mov()
        loss = loss + 1e-3 * w.mean() + 1e-2 * mov.abs().mean() + 1e2 * (model.get_volume_variance() ** 2).sum()
        lap = model.get_featlap(torch.cat([c, w, mov], dim=-1)).sum(0)
        loss = loss + torch.dot(lap, wvec) * 1e-4
        loss.backward()
        opt_g.step(); opt_d.step()
    med, mn = timeit(step, 8, 3, flush)
    F, T, V = int(model.tff_fx3.shape[0]), int(model.tftet_tx4.shape[0]), int(model.n_point)
    emit(row="config 5: diff_render optimisation step (fused projection + rasterizer + compositor, regularisers, Adam)", pixels=W * W, faces=F, T=T,
         V=V, K=K, ms=med, ms_min=mn, views_per_s=1e3 / med, tets_per_ms=T / med)
    with torch.no_grad():
        med, _ = timeit(lambda: model(sample, rots[:1], poss[:1], proj, diffrender.rendermeshcolor, knum=K), 5, 2, flush)
    emit(row="config 5: diff_render forward only (one 800x800 view)" if W == 800 else "config 5 (quick): forward only", pixels=W * W, faces=F, K=K, ms=med)
    # topology edit of the same model: 1->8 subdivision incl. rebuilding every table (Deftet.subdivision)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.subdivision(None)
    torch.cuda.synchronize()
    emit(row="config 5: Deftet.subdivision(None) incl. updategeometry (wall clock)", T_in=T, T_out=int(model.tftet_tx4.shape[0]), V_out=int(model.n_point),
         ms=(time.perf_counter() - t0) * 1e3)


if __name__ == "__main__":
    main()
