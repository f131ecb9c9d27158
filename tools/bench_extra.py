"""bench.py --config 4 / --config 5: the two BASELINE.json configurations beside the metric's own.

config 4  res 100, GLOBAL batch 4, "adjacency rebuild + vertex collapse every step": every step deletes a (fixed, per input set)
          random 5 % of the tets (mimics Deftet.deletetet), rebuilds the whole topology on the GPU -- face table A11 + rest inverses,
          vertex adjacency A10, tet-tet sharing A12, face-face adjacency A13 -- collapses the (4T,3) tet-soup vertices A14, and runs the
          full loss forward+backward on the rebuilt topology.  Strong scaling of the 4 samples: N <= 4 ranks take 4/N samples each;
          at N = 8 ranks 2k and 2k+1 SHARE sample k: each takes half of its query points, half of its GT points and half of the
          surface samples per boundary face, with loss weight 1/2 (energies, normal and surface-distance terms add up to exactly the
          unsplit loss; the chamfer and the masked occupancy means become the average of two half-sample estimates); the topology rebuild
          and the per-tet energies are replicated work on every rank (recompute instead of broadcasting).
config 5  diff_render: res-40 grid x tetcoef 2.5, 64 cameras on a radius-4 sphere, 800x800, ALL pixels, K = 300; the 64 views are
          sharded over the ranks, each rank renders + back-propagates its views, ONE all-reduce of [d pointmov (V,3), d features
          (V,4)] = 28 V bytes, Adam step on every rank (6_optim/optim_with_mask_subdiv_from_gridmov.py:186-283 runs one random view
          per iteration; a step here is one pass over all 64).

Both print ONE JSON line in bench.py's format (their own metric names; the driver's line is config 3)."""
import json
import os
import time

import numpy as np
import torch


def _init_dist(world, dev, timeout_s=120):
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=timeout_s))
        return dist
    return None


def _timed(run, steps, warmup, world, dist, dev):
    for w in range(max(3, warmup)):
        run(w)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        run(k)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / steps


# ===================================================================================================== config 4
def run_config4(args, rank, world, local_rank, ClockSampler):
    from bench import Step
    from deftet_b200 import builders, energies
    from deftet_b200.engine import GeometryEngine
    from deftet_b200.grid import acute_lattice_grid
    from deftet_b200.synthetic import analytic_scene
    assert world in (1, 2, 4, 8), "config 4 shards a global batch of 4: N in {1,2,4,8}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = _init_dist(world, dev)
    res = args.res or 100
    GB, P, S = 4, args.points, args.points
    grid = acute_lattice_grid(res)
    T, V = grid.n_tet, grid.n_vert
    pair = world == 8                              # two ranks per sample
    B = 1 if pair else GB // world
    lo = rank // 2 if pair else rank * B
    Fmax, S_face = 32768, (10 if pair else 20)
    NSETS = 2
    init_pos = torch.from_numpy(grid.centred()).to(dev)
    tets = torch.from_numpy(grid.tets).to(dev).to(torch.int32)
    sets = []
    for s in range(NSETS):
        full = analytic_scene(grid, GB, P, S, 1000 * 4 + s, dev)             # identical on every rank (same seed), then sharded
        sc = {k: t[lo:lo + B].contiguous() for k, t in full.items()}
        if pair:
            h = rank % 2
            sc["pts"], sc["target"] = sc["pts"][:, h::2].contiguous(), sc["target"][:, h::2].contiguous()
            sc["gt_all"] = sc["gt"]                                          # chamfer targets: every GT point (the queries are split)
            sc["gt"] = sc["gt"][:, h::2].contiguous()
        gen = torch.Generator(device=dev).manual_seed(40 + s)
        keep = torch.rand(T, device=dev, generator=gen) >= 0.05              # the same 5 % on every rank
        sc["keep_idx"] = torch.nonzero(keep).reshape(-1)
        sc["occ_kept"] = sc["occ"][:, sc["keep_idx"]].contiguous()
        gen2 = torch.Generator(device=dev).manual_seed(7 + rank + 100 * s)
        sc["u"] = torch.sqrt(torch.rand(B, Fmax, S_face, device=dev, generator=gen2))
        sc["v"] = torch.rand(B, Fmax, S_face, device=dev, generator=gen2)
        del full
        sets.append(sc)
    delta = torch.zeros(V, 3, device=dev, requires_grad=True)
    stats = {}

    def run(k):
        sc = sets[k % NSETS]
        delta.grad = None
        # ---- delete + rebuild (replicated on every rank) -------------------------------------------------------------------------
        tet_k = tets[sc["keep_idx"]].contiguous()                             # compaction of the tet list
        eng = GeometryEngine(init_pos, tet_k, max_boundary_faces=Fmax, samples_per_face=S_face, device=dev)     # A11 + rest inverses
        edges = builders.tet_point_adj(tet_k, V, normalize=True)              # A10 (+ 1/deg weights)
        share = builders.tet_adj_share(tet_k, V)                              # A12
        fadj = builders.tet_face_adj(tet_k, V)                                # A13
        pos0 = sc["pos"][0] + delta.detach()
        soup = pos0[tet_k.long().reshape(-1)]                                 # (4T,3) tet-soup vertices of sample 0
        cmap, cinv = builders.collapse_vertices(soup)                         # A14
        stats.update(T_kept=int(tet_k.shape[0]), edges=int(edges[0].shape[0]), shared_rows=int(share.shape[0]), face_pairs=int(fadj.shape[0]),
                     collapsed_to=int(cinv.shape[0]), F_s=int(eng.face_table.n_face))
        # ---- full loss on the rebuilt topology -------------------------------------------------------------------------------
        step = Step(eng, None, Fmax, S_face, loss_scale=0.5 if pair else 1.0)
        step.delta = delta
        scene = dict(pos=sc["pos"], occ=sc["occ_kept"], gt=sc["gt"], pts=sc["pts"], target=sc["target"], vfield=sc["vfield"],
                     gt_chamfer=sc.get("gt_all"))
        loss, counts, ovf = step.forward_backward(scene, sc["u"], sc["v"])
        if world > 1:
            dist.all_reduce(delta.grad)
        return loss, ovf

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    loss, ovf = run(0)
    assert int(ovf.item()) == 0, "boundary face capacity exceeded"
    ms = _timed(run, args.steps if args.steps != 200 else 20, args.warmup, world, dist, dev)
    clocks = sampler.stop() if rank == 0 else None
    # builder-only share of the step (rank 0, after the timed region)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    sc = sets[0]
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        tet_k = tets[sc["keep_idx"]].contiguous()
        eng = GeometryEngine(init_pos, tet_k, max_boundary_faces=Fmax, device=dev)
        builders.tet_point_adj(tet_k, V, normalize=True); builders.tet_adj_share(tet_k, V); builders.tet_face_adj(tet_k, V)
        builders.collapse_vertices((sc["pos"][0])[tet_k.long().reshape(-1)])
    ev[1].record()
    torch.cuda.synchronize()
    rebuild_ms = ev[0].elapsed_time(ev[1]) / 5
    steps = args.steps if args.steps != 200 else 20
    line = {"metric": "tets/ms fwd+bwd incl. adjacency rebuild + vertex collapse, res-%d" % res, "value": GB * T / ms, "unit": "tets/ms",
            "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "res=%d GLOBAL batch 4, delete 5 %% of the tets + rebuild A10-A13 + collapse A14 + full loss, every step "
                                   "(BASELINE.json configs[3])" % res, "bench_config": "4", "grid": "synthetic acute lattice V=%d T=%d" % (V, T),
                       "global_batch": GB, "query_points": P, "gt_points": S,
                       "split": ("two ranks per sample: each half of the query points, GT points and surface samples, loss weight 1/2; "
                                 "rebuild + energies replicated" if pair else "%d sample(s) per rank; rebuild replicated" % B),
                       "parallelism": "dp%d" % world, "graph": "eager (the builders size their outputs on the host)", "rebuilt": stats},
            "clocks": clocks, "rebuild_ms": rebuild_ms, "loss": float(loss.item())}
    print(json.dumps(line))


# ===================================================================================================== config 5
def run_config5(args, rank, world, local_rank, ClockSampler):
    import tempfile
    from deftet_b200 import diffrender
    from deftet_b200.dist import GradBucket, shard_range
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = _init_dist(world, dev)
    W, K, NV = 800, 300, 64
    with tempfile.TemporaryDirectory() as d:
        model = diffrender.Deftet(d, res=args.res or 40, coef=2.5, feature_dim=4, seed=0, device=dev)          # same init on every rank
    model.sethw(W, W, 1000)
    focal = 0.5 * W / np.tan(0.5 * 0.6911)                                       # NeRF-synthetic field of view
    proj = torch.tensor([focal / (0.5 * W), focal / (0.5 * W), -1.0], device=dev).reshape(3, 1)
    # 64 cameras on a radius-4 sphere (2_data/load_blender.py:45-52,95-98): azimuth sweep at three elevations
    cams = []
    for i in range(NV):
        th, ph = 2 * np.pi * i / NV, np.deg2rad([-30.0, -10.0, 15.0][i % 3])
        ry = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
        rx = np.array([[1, 0, 0], [0, np.cos(ph), -np.sin(ph)], [0, np.sin(ph), np.cos(ph)]])
        cams.append((rx @ ry).astype(np.float32))
    rots = torch.from_numpy(np.stack(cams)).to(dev)
    poss = torch.stack([rots[b].t() @ torch.tensor([0.0, 0.0, 4.0], device=dev) for b in range(NV)])
    lo, hi = shard_range(NV, rank, world)
    sample = torch.ones(W, W, dtype=torch.bool, device=dev)
    gen = torch.Generator(device=dev).manual_seed(5)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, W, device=dev), torch.linspace(-1, 1, W, device=dev), indexing="ij")
    disc = ((xx ** 2 + yy ** 2) < 0.3).float().reshape(1, W * W, 1)
    gt_mask = disc                                                               # synthetic target: a disc silhouette, flat colours
    gt_im = torch.cat([disc * 0.8, disc * 0.5, disc * 0.2], dim=-1) + (1 - disc)
    params = list(model.parameters())
    opt_g = torch.optim.Adam(params[1:], lr=1e-2)
    opt_d = torch.optim.Adam(params[:1], lr=1e-4)
    bucket = GradBucket(params)
    wvec = torch.tensor([1.0, 1.0, 1.0, 1.0, 10.0, 10.0, 10.0], device=dev)
    VB = 2                                                                       # views per launch

    def run(k):
        opt_g.zero_grad(set_to_none=True); opt_d.zero_grad(set_to_none=True)
        total = None
        for v0 in range(lo, hi, VB):
            v1 = min(v0 + VB, hi)
            col, mask = model(sample, rots[v0:v1], poss[v0:v1], proj, diffrender.rendermeshcolor, knum=K)
            loss = (torch.nn.functional.l1_loss(col, gt_im.expand(v1 - v0, -1, -1)) + torch.nn.functional.l1_loss(mask, gt_mask.expand(v1 - v0, -1, -1))) * ((v1 - v0) / NV)
            loss.backward()                                                      # frees the view chunk's graph before the next one
            total = loss.detach() if total is None else total + loss.detach()
        # regularisers: once per step, 1/world on every rank so that the all-reduced sum counts them once
        w, c = diffrender.preprocess_save(None, model.get_feat())
        mov = model.get_mov()
        reg = 1e-3 * w.mean() + 1e-2 * mov.abs().mean() + 1e2 * (model.get_volume_variance() ** 2).sum()
        lap = model.get_featlap(torch.cat([c, w, mov], dim=-1)).sum(0)
        reg = (reg + torch.dot(lap, wvec) * 1e-4) / world
        reg.backward()
        bucket.all_reduce()                                                      # ONE collective: 28 V bytes
        opt_g.step(); opt_d.step()
        return total + reg.detach()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    steps = args.steps if args.steps != 200 else 5
    ms = _timed(run, steps, min(args.warmup, 3), world, dist, dev)
    clocks = sampler.stop() if rank == 0 else None
    loss = run(0)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    F, T, V = int(model.tff_fx3.shape[0]), int(model.tftet_tx4.shape[0]), int(model.n_point)
    line = {"metric": "views/s diff_render optimisation (64 views 800x800, K=300, fwd+bwd+Adam)", "value": NV / (ms * 1e-3), "unit": "views/s",
            "n_gpus": world, "steps": steps, "warmup": 3, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "diff_render res-40 grid x 2.5, 64 views of 800x800 (all 640 000 pixels), K=300, L1 image + mask loss, "
                                   "regularisers, one %d-byte all-reduce, Adam (BASELINE.json configs[4])" % (28 * V), "bench_config": "5",
                       "V": V, "T": T, "faces": F, "views_per_rank": hi - lo, "views_per_launch": VB, "parallelism": "views sharded dp%d" % world,
                       "graph": "eager"},
            "clocks": clocks, "tets_per_ms": NV * T / ms, "mpixel_per_s": NV * W * W / (ms * 1e-3) / 1e6, "allreduce_bytes": 28 * V, "loss": float(loss.item())}
    print(json.dumps(line))
