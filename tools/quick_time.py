"""Quick per-op timing at bench scale (CUDA events, L2 flushed between iterations). Dev tool, not the bench."""
import argparse
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deftet_b200 import energies, search
from tests.util import deformed_grid


def timeit(fn, iters=10, warm=3, flush=None):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)), float(np.min(ts))


def numpy_face_table(tets, n_vert):
    """interior faces + their two tets (any order) -- dev helper until the GPU builder lands"""
    loc = np.array([[0, 1, 2], [1, 0, 3], [2, 3, 0], [3, 2, 1]])
    tri = tets[:, loc]                                   # T,4,3
    srt = np.sort(tri, axis=-1).reshape(-1, 3)
    key = (srt[:, 0] * n_vert + srt[:, 2]) * n_vert + srt[:, 1]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    same = ks[1:] == ks[:-1]
    first = order[:-1][same]
    second = order[1:][same]
    return tri.reshape(-1, 3)[first].astype(np.int64), np.stack([first // 4, second // 4], 1).astype(np.int64)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=70)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--points", type=int, default=100000)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g, pos, tet = deformed_grid(a.res, a.batch, seed=1)
    B, V, T = a.batch, g.n_vert, g.n_tet
    print("res", a.res, "B", B, "V", V, "T", T)
    pos = pos.to(dev)
    tet32 = tet.to(dev).to(torch.int32)
    inv = energies.tet_inverse_v(torch.from_numpy(g.centred()).to(dev), tet32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator().manual_seed(0)
    pts = ((torch.rand(B, a.points, 3, generator=gen) - 0.5) * 1.05).to(dev)
    d = torch.randn(B, a.points, 3, generator=gen)
    surf = (d / d.norm(dim=-1, keepdim=True) * 0.3).to(dev)
    q = (surf + 0.01 * torch.randn(B, a.points, 3, generator=gen).to(dev))

    p = pos.clone().requires_grad_(True)

    def e_fwd():
        return energies.tet_energies(p, tet32, inv)

    def e_fwdbwd():
        am, ed, vv = energies.tet_energies(p, tet32, inv)
        (am + ed + vv).sum().backward()

    def pit():
        return search.point_in_tet(pos, tet32, pts)

    def pit_fb():
        c, w = search.point_in_tet(p, tet32, pts)
        (w * w).sum().backward()

    def nn():
        return search.nearest_neighbor_index(q, surf)

    for name, fn in [("energies fwd", e_fwd), ("energies fwd+bwd", e_fwdbwd), ("point_in_tet fwd", pit), ("point_in_tet fwd+bwd", pit_fb),
                     ("nearest_neighbor", nn)]:
        med, mn = timeit(fn, a.iters, flush=flush)
        print("%-24s median %.3f ms  min %.3f ms  -> %.1f k tets/ms" % (name, med, mn, B * T / med / 1e3))
    c, w = search.point_in_tet(pos, tet32, pts)
    print("inside fraction", float((c >= 0).float().mean()))

    # ---- surface stage ----
    from deftet_b200 import surface
    from tests.util import sphere_occupancy
    t0 = time.time()
    f3, ft2 = numpy_face_table(g.tets, V)
    print("numpy face table %.2fs, F_s=%d" % (time.time() - t0, f3.shape[0]))
    occ = sphere_occupancy(pos.cpu(), tet, [[0, 0, 0]] * B, [0.3] * B).to(dev)
    table = surface.FaceTable(torch.from_numpy(f3).to(dev), torch.from_numpy(ft2).to(dev))
    Fmax = 16384
    faces, counts, ovf = surface.boundary_faces(table, occ, Fmax)
    print("boundary counts", counts.tolist(), "overflow", int(ovf.item()))
    S = 20
    u = torch.sqrt(torch.rand(B, Fmax, S, device=dev))
    v = torch.rand(B, Fmax, S, device=dev)

    def bf():
        return surface.boundary_faces(table, occ, Fmax)

    def ch_fb():
        l = surface.surface_chamfer(p, faces, counts, u, v, surf)
        l.sum().backward()

    def sd_fb():
        l = surface.surface_distance(p, faces, counts, surf)
        l.sum().backward()

    def nl_fb():
        l = surface.surface_normal_loss(p, faces, counts)
        l.sum().backward()

    for name, fn in [("boundary_faces", bf), ("chamfer fwd+bwd", ch_fb), ("surface_distance fwd+bwd", sd_fb), ("normal_loss fwd+bwd", nl_fb)]:
        med, mn = timeit(fn, a.iters, flush=flush)
        print("%-24s median %.3f ms  min %.3f ms  -> %.1f k tets/ms" % (name, med, mn, B * T / med / 1e3))


if __name__ == "__main__":
    main()
