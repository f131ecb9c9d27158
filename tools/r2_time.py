"""Dev tool (round 2): per-kernel timings at bench scale under the A/B switches of the library.
   python tools/r2_time.py [energies] [nn] [pit] [a4]     -> JSON lines (CUDA events of the library's own profile hooks)"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deftet_b200 import _lib, energies, search, surface
from deftet_b200.engine import GeometryEngine
from deftet_b200.grid import acute_lattice_grid
from deftet_b200.synthetic import analytic_scene

TAGS = {"energies_fwd": 0, "energies_bwd": 1, "pit_tet": 2, "nn_query": 3, "pfd_forward": 4, "bary_bwd": 5}


def kernel_ms(L, tag, fn, iters=8, flush=None):
    ts = []
    for k in range(iters + 2):
        if flush is not None:
            flush.zero_()
        fn()
        torch.cuda.synchronize()
        ms = ctypes.c_float(-1)
        L.dtb_profile_elapsed(TAGS[tag], ctypes.byref(ms))
        if k >= 2:
            ts.append(ms.value)
    return float(np.median(ts)), float(np.min(ts))


def main():
    what = set(sys.argv[1:]) or {"energies", "nn"}
    res = int(os.environ.get("R2_RES", "70"))
    dev = torch.device("cuda:0")
    grid = acute_lattice_grid(res)
    B, P, S = 8, 100000, 100000
    eng = GeometryEngine(grid.centred(), grid.tets, device=dev, max_boundary_faces=16384)
    sc = analytic_scene(grid, B, P, S, 3000, dev)
    L = _lib.lib()
    L.dtb_profile_enable.argtypes = [ctypes.c_int]
    L.dtb_profile_elapsed.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    L.dtb_profile_enable(1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    Fmax = 16384
    faces, counts, _ = surface.boundary_faces(eng.face_table, sc["occ"], Fmax)
    gen = torch.Generator(device=dev).manual_seed(1)
    u = torch.sqrt(torch.rand(B, Fmax, 20, device=dev, generator=gen)); v = torch.rand(B, Fmax, 20, device=dev, generator=gen)
    print(json.dumps({"res": res, "T": eng.n_tet, "V": eng.n_vert, "F_b": counts.tolist(), "nloc_max": eng.tet_tiles.nloc_max}))
    if "energies" in what:
        for path, group in (("direct", ""), ("tiled", "1"), ("tiled", "2"), ("tiled", "4"), ("tiled", "8")):
            os.environ["DTB_ENERGY_PATH"] = path
            if group:
                os.environ["DTB_ENERGY_GROUP"] = group

            def fb():
                p = sc["pos"].detach().requires_grad_(True)
                am, ed, vv = energies.tet_energies(p, eng.tet, eng.inverse_v, tiles=eng.tet_tiles)
                (am + ed + 1e6 * vv).sum().backward()
                return am, ed, vv, p.grad
            out = fb()
            f = kernel_ms(L, "energies_fwd", fb, flush=flush)
            b = kernel_ms(L, "energies_bwd", fb, flush=flush)
            print(json.dumps({"op": "energies", "path": path, "group": group, "fwd_ms": f, "bwd_ms": b,
                              "amips0": float(out[0][0]), "vv0": float(out[2][0]), "gsum": float(out[3].double().abs().sum())}))
        os.environ.pop("DTB_ENERGY_PATH", None); os.environ.pop("DTB_ENERGY_GROUP", None)
    if "nn" in what:
        for kern in ("group", "brick", "thread"):
            os.environ["DTB_NN_KERNEL"] = kern
            for G in ((0, 32, 40, 48, 56, 64, 80, 96, 128) if kern == "group" else (0, 48)):
                def fn():
                    return surface.sample_and_match(sc["pos"], faces, counts, u, v, sc["gt"], G)
                q, nn = fn()
                h = int(nn.long().sum())
                k = kernel_ms(L, "nn_query", fn, flush=flush)
                # whole op (binning included), CUDA events around the call
                ts = []
                for _ in range(6):
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); b.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
                print(json.dumps({"op": "nn", "kernel": kern, "G": G, "kernel_ms": k, "op_ms_median": float(np.median(ts)), "checksum": h}))
        os.environ.pop("DTB_NN_KERNEL", None)
    if "pit" in what:
        for G in (0, 24, 31, 40, 48, 62):
            def fn():
                return search.point_in_tet(sc["pos"], eng.tet, sc["pts"], G)
            c, w = fn()
            k = kernel_ms(L, "pit_tet", fn, flush=flush)
            print(json.dumps({"op": "pit", "G": G, "kernel_ms": k, "checksum": float(c.double().sum())}))
    if "a4" in what:
        for G in (0, 32, 48, 64):
            def fn():
                return surface.closest_faces(sc["pos"], faces, counts, sc["gt"], G)
            s_, cd, cf = fn()
            k = kernel_ms(L, "pfd_forward", fn, flush=flush)
            print(json.dumps({"op": "a4", "G": G, "kernel_ms": k, "checksum": float(cf.double().sum()), "dsum": float(cd.double().sum())}))


if __name__ == "__main__":
    main()
