"""Dev tool: time one op of the bench workload under several builds of the library (DEFTET_B200_LIB) and check that the results are
bit-identical across them.   python tools/ab_run.py a4|nn|pit name1 name2 ...   (names of build/variants/lib_<name>.so)"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(op):
    sys.path.insert(0, ROOT)
    import torch
    from bench import analytic_scene
    from deftet_b200 import search, surface
    from deftet_b200.engine import GeometryEngine
    from deftet_b200.grid import acute_lattice_grid
    from tools.quick_time import timeit
    dev = torch.device("cuda:0")
    grid = acute_lattice_grid(70)
    B, P, S = 8, 100000, 100000
    eng = GeometryEngine(grid.centred(), grid.tets, device=dev)
    scs = [analytic_scene(grid, B, P, S, 3000 + k, dev) for k in range(2)]
    Fmax = 16384
    fcs = [surface.boundary_faces(eng.face_table, sc["occ"], Fmax)[:2] for sc in scs]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    u = torch.sqrt(torch.rand(B, Fmax, 20, device=dev, generator=gen)); v = torch.rand(B, Fmax, 20, device=dev, generator=gen)
    h = hashlib.sha1()
    state = {"k": 0}
    if op == "a4":
        def fn():
            k = state["k"] % 2; state["k"] += 1
            return surface.surface_distance(scs[k]["pos"], fcs[k][0], fcs[k][1], scs[k]["gt"])
        # drop-in form for the checksum: closest face + distance
        for k in range(2):
            idx = fcs[k][0].long().clamp(min=0)
            soup = torch.gather(scs[k]["pos"].unsqueeze(2).expand(-1, -1, 3, -1), 1, idx.unsqueeze(-1).expand(-1, -1, -1, 3)).contiguous()
            d, f = surface.tet_analytic_distance_f_batch(scs[k]["gt"], soup, fcs[k][1].float())
            h.update(d.cpu().numpy().tobytes()); h.update(f.cpu().numpy().tobytes())
    elif op == "nn":
        def fn():
            k = state["k"] % 2; state["k"] += 1
            return surface.surface_chamfer(scs[k]["pos"], fcs[k][0], fcs[k][1], u, v, scs[k]["gt"], int(os.environ.get("DTB_AB_G", "0")))
        for k in range(2):
            q = scs[k]["gt"] + 0.01
            h.update(search.nearest_neighbor_index(q, scs[k]["gt"]).cpu().numpy().tobytes())
            h.update(surface.surface_chamfer(scs[k]["pos"], fcs[k][0], fcs[k][1], u, v, scs[k]["gt"]).cpu().numpy().tobytes())
    else:
        def fn():
            k = state["k"] % 2; state["k"] += 1
            return search.point_in_tet(scs[k]["pos"], eng.tet, scs[k]["pts"])
        for k in range(2):
            c, w = search.point_in_tet(scs[k]["pos"], eng.tet, scs[k]["pts"])
            h.update(c.cpu().numpy().tobytes()); h.update(w.cpu().numpy().tobytes())
    med, mn = timeit(fn, 20, 4, flush)
    print(json.dumps({"op": op, "lib": os.path.basename(os.environ.get("DEFTET_B200_LIB", "default")), "G": os.environ.get("DTB_AB_G", "auto"), "ms_median": med, "ms_min": mn, "sha1": h.hexdigest()[:16]}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        op = sys.argv[1]
        for name in sys.argv[2:]:
            env = dict(os.environ)
            if "@" in name:                      # name@G: grid resolution override for the op
                name, env["DTB_AB_G"] = name.split("@")
            if name != "default":
                env["DEFTET_B200_LIB"] = os.path.join(ROOT, "build", "variants", "lib_%s.so" % name)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", op], env=env, capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED %s: %s" % (name, r.stderr[-400:]))
