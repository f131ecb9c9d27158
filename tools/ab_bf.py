"""Dev tool: time surface.boundary_faces (A9) under library variants: python tools/ab_bf.py name1 name2 ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import hashlib, torch
    from deftet_b200 import surface
    from deftet_b200.engine import GeometryEngine
    from deftet_b200.grid import acute_lattice_grid
    from deftet_b200.synthetic import analytic_scene
    from tools.quick_time import timeit
    dev = torch.device("cuda:0")
    grid = acute_lattice_grid(70)
    eng = GeometryEngine(grid.centred(), grid.tets, device=dev)
    scs = [analytic_scene(grid, 8, 1000, 1000, 3000 + k, dev) for k in range(2)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = {"k": 0}
    def fn():
        k = st["k"] % 2; st["k"] += 1
        return surface.boundary_faces(eng.face_table, scs[k]["occ"], 16384)
    h = hashlib.sha1()
    for k in range(2):
        f, c, o = surface.boundary_faces(eng.face_table, scs[k]["occ"], 16384)
        for b in range(8):
            h.update(f[b, :int(c[b])].cpu().numpy().tobytes())
        h.update(c.cpu().numpy().tobytes())
    med, mn = timeit(fn, 30, 5, flush)
    print(json.dumps({"op": "bf", "lib": os.path.basename(os.environ.get("DEFTET_B200_LIB", "default")), "ms_median": med, "ms_min": mn, "sha1": h.hexdigest()[:16]}))
else:
    for name in sys.argv[1:]:
        env = dict(os.environ, DEFTET_B200_LIB=os.path.join(ROOT, "build", "variants", "lib_%s.so" % name))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED %s: %s" % (name, r.stderr[-300:]))
