"""Grid-resolution sweeps for the binned search kernels at bench scale (dev tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import analytic_scene
from deftet_b200 import search, surface
from deftet_b200.engine import GeometryEngine
from deftet_b200.grid import acute_lattice_grid
from tools.quick_time import timeit


def main():
    dev = torch.device("cuda:0")
    grid = acute_lattice_grid(70)
    B, P, S = 8, 100000, 100000
    eng = GeometryEngine(grid.centred(), grid.tets, device=dev)
    sc = analytic_scene(grid, B, P, S, 3000, dev)
    Fmax = 16384
    faces, counts, _ = surface.boundary_faces(eng.face_table, sc["occ"], Fmax)
    print("counts", counts.tolist())
    u = torch.sqrt(torch.rand(B, Fmax, 20, device=dev)); v = torch.rand(B, Fmax, 20, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pos = sc["pos"]
    for G in ((40, 48, 56) if '--nn' in sys.argv else (24, 32, 40, 48, 64, 80, 96)):
        med, _ = timeit(lambda: surface.surface_distance(pos, faces, counts, sc["gt"], G), 5, 2, flush)
        print("A4 fwd  G=%3d  %.3f ms" % (G, med))
    for G in (24, 32, 40, 48, 64, 80):
        med, _ = timeit(lambda: surface.surface_chamfer(pos, faces, counts, u, v, sc["gt"], G), 5, 2, flush)
        print("chamfer fwd G=%3d  %.3f ms" % (G, med))
    if "--a1" in sys.argv:
        for G in (24, 32, 48, 63, 80, 100):
            med, _ = timeit(lambda: search.point_in_tet(pos, eng.tet, sc["pts"], G), 5, 2, flush)
            print("A1 fwd G=%3d  %.3f ms" % (G, med))


if __name__ == "__main__":
    main()
