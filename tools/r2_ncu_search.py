"""Dev tool: launch the three search kernels (A1, A2 grouped, A4) at bench scale (for an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deftet_b200 import search, surface
from deftet_b200.engine import GeometryEngine
from deftet_b200.grid import acute_lattice_grid
from deftet_b200.synthetic import analytic_scene
dev = torch.device("cuda:0")
grid = acute_lattice_grid(70)
B, P, S = 8, 100000, 100000
eng = GeometryEngine(grid.centred(), grid.tets, device=dev, max_boundary_faces=16384)
sc = analytic_scene(grid, B, P, S, 3000, dev)
Fmax = 16384
faces, counts, _ = surface.boundary_faces(eng.face_table, sc["occ"], Fmax)
gen = torch.Generator(device=dev).manual_seed(1)
u = torch.sqrt(torch.rand(B, Fmax, 20, device=dev, generator=gen)); v = torch.rand(B, Fmax, 20, device=dev, generator=gen)
for rep in range(2):
    search.point_in_tet(sc["pos"], eng.tet, sc["pts"])
    surface.sample_and_match(sc["pos"], faces, counts, u, v, sc["gt"], 0)
    surface.closest_faces(sc["pos"], faces, counts, sc["gt"])
torch.cuda.synchronize()
