"""Dev tool: launch each A/B kernel variant once at bench scale (for an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deftet_b200 import energies, surface
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
    for path, group in (("direct", ""), ("tiled", "2"), ("tiled", "8")):
        os.environ["DTB_ENERGY_PATH"] = path
        if group:
            os.environ["DTB_ENERGY_GROUP"] = group
        p = sc["pos"].detach().requires_grad_(True)
        am, ed, vv = energies.tet_energies(p, eng.tet, eng.inverse_v, tiles=eng.tet_tiles)
        (am + ed + 1e6 * vv).sum().backward()
    for kern, G in (("thread", 0), ("brick", 0), ("brick", 64)):
        os.environ["DTB_NN_KERNEL"] = kern
        surface.sample_and_match(sc["pos"], faces, counts, u, v, sc["gt"], G)
torch.cuda.synchronize()
