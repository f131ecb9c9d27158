"""Dev tool: one launch of every kernel that the bench's ncu summaries of round 1 did not cover (for an ncu --set full capture):
the rest of the bench step (boundary, chamfer, backward kernels, face adjacency, energies), the builders' radix sort, the
rasterizer / compositor, check_sign, the CSR SpMM."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import Step
from deftet_b200 import builders, diffrender, graph, render
from deftet_b200.engine import GeometryEngine
from deftet_b200.grid import acute_lattice_grid
from deftet_b200.synthetic import analytic_scene, icosphere
dev = torch.device("cuda:0")
grid = acute_lattice_grid(70)
B, P, S, Fmax = 8, 100000, 100000, 16384
eng = GeometryEngine(grid.centred(), grid.tets, device=dev, max_boundary_faces=Fmax)
sc = analytic_scene(grid, B, P, S, 3000, dev)
gen = torch.Generator(device=dev).manual_seed(1)
u = torch.sqrt(torch.rand(B, Fmax, 20, device=dev, generator=gen)); v = torch.rand(B, Fmax, 20, device=dev, generator=gen)
step = Step(eng, None, Fmax, 20)
for rep in range(1):
    step.delta.grad = None
    step.forward_backward(sc, u, v, concurrent=False)
    tet = eng.tet
    # A16
    vi, fi = icosphere(5)
    cen = sc["pos"][:, tet.long().reshape(-1)].reshape(B, -1, 4, 3).mean(dim=2)
    render.check_sign(torch.from_numpy(vi * 0.3).to(dev).unsqueeze(0).repeat(B, 1, 1), torch.from_numpy(fi).to(dev), cen)
    # N2
    csr = graph.adjacency_csr(tet, grid.n_vert, True)
    x = torch.rand(B, grid.n_vert, 256, device=dev, requires_grad=True)
    graph.sparse_batch_matmul(csr, x).sum().backward()
    # A15 / config 5: one 800x800 view, K = 300, fused forward + backward
    with tempfile.TemporaryDirectory() as d:
        model = diffrender.Deftet(d, res=40, coef=2.5, feature_dim=4, seed=0, device=dev)
    W = 800
    model.sethw(W, W, 1000)
    focal = 0.5 * W / np.tan(0.5 * 0.6911)
    proj = torch.tensor([focal / (0.5 * W), focal / (0.5 * W), -1.0], device=dev).reshape(3, 1)
    rot = torch.eye(3, device=dev).unsqueeze(0)
    pos = torch.tensor([[0.0, 0.0, 4.0]], device=dev)
    col, mask = model(torch.ones(W, W, dtype=torch.bool, device=dev), rot, pos, proj, diffrender.rendermeshcolor, knum=300)
    (col.sum() + mask.sum()).backward()
    builders.tet_face_adj(tet, grid.n_vert)                 # A13: the longest radix sort of the builders
torch.cuda.synchronize()
