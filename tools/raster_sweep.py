"""Rasterizer grid-resolution sweep at config-5 scale (dev tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deftet_b200 import builders, render
from deftet_b200.grid import acute_lattice_grid
from tools.quick_time import timeit

dev = torch.device("cuda:0")
g40 = acute_lattice_grid(40)
f3, ft2, fs2, bnd = builders.tet_to_face(g40.n_vert, torch.from_numpy(g40.tets).to(dev))
faces = torch.cat([f3, bnd]).long()
F = faces.shape[0]
vpos = torch.from_numpy(g40.centred()).to(dev) * 2.5
W = 800
focal = 0.5 * W / np.tan(0.5 * 0.6911)
cam = vpos + torch.tensor([0.0, 0.0, -4.0], device=dev)
xy = cam[:, :2] / (-cam[:, 2:3]) * focal / (0.5 * W)
fz = cam[faces][..., 2].unsqueeze(0).contiguous()
fxy = (xy[faces] * 1000).unsqueeze(0).contiguous().requires_grad_(True)
feat = torch.rand(1, F, 3, 4, device=dev, requires_grad=True)
ys, xs = torch.meshgrid(torch.linspace(-1, 1, W, device=dev), torch.linspace(-1, 1, W, device=dev), indexing="ij")
pix = (torch.stack([xs, ys], -1).reshape(1, -1, 2) * 1000).contiguous()
rng = torch.tensor([-1000.0, 0.0], device=dev).reshape(1, 1, 2).expand(1, pix.shape[1], 2).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for K in (64, 300):
    for R in (48, 64, 96, 128, 192, 256):
        with torch.no_grad():
            med, _ = timeit(lambda: render.deftet_sparse_render(pix, rng, fz, fxy, feat, knum=K, grid_res=R), 3, 1, flush)
        print("K=%3d R=%3d fwd %.3f ms" % (K, R, med))
out, idx = render.deftet_sparse_render(pix, rng, fz, fxy, feat, knum=64, grid_res=256)
gout = torch.rand_like(out)
med, _ = timeit(lambda: torch.autograd.grad((out * gout).sum(), (fxy, feat), retain_graph=True), 3, 1, flush)
print("K=64 backward (incl. mul/sum) %.3f ms" % med)
for R in (64, 128, 192):
    with torch.no_grad():
        med, _ = timeit(lambda: render.render_composite(pix, rng, fz, fxy, feat, knum=300, grid_res=R), 3, 1, flush)
    print("fused K=300 R=%3d fwd %.3f ms" % (R, med))
