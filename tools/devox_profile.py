"""Dev tool: one warm-up and one measured forward+backward of sample_f (three encoder levels, res-70 vertices, batch 8) for ncu:
  ncu --set full --clock-control none -k regex:devox --launch-skip 9 -c 9 -o gpurun_out/devox_full python tools/devox_profile.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from deftet_b200 import devox
from deftet_b200.grid import acute_lattice_grid

res, B = 70, 8
dev = torch.device("cuda:0")
g = acute_lattice_grid(res)
pos = torch.from_numpy(g.centred().astype(np.float32)).to(dev).unsqueeze(0).repeat(B, 1, 1)
pos = (pos + 0.1 / res * (torch.rand_like(pos) - 0.5)).contiguous().requires_grad_(True)
vols = [torch.randn(B, c, r, r, r, device=dev, requires_grad=True) for c, r in ((64, 32), (128, 16), (512, 8))]
go = torch.randn(B, 704, pos.shape[1], device=dev)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for _ in range(2):
    out = devox.sample_f(pos, vols, flags)
    torch.autograd.grad(out, [pos] + vols, go)
torch.cuda.synchronize()
print("ok", out.shape)
