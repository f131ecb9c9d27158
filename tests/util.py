"""Shared synthetic-input builders for the parity tests (seeded; sizes the oracle finishes in seconds)."""
import numpy as np
import torch

from deftet_b200.grid import acute_lattice_grid


def deformed_grid(res, B, seed=0, amp=0.25, device="cpu"):
    """Grid + per-sample vertex deformation U(-amp/res, amp/res) on interior coordinates."""
    g = acute_lattice_grid(res)
    gen = torch.Generator().manual_seed(seed)
    base = torch.from_numpy(g.centred())
    mask = torch.from_numpy(g.mask.astype(np.float32))
    delta = (torch.rand(B, g.n_vert, 3, generator=gen) * 2 - 1) * (amp / res)
    pos = base.unsqueeze(0) + delta * mask.unsqueeze(0)
    tet = torch.from_numpy(g.tets)
    return g, pos.float().to(device), tet.to(device)


def rel_err(a, b, floor=0.0):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    scale = max(float(b.abs().max()), floor, 1e-300)
    return float((a - b).abs().max()) / scale


def sphere_occupancy(pos, tet, centres, radii):
    """occ (B,T) float {0,1}: tet centroid inside the sample's sphere (analytic stand-in for kal check_sign)."""
    B = pos.shape[0]
    cen = pos[:, tet.long().reshape(-1)].reshape(B, -1, 4, 3).mean(dim=2)
    c = torch.as_tensor(centres, dtype=pos.dtype).reshape(B, 1, 3)
    r = torch.as_tensor(radii, dtype=pos.dtype).reshape(B, 1)
    return ((cen - c).norm(dim=-1) < r).float()


def sphere_points(B, n, centres, radii, seed=0):
    gen = torch.Generator().manual_seed(seed)
    d = torch.randn(B, n, 3, generator=gen)
    d = d / d.norm(dim=-1, keepdim=True)
    return d * torch.as_tensor(radii).reshape(B, 1, 1) + torch.as_tensor(centres, dtype=torch.float32).reshape(B, 1, 3)
