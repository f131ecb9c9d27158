"""Synthetic stand-in for one ShapeNet batch -- the workload bench.py, the scale parity tests and the
multi-GPU checks all draw from, so that they see the same data for the same seed.

Per sample:
  * vertex deformation U(-0.25/res, 0.25/res) on interior coordinates (keeps every tet positively oriented);
  * GT shape = union of 1..3 axis-aligned ellipsoids, semi-axes U(0.15, 0.35), centres within +-0.12 of the cube centre
    (the analytic inside test replaces kal.ops.mesh.check_sign, layers/DefTet/deftet.py:33-49);
  * S GT surface points uniform (by area) on the surface of the union: area-weighted thinning of sphere directions mapped
    to each ellipsoid, points buried inside another ellipsoid removed;
  * P SDF query points 1.05 * (U[0,1)^3 - 0.5) as dataloader.py:108;
  * occupancy labels: tet centroid inside (occ), query point inside (target), vertex inside (vfield).

Everything is generated with a seeded CPU torch.Generator (identical on every device) and moved to `device` at the end.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["analytic_scene", "inside_union"]


def inside_union(x, centres, axes):
    """x (..., 3); centres, axes (E, 3) -> bool (...): inside any ellipsoid."""
    d = (x.unsqueeze(-2) - centres) / axes                       # (..., E, 3)
    return ((d * d).sum(-1) < 1.0).any(-1)


def _surface_points(S, centres, axes, g):
    """S points uniform by area on the boundary of the union of ellipsoids."""
    E = centres.shape[0]
    # area element of d -> c + a*d on the unit sphere: |(a_y a_z d_x, a_x a_z d_y, a_x a_y d_z)|
    cof = torch.stack([axes[:, 1] * axes[:, 2], axes[:, 0] * axes[:, 2], axes[:, 0] * axes[:, 1]], dim=1)   # (E,3)
    wmax = float(cof.max())
    kept = []
    n_have = 0
    for _ in range(64):
        for e in range(E):
            d = torch.randn(4 * S, 3, generator=g)
            d = d / d.norm(dim=-1, keepdim=True)
            w = (cof[e] * d).norm(dim=-1) / wmax
            acc = torch.rand(4 * S, generator=g) < w
            x = centres[e] + axes[e] * d[acc]
            if E > 1:
                others = [k for k in range(E) if k != e]
                x = x[~inside_union(x, centres[others], axes[others])]
            kept.append(x)
            n_have += x.shape[0]
        if n_have >= S:
            break
    pool = torch.cat(kept, dim=0)
    perm = torch.randperm(pool.shape[0], generator=g)[:S]
    return pool[perm]


def analytic_scene(grid, B, P, S, seed, device, shapes=(1, 3), info=None):
    """-> dict(pos (B,V,3), occ (B,T), gt (B,S,3), pts (B,P,3), target (B,P), vfield (B,V)), float32 on `device`.
    `shapes` = (lo, hi): number of ellipsoids per sample drawn uniformly from lo..hi (more shapes -> more boundary faces)."""
    g = torch.Generator().manual_seed(int(seed))
    base = torch.from_numpy(grid.centred())
    mask = torch.from_numpy(grid.mask.astype(np.float32))
    res = grid.res
    deform = (torch.rand(B, grid.n_vert, 3, generator=g) * 2 - 1) * (0.25 / res) * mask
    pos = base.unsqueeze(0) + deform
    tet = torch.from_numpy(grid.tets)
    cen = pos[:, tet.reshape(-1)].reshape(B, -1, 4, 3).mean(dim=2)
    pts = (torch.rand(B, P, 3, generator=g) - 0.5) * 1.05                       # dataloader.py:108
    occ, gt, target, vfield, n_shapes = [], [], [], [], []
    lo, hi = int(shapes[0]), int(shapes[1])
    for b in range(B):
        E = int(torch.randint(lo, hi + 1, (1,), generator=g))
        if E <= 3:                                                               # 1-3 shapes, radii U(0.15, 0.35)
            spread, amin, amax = 0.12, 0.15, 0.35
        else:                                                                    # F_b sweep: many small blobs = more surface
            spread, amin, amax = 0.36, 0.08, 0.13
        centres = (torch.rand(E, 3, generator=g) * 2 - 1) * spread
        axes = amin + (amax - amin) * torch.rand(E, 3, generator=g)
        occ.append(inside_union(cen[b], centres, axes).float())
        target.append(inside_union(pts[b], centres, axes).float())
        vfield.append(inside_union(pos[b], centres, axes).float())
        gt.append(_surface_points(S, centres, axes, g))
        n_shapes.append(E)
    out = dict(pos=pos, occ=torch.stack(occ), gt=torch.stack(gt), pts=pts, target=torch.stack(target), vfield=torch.stack(vfield))
    out = {k: v.float().contiguous().to(device) for k, v in out.items()}
    if info is not None:
        info["n_shapes"] = n_shapes
    return out


def icosphere(level=4):
    """Unit icosphere: (V,3) float32 vertices, (F,3) int64 faces (outward winding), 20 * 4**level faces -- a watertight GT mesh
    for the paths that label occupancy with check_sign (DefTet.check_tet_inside_sdfs, layers/DefTet/deftet.py:33-49)."""
    t = (1.0 + 5 ** 0.5) / 2
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1],
                  [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
                  [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)            # 3F edges, (ab, bc, ca) blocks
        uniq, inv = np.unique(e, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        m = inv.reshape(3, -1) + v.shape[0]
        ab, bc, ca = m[0], m[1], m[2]
        v = np.concatenate([v, mid])
        f = np.concatenate([np.stack([f[:, 0], ab, ca], 1), np.stack([f[:, 1], bc, ab], 1), np.stack([f[:, 2], ca, bc], 1), np.stack([ab, bc, ca], 1)])
    return v.astype(np.float32), f
