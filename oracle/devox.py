"""Oracle for the voxel-feature sampling (second half) -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's live `trilinear_devoxelize` (layers/pv_module/functional/devoxelization.py:47-53) and of
`sample_f` (layers/pc_model.py:182-194).  The reference function is six lines around torch's `F.grid_sample` (torch is the
reference's own pinned dependency, README.md:11-20): `grid = flip((coords*2+1)/r - 1)`, mode bilinear (= trilinear on a volume),
padding_mode 'border', align_corners False.  grid_sample's published algorithm (ATen GridSampler: unnormalise
`((g+1)*size-1)/2`, clip to [0,size-1] with zero gradient on the clipped side, 8 corner weights formed as products of
`(corner+1-x)` / `(x-corner)`, out-of-bounds corners skipped) is restated below without calling it.

PINNED: tests/golden/devox.npz holds inputs, outputs and autograd gradients produced by the reference's own file on CPU
(tests/golden/make_golden_devox.py); tests/test_golden.py checks both functions below against it.
"""
import numpy as np
import torch


def _axis(c, r, xp):
    """coordinate along one axis -> (low index, high index, weight of low, weight of high, d u / d c) in float32."""
    f32 = np.float32
    if xp is np:
        c = c.astype(f32)
        g = (c * f32(2) + f32(1)) / f32(r) - f32(1)                     # devoxelization.py:48
        u = ((g + f32(1)) * f32(r) - f32(1)) / f32(2)                   # unnormalise, align_corners=False
        inside = (u > 0) & (u < f32(r - 1))
        u = np.minimum(f32(r - 1), np.maximum(u, f32(0)))               # padding_mode='border'
        lo = np.floor(u)
        w_hi = u - lo
        w_lo = (lo + f32(1)) - u
        lo_i = lo.astype(np.int64)
        return lo_i, np.minimum(lo_i + 1, r - 1), w_lo, w_hi, inside
    g = (c * 2 + 1.0) / r - 1.0
    u = ((g + 1) * r - 1) / 2
    inside = (u > 0) & (u < r - 1)
    u = torch.where(inside, u, u.detach().clamp(0, r - 1))             # zero gradient where clipped (clip_coordinates_set_grad)
    lo = torch.floor(u).detach()
    w_hi = u - lo
    w_lo = (lo + 1) - u
    lo_i = lo.long()
    return lo_i, torch.clamp(lo_i + 1, max=r - 1), w_lo, w_hi, inside


def trilinear_devoxelize(feat, coords, r):
    """numpy float32, one rounding per operator in ATen's order.  feat (B,C,R,R,R), coords (B,3,N) voxel coordinates -> (B,C,N)."""
    feat = np.asarray(feat, dtype=np.float32)
    coords = np.asarray(coords, dtype=np.float32)
    B, C = feat.shape[:2]
    N = coords.shape[2]
    out = np.zeros((B, C, N), dtype=np.float32)
    for b in range(B):
        l0, h0, g0, f0, _ = _axis(coords[b, 0], r, np)       # slowest volume axis ("t/b" in ATen's corner names) after the flip
        l1, h1, g1, f1, _ = _axis(coords[b, 1], r, np)
        l2, h2, g2, f2, _ = _axis(coords[b, 2], r, np)       # fastest axis (x)
        v = feat[b]
        corners = [(l0, l1, l2, g2 * g1 * g0), (l0, l1, h2, f2 * g1 * g0), (l0, h1, l2, g2 * f1 * g0), (l0, h1, h2, f2 * f1 * g0),
                   (h0, l1, l2, g2 * g1 * f0), (h0, l1, h2, f2 * g1 * f0), (h0, h1, l2, g2 * f1 * f0), (h0, h1, h2, f2 * f1 * f0)]
        acc = np.zeros((C, N), dtype=np.float32)
        for i0, i1, i2, w in corners:                        # tnw, tne, tsw, tse, bnw, bne, bsw, bse
            acc = acc + v[:, i0, i1, i2] * w.astype(np.float32)[None, :]
        out[b] = acc
    return out


def trilinear_devoxelize_torch(feat, coords, r):
    """Differentiable torch-CPU restatement (any float dtype); gradients by autograd."""
    B, C = feat.shape[:2]
    flat = feat.reshape(B, C, -1)
    l0, h0, g0, f0, _ = _axis(coords[:, 0], r, torch)
    l1, h1, g1, f1, _ = _axis(coords[:, 1], r, torch)
    l2, h2, g2, f2, _ = _axis(coords[:, 2], r, torch)
    out = 0
    for i0, i1, i2, w in [(l0, l1, l2, g2 * g1 * g0), (l0, l1, h2, f2 * g1 * g0), (l0, h1, l2, g2 * f1 * g0), (l0, h1, h2, f2 * f1 * g0),
                          (h0, l1, l2, g2 * g1 * f0), (h0, l1, h2, f2 * g1 * f0), (h0, h1, l2, g2 * f1 * f0), (h0, h1, h2, f2 * f1 * f0)]:
        idx = ((i0 * r + i1) * r + i2).unsqueeze(1).expand(B, C, -1)
        out = out + torch.gather(flat, 2, idx) * w.unsqueeze(1)
    return out


def sample_f(point_pos, c_list, fn=trilinear_devoxelize_torch):
    """layers/pc_model.py:182-194, point-cloud branch: point_pos (B,N,3) in [-0.5,0.5]^3, c_list of (B,C_i,R_i,R_i,R_i) -> (B,sum C_i,N)."""
    p = (point_pos + 0.5).permute(0, 2, 1)
    outs = []
    for c in c_list:
        r = c.shape[-1]
        outs.append(fn(c, torch.clamp(p * r, 0, r - 1), r))
    return torch.cat(outs, dim=1)
