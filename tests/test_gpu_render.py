"""A15 / A16 / A17: rasterizer, inside test and Laplacian vs their oracles (A15/A16: parity unpinned -- the
oracle restates the Kaolin call-site contract, see oracle/render_oracle.c)."""
import numpy as np
import pytest
import torch

from oracle import native as orc
from tests.util import deformed_grid, rel_err

pytestmark = pytest.mark.gpu


def _render_scene(B, F, P, D, seed):
    gen = torch.Generator().manual_seed(seed)
    c = torch.rand(B, F, 1, 2, generator=gen) * 2 - 1
    xy = c + (torch.rand(B, F, 3, 2, generator=gen) - 0.5) * 0.5
    z = -torch.rand(B, F, 3, generator=gen) * 5 - 0.5
    feat = torch.rand(B, F, 3, D, generator=gen)
    pix = torch.rand(B, P, 2, generator=gen) * 2 - 1
    rng = torch.tensor([-4.0, -0.6]).reshape(1, 1, 2).expand(B, P, 2).contiguous()
    return pix, rng, z, xy, feat


@pytest.mark.parametrize("B,F,P,D,K", [(1, 300, 500, 4, 16), (2, 2000, 700, 3, 300), (1, 50, 64, 1, 4)])
def test_sparse_render_forward(B, F, P, D, K):
    from deftet_b200 import render
    pix, rng, z, xy, feat = _render_scene(B, F, P, D, F)
    ref_feat, ref_idx = orc.sparse_render(pix.numpy(), rng.numpy(), z.numpy(), xy.numpy(), feat.numpy(), K)
    out, idx = render.deftet_sparse_render(pix.cuda(), rng.cuda(), z.cuda(), xy.cuda(), feat.cuda(), knum=K)
    assert out.shape == (B, P, K, D) and idx.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert rel_err(out, ref_feat) < 1e-5
    assert (ref_idx >= 0).sum() > 0
    for R in (7, 64):
        out2, idx2 = render.deftet_sparse_render(pix.cuda(), rng.cuda(), z.cuda(), xy.cuda(), feat.cuda(), knum=K, grid_res=R)
        assert np.array_equal(idx2.cpu().numpy(), ref_idx)


def test_sparse_render_backward_matches_autograd():
    from deftet_b200 import render
    B, F, P, D, K = 1, 400, 300, 4, 12
    pix, rng, z, xy, feat = _render_scene(B, F, P, D, 3)
    dxy = xy.cuda().requires_grad_(True)
    dfeat = feat.cuda().requires_grad_(True)
    out, idx = render.deftet_sparse_render(pix.cuda(), rng.cuda(), z.cuda(), dxy, dfeat, knum=K)
    gen = torch.Generator().manual_seed(0)
    g = torch.randn(B, P, K, D, generator=gen)
    (out * g.cuda()).sum().backward()
    # torch autograd through the same barycentric expression, on the hit list found by the kernel (float64)
    rxy = xy.double().requires_grad_(True)
    rfeat = feat.double().requires_grad_(True)
    ci = idx.cpu()
    valid = ci >= 0
    fsel = ci.clamp(min=0)
    fx = rxy[0][fsel[0]]                      # (P,K,3,2)
    ff = rfeat[0][fsel[0]]                    # (P,K,3,D)
    p = pix[0].double().unsqueeze(1)          # (P,1,2)
    a, b, c = fx[..., 0, :], fx[..., 1, :], fx[..., 2, :]
    m, pp, n, q = b[..., 0] - a[..., 0], b[..., 1] - a[..., 1], c[..., 0] - a[..., 0], c[..., 1] - a[..., 1]
    s, t = p[..., 0] - a[..., 0], p[..., 1] - a[..., 1]
    k1, k2, k3 = s * q - n * t, m * t - s * pp, m * q - n * pp
    w1, w2 = k1 / (k3 + 1e-8), k2 / (k3 + 1e-8)
    w0 = 1 - w1 - w2
    ref = (w0.unsqueeze(-1) * ff[..., 0, :] + w1.unsqueeze(-1) * ff[..., 1, :] + w2.unsqueeze(-1) * ff[..., 2, :]) * valid[0].unsqueeze(-1)
    (ref * g[0].double()).sum().backward()
    assert rel_err(out[0], ref.detach()) < 1e-5
    assert rel_err(dfeat.grad, rfeat.grad) < 1e-5
    assert rel_err(dxy.grad, rxy.grad) < 1e-4


def _icosphere(level=3):
    t = (1.0 + 5 ** 0.5) / 2
    v = [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]]
    f = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2],
         [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]]
    v = [np.array(x, dtype=np.float64) / np.linalg.norm(x) for x in v]
    for _ in range(level):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                x = (v[a] + v[b]) / 2
                v.append(x / np.linalg.norm(x))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f = nf
    return np.array(v, dtype=np.float32), np.array(f, dtype=np.int64)


def test_check_sign_sphere_mesh():
    from deftet_b200 import render
    v, f = _icosphere(3)
    verts = np.stack([v * 0.3, v * 0.2 + 0.05]).astype(np.float32)          # B = 2 meshes sharing the face list
    gen = torch.Generator().manual_seed(0)
    pts = (torch.rand(2, 4000, 3, generator=gen) - 0.5) * 0.8
    # lattice-aligned points exercise rays through vertices and edges
    pts[0, :300] = torch.from_numpy(verts[0][:300]) * torch.tensor([1.0, 1.0, 0.5])
    ref = orc.check_sign(verts, f, pts.numpy())
    for R in (512, 64, 5):
        out = render.check_sign(torch.from_numpy(verts).cuda(), torch.from_numpy(f).cuda(), pts.cuda(), hash_resolution=R)
        assert out.dtype == torch.bool and np.array_equal(out.cpu().numpy(), ref)
    # semantic check against the analytic ball away from the faceted surface
    r = pts[0].norm(dim=-1).numpy()
    clear = np.abs(r - 0.3) > 0.01
    assert np.array_equal(ref[0][clear], (r < 0.3)[clear])


def test_laplacian_matches_sparse_mm():
    from deftet_b200 import builders, render
    g, pos, tet = deformed_grid(8, 2, seed=0)
    edges, w = builders.tet_point_adj(tet.cuda(), g.n_vert, normalize=True)
    gen = torch.Generator().manual_seed(1)
    off = torch.randn(2, g.n_vert, 3, generator=gen)
    doff = off.cuda().requires_grad_(True)
    loss = render.laplacian_loss(doff, edges, w)
    (loss * torch.tensor([1.0, 2.0]).cuda()).sum().backward()
    # reference: dense D^-1 A
    A = torch.zeros(g.n_vert, g.n_vert, dtype=torch.float64)
    e = edges.cpu().long()
    A[e[:, 0], e[:, 1]] = w.cpu().double()
    roff = off.double().requires_grad_(True)
    nei = torch.einsum("vj,bjc->bvc", A, roff)
    ref = ((nei - roff) ** 2).sum(-1).sum(-1)
    (ref * torch.tensor([1.0, 2.0], dtype=torch.float64)).sum().backward()
    assert rel_err(loss, ref.detach()) < 1e-5
    assert rel_err(doff.grad, roff.grad) < 1e-5


@pytest.mark.parametrize("K", [8, 40])
def test_render_composite_fused_matches_render_then_peel2mask(K):
    """Fused fast path == deftet_sparse_render followed by the reference's peel2mask (forward and gradients)."""
    from deftet_b200 import render
    from oracle import surface as orc_s
    B, F, P, D = 1, 1500, 600, 4
    pix, rng, z, xy, feat = _render_scene(B, F, P, D, 17)
    feat = feat * 0.6                                            # opacities spread over (0, 0.6): several layers stay visible
    feat[0, :50, :, 0] = 1.5                                     # some opacities beyond the clamp
    feat[0, 50:100, :, 0] = -0.2
    gen = torch.Generator().manual_seed(2)
    gcol, gmask = torch.randn(B, P, D - 1, generator=gen), torch.randn(B, P, 1, generator=gen)
    # unfused chain on the GPU kernels already validated above, composited by torch (float64)
    rxy = xy.cuda().requires_grad_(True)
    rfeat = feat.cuda().requires_grad_(True)
    ims, _ = render.deftet_sparse_render(pix.cuda(), rng.cuda(), z.cuda(), rxy, rfeat, knum=K)
    rc, rm = orc_s.peel2mask(ims.double())
    ((rc * gcol.cuda().double()).sum() + (rm * gmask.cuda().double()).sum()).backward()
    dxy = xy.cuda().requires_grad_(True)
    dfeat = feat.cuda().requires_grad_(True)
    color, mask = render.render_composite(pix.cuda(), rng.cuda(), z.cuda(), dxy, dfeat, knum=K)
    assert color.shape == (B, P, D - 1) and mask.shape == (B, P, 1)
    ((color * gcol.cuda()).sum() + (mask * gmask.cuda()).sum()).backward()
    assert rel_err(color, rc.detach()) < 1e-5 and rel_err(mask, rm.detach()) < 1e-5
    assert rel_err(dfeat.grad, rfeat.grad) < 1e-4
    assert rel_err(dxy.grad, rxy.grad) < 1e-4
