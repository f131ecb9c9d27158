"""Oracle for the barycentric weights of a point in a tet and their gradient: the reference's pure-torch
``bary_centric_tet`` (utils/tet_utils.py:24-45) restated, differentiated by autograd.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import torch


def _triple(a, b, c):
    return torch.sum(a * torch.linalg.cross(b, c, dim=-1), dim=-1)


def bary_centric_tet(a, b, c, d, p):
    vap, vbp = p - a, p - b
    vab, vac, vad = b - a, c - a, d - a
    vbc, vbd = c - b, d - b
    va6 = _triple(vbp, vbd, vbc)
    vb6 = _triple(vap, vac, vad)
    vc6 = _triple(vap, vad, vab)
    vd6 = _triple(vap, vab, vac)
    v6 = 1 / _triple(vab, vac, vad)
    return va6 * v6, vb6 * v6, vc6 * v6, vd6 * v6


def weights_with_grad(pos, tet_tx4, points, cond, g_w, dtype=torch.float32):
    """pos (B,V,3), points (B,P,3), cond (B,P) tet id or -1, g_w (B,P,4) -> (w (B,P,4), grad_pos, grad_points)."""
    pos = pos.detach().clone().to(dtype).requires_grad_(True)
    points = points.detach().clone().to(dtype).requires_grad_(True)
    B, P = cond.shape
    tid = cond.long().clamp(min=0)
    vid = tet_tx4.long()[tid]                                  # (B,P,4)
    v = torch.gather(pos, 1, vid.reshape(B, -1, 1).expand(-1, -1, 3)).reshape(B, P, 4, 3)
    w = torch.stack(bary_centric_tet(v[:, :, 0], v[:, :, 1], v[:, :, 2], v[:, :, 3], points), dim=-1)
    w = torch.where((cond >= 0).unsqueeze(-1), w, torch.zeros_like(w))
    loss = (w * g_w.to(dtype)).sum()
    gp, gq = torch.autograd.grad(loss, (pos, points))
    return w.detach(), gp, gq
