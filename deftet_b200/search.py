"""Point-in-tet occupancy query (+ barycentric weights / backward) and 1-NN, over csrc/search.cu.

Host-side mirror of
  ``check_condition_f_base`` = ``TriRender2D.apply``  (reference layers/DefTet/check_condition_tetrahedron_base/utils.py:38-62)
  ``NearestNeighbor`` / ``NearestNeighborFunction``     (reference layers/nearest_neighbor/nearest_neighbor.py:21-60)
plus the engine forms that take (pos, tet) instead of the materialised (B,T,4,3) tensor.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


@_lib.register_signatures
def _search_sigs(lib, sig):
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    sig("dtb_point_in_tet_grid_res", i, i, i)
    sig("dtb_point_in_tet_workspace", sz, i, i, i, i)
    sig("dtb_point_in_tet", i, vp, vp, vp, i, i, i, i, i, vp, vp, vp, sz, vp)
    sig("dtb_point_in_tet_soup", i, vp, vp, i, i, i, i, vp, vp, vp, sz, vp)
    sig("dtb_tet_barycentric_backward", i, vp, vp, vp, vp, vp, i, i, i, i, vp, i, vp, vp)
    sig("dtb_tet_interpolate_forward", i, vp, vp, vp, vp, i, i, i, i, vp, vp)
    sig("dtb_tet_interpolate_backward", i, vp, vp, vp, vp, vp, i, i, i, i, vp, vp, vp)
    sig("dtb_masked_mse_forward", i, vp, vp, vp, i, i, vp, vp, vp)
    sig("dtb_masked_mse_backward", i, vp, vp, vp, vp, vp, i, i, vp, vp)
    sig("dtb_nearest_neighbor_grid_res", i, i)
    sig("dtb_nearest_neighbor_workspace", sz, i, i, i, i)
    sig("dtb_nearest_neighbor", i, vp, vp, vp, i, i, i, i, vp, sz, vp)


def _f32c(t):
    return _lib.aligned(t if t.dtype == torch.float32 else t.float())


# --------------------------------------------------------------------------------------------------- A1
def point_in_tet_soup(tet_bxfx4x3, point_pos_bxnx3, grid_res=0, want_bary=False):
    """Drop-in semantics of ``check_condition_f_base``: -> condition (B,P,1) f32 (tet id or -1)."""
    _lib.require_cuda(tet_bxfx4x3, point_pos_bxnx3)
    if tet_bxfx4x3.dim() != 4 or tet_bxfx4x3.shape[2:] != (4, 3):
        raise RuntimeError("tet_bxfx4x3 must be same im size")          # CHECK_DIM3, check_condition_tet.cpp:40
    if point_pos_bxnx3.dim() != 3 or point_pos_bxnx3.shape[2] != 3 or point_pos_bxnx3.shape[0] != tet_bxfx4x3.shape[0]:
        raise RuntimeError("point_pos_bxnx3 must be same point size")   # CHECK_DIM2, :41
    tet = _f32c(tet_bxfx4x3)
    pts = _f32c(point_pos_bxnx3)
    B, T = tet.shape[0], tet.shape[1]
    P = pts.shape[1]
    dev = pts.device
    L = _lib.lib()
    cond = torch.empty(B, P, 1, device=dev, dtype=torch.float32)
    bary = torch.empty(B, P, 4, device=dev, dtype=torch.float32) if want_bary else None
    wsz = L.dtb_point_in_tet_workspace(B, P, T, grid_res)
    ws = torch.empty(wsz, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_point_in_tet_soup(_lib.ptr(tet), _lib.ptr(pts), B, T, P, grid_res, _lib.ptr(cond), _lib.ptr(bary),
                                           _lib.ptr(ws), wsz, _lib.stream_ptr()), "dtb_point_in_tet_soup")
    return (cond, bary) if want_bary else cond


class _PointInTet(torch.autograd.Function):
    """(pos, tet, points) -> (cond (B,P,1), bary (B,P,4)); bary is differentiable w.r.t. pos and points."""

    @staticmethod
    def forward(ctx, pos, tet32, points, grid_res):
        _lib.require_cuda(pos, tet32, points)
        pos, points = _f32c(pos), _f32c(points)
        B, V, _ = pos.shape
        T, P = tet32.shape[0], points.shape[1]
        dev = pos.device
        L = _lib.lib()
        cond = torch.empty(B, P, 1, device=dev, dtype=torch.float32)
        bary = torch.empty(B, P, 4, device=dev, dtype=torch.float32)
        wsz = L.dtb_point_in_tet_workspace(B, P, T, grid_res)
        ws = torch.empty(wsz, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _lib.check(L.dtb_point_in_tet(_lib.ptr(pos), _lib.ptr(tet32), _lib.ptr(points), B, V, T, P, grid_res,
                                          _lib.ptr(cond), _lib.ptr(bary), _lib.ptr(ws), wsz, _lib.stream_ptr()),
                       "dtb_point_in_tet")
        ctx.save_for_backward(pos, tet32, points, cond)
        ctx.mark_non_differentiable(cond)
        return cond, bary

    @staticmethod
    def backward(ctx, _g_cond, g_bary):
        pos, tet32, points, cond = ctx.saved_tensors
        B, V, _ = pos.shape
        T, P = tet32.shape[0], points.shape[1]
        g_bary = _f32c(g_bary)
        # padded (B,V,4) accumulator -> one vector reduction per vertex update (see include/deftet_b200.h, grad_stride)
        grad_pos = torch.zeros(B, V, 4, device=pos.device) if ctx.needs_input_grad[0] else None
        grad_pts = torch.empty_like(points) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(pos.device):
            _lib.check(_lib.lib().dtb_tet_barycentric_backward(_lib.ptr(pos), _lib.ptr(tet32), _lib.ptr(points), _lib.ptr(cond),
                                                               _lib.ptr(g_bary), B, V, T, P, _lib.ptr(grad_pos), 4,
                                                               _lib.ptr(grad_pts), _lib.stream_ptr()),
                       "dtb_tet_barycentric_backward")
        return (None if grad_pos is None else grad_pos[..., :3]), None, grad_pts, None


def point_in_tet(pos, tet, points, grid_res=0):
    if tet.dtype != torch.int32:
        tet = tet.to(torch.int32)
    return _PointInTet.apply(pos, _lib.aligned(tet), points, int(grid_res))


class _TetInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, field, tet32, cond, bary):
        field, bary = _f32c(field), _f32c(bary)
        B, V, Cn = field.shape
        P = cond.shape[1]
        out = torch.empty(B, P, Cn, device=field.device)
        with torch.cuda.device(field.device):
            _lib.check(_lib.lib().dtb_tet_interpolate_forward(_lib.ptr(field), _lib.ptr(tet32), _lib.ptr(cond), _lib.ptr(bary), B, V, Cn, P,
                                                              _lib.ptr(out), _lib.stream_ptr()), "dtb_tet_interpolate_forward")
        ctx.save_for_backward(field, tet32, cond, bary)
        return out

    @staticmethod
    def backward(ctx, g_out):
        field, tet32, cond, bary = ctx.saved_tensors
        B, V, Cn = field.shape
        P = cond.shape[1]
        g_out = _f32c(g_out)
        g_field = torch.zeros_like(field) if ctx.needs_input_grad[0] else None
        g_bary = torch.empty_like(bary) if ctx.needs_input_grad[3] else None
        with torch.cuda.device(field.device):
            _lib.check(_lib.lib().dtb_tet_interpolate_backward(_lib.ptr(field), _lib.ptr(tet32), _lib.ptr(cond), _lib.ptr(bary),
                                                               _lib.ptr(g_out), B, V, Cn, P, _lib.ptr(g_field), _lib.ptr(g_bary),
                                                               _lib.stream_ptr()), "dtb_tet_interpolate_backward")
        return g_field, None, None, g_bary


def tet_interpolate(field_bxvxc, tet, cond, bary):
    """Interpolate a per-vertex field at the query points located by ``point_in_tet``: -> (B,P,C)."""
    if tet.dtype != torch.int32:
        tet = tet.to(torch.int32)
    return _TetInterpolate.apply(field_bxvxc, _lib.aligned(tet), cond.contiguous(), bary)


class _MaskedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, cond):
        x, target = _f32c(x), _f32c(target)
        B, P = x.shape
        acc = torch.empty(B, 2, device=x.device, dtype=torch.float64)
        loss = torch.empty(B, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().dtb_masked_mse_forward(_lib.ptr(x), _lib.ptr(target), _lib.ptr(cond), B, P, _lib.ptr(acc), _lib.ptr(loss),
                                                         _lib.stream_ptr()), "dtb_masked_mse_forward")
        ctx.save_for_backward(x, target, cond, acc)
        return loss

    @staticmethod
    def backward(ctx, g):
        x, target, cond, acc = ctx.saved_tensors
        B, P = x.shape
        gx = torch.empty_like(x)
        g = _f32c(g)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().dtb_masked_mse_backward(_lib.ptr(x), _lib.ptr(target), _lib.ptr(cond), _lib.ptr(acc), _lib.ptr(g), B, P,
                                                          _lib.ptr(gx), _lib.stream_ptr()), "dtb_masked_mse_backward")
        return gx, None, None


def located_mse(pred_bxp, target_bxp, cond_bxpx1):
    """Mean squared error over the points that fell inside some tet (cond >= 0): -> (B,)."""
    return _MaskedMSE.apply(pred_bxp, target_bxp, cond_bxpx1.reshape(pred_bxp.shape).contiguous())


def paste_occ(pred_tet_occ, condition):
    """``DefTet.paste_occ`` (deftet.py:132-136), including its in-place clamp of misses (-1) to tet 0."""
    condition[condition < 0] = 0
    return torch.gather(input=pred_tet_occ, index=condition.long().squeeze(-1), dim=1)


# --------------------------------------------------------------------------------------------------- A2
def nearest_neighbor_index(queries, points, grid_res=0):
    """(B,Q,3), (B,M,3) -> int32 (B,Q)."""
    _lib.require_cuda(queries, points)
    batch_size, num_queries, dim = queries.shape
    assert dim == 3, "Currently only 3D points are supported"          # nearest_neighbor.py:26
    assert batch_size == points.shape[0]
    assert dim == points.shape[2]
    q, p = _f32c(queries.detach()), _f32c(points.detach())
    M = p.shape[1]
    dev = q.device
    L = _lib.lib()
    result = torch.zeros(batch_size, num_queries, device=dev, dtype=torch.int32)
    wsz = L.dtb_nearest_neighbor_workspace(batch_size, num_queries, M, grid_res)
    ws = torch.empty(wsz, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_nearest_neighbor(_lib.ptr(q), _lib.ptr(p), _lib.ptr(result), batch_size, num_queries, M, grid_res,
                                          _lib.ptr(ws), wsz, _lib.stream_ptr()), "dtb_nearest_neighbor")
    return result


class NearestNeighborFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, queries, points):
        return nearest_neighbor_index(queries, points).long()

    @staticmethod
    def backward(*args):
        raise NotImplementedError


class NearestNeighbor(torch.nn.Module):
    def forward(self, queries, points):
        """queries (B,Q,3), points (B,M,3) -> long (B,Q)"""
        return NearestNeighborFunction.apply(queries, points)
