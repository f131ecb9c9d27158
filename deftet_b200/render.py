"""Stand-ins for the two Kaolin operations on the hot path, with Kaolin's call signatures:

  deftet_sparse_render(pixel_coords, render_ranges, face_vertices_z, face_vertices_image, face_features, knum=300, eps=1e-8)
      used at diff_render/diftet_6_subdiv/5_rendereq/deftetrneder.py:97-100
  check_sign(verts, faces, points, hash_resolution=512)
      used at layers/DefTet/deftet.py:46, eval.py:239, dataloader.py:92
and the Laplacian smoothness loss of DefTet.laplacian_sparse (layers/DefTet/deftet.py:340-343).

Kaolin is un-vendored and un-pinned in the reference: parity for these two is defined by oracle/render_oracle.c
("parity unpinned")."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .search import _f32c


@_lib.register_signatures
def _render_sigs(lib, sig):
    vp, i, sz, ll, f = C.c_void_p, C.c_int, C.c_size_t, C.c_longlong, C.c_float
    sig("dtb_sparse_render_workspace", sz, i, i, i, i, ll)
    sig("dtb_sparse_render_pair_count", i, vp, vp, i, i, i, i, vp, vp, sz, vp)
    sig("dtb_sparse_render_forward", i, vp, vp, vp, vp, vp, i, i, i, i, i, f, i, ll, vp, vp, vp, vp, sz, vp)
    sig("dtb_sparse_render_backward", i, vp, vp, vp, vp, vp, i, i, i, i, i, f, vp, vp, vp)
    sig("dtb_render_composite_forward", i, vp, vp, vp, vp, vp, i, i, i, i, i, f, i, ll, vp, vp, vp, vp, sz, vp)
    sig("dtb_render_composite_backward", i, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, f, i, ll, vp, vp, vp, sz, vp)
    sig("dtb_check_sign_workspace", sz, i, i, i)
    sig("dtb_check_sign", i, vp, vp, vp, i, i, i, i, i, vp, vp, sz, vp)
    sig("dtb_check_sign_probe", i, vp, vp, vp, i, i, i, i, i, vp, vp, vp, sz, vp)
    sig("dtb_check_sign_fixed", i, vp, vp, vp, i, i, i, i, i, vp, vp, sz, vp)
    sig("dtb_laplacian_forward", i, vp, vp, vp, i, i, i, vp, vp, vp, vp, vp)
    sig("dtb_laplacian_backward", i, vp, vp, vp, vp, vp, i, i, vp, vp)


def _pair_capacity(pix, fxy, grid_res):
    """Exact (cell, face) pair count of the binning (one small kernel sequence + one host read)."""
    B, P, F = pix.shape[0], pix.shape[1], fxy.shape[1]
    n = torch.zeros(1, device=pix.device, dtype=torch.int32)
    wsz = (4 * B + B * F) * 4 + 1024
    ws = torch.empty(wsz, device=pix.device, dtype=torch.uint8)
    _lib.check(_lib.lib().dtb_sparse_render_pair_count(_lib.ptr(pix), _lib.ptr(fxy), B, P, F, grid_res, _lib.ptr(n), _lib.ptr(ws), wsz,
                                                       _lib.stream_ptr()), "dtb_sparse_render_pair_count")
    return max(int(n.item()), 1024)


class _SparseRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pixel_coords, render_ranges, face_vertices_z, face_vertices_image, face_features, knum, eps, grid_res):
        _lib.require_cuda(pixel_coords, face_vertices_image)
        pix, rng = _f32c(pixel_coords), _f32c(render_ranges)
        fz, fxy, ff = _f32c(face_vertices_z), _f32c(face_vertices_image), _f32c(face_features)
        B, P = pix.shape[0], pix.shape[1]
        F, D = fz.shape[1], ff.shape[-1]
        dev = pix.device
        L = _lib.lib()
        out = torch.empty(B, P, knum, D, device=dev)
        idx = torch.empty(B, P, knum, device=dev, dtype=torch.int64)
        overflow = torch.zeros(1, device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            cap = _pair_capacity(pix, fxy, grid_res)
            for _ in range(2):
                wsz = L.dtb_sparse_render_workspace(B, P, F, grid_res, cap)
                ws = torch.empty(wsz, device=dev, dtype=torch.uint8)
                _lib.check(L.dtb_sparse_render_forward(_lib.ptr(pix), _lib.ptr(rng), _lib.ptr(fz), _lib.ptr(fxy), _lib.ptr(ff), B, P, F, D,
                                                       knum, eps, grid_res, cap, _lib.ptr(out), _lib.ptr(idx), _lib.ptr(overflow),
                                                       _lib.ptr(ws), wsz, _lib.stream_ptr()), "dtb_sparse_render_forward")
                if int(overflow.item()) == 0:
                    break
                cap *= 4
            else:
                raise _lib.DeftetB200Error("deftet_sparse_render: face-binning capacity exceeded")
        ctx.save_for_backward(pix, fxy, ff, idx)
        ctx.eps = eps
        ctx.mark_non_differentiable(idx)
        return out, idx

    @staticmethod
    def backward(ctx, g_out, _g_idx):
        pix, fxy, ff, idx = ctx.saved_tensors
        B, P, K = idx.shape
        F, D = fxy.shape[1], ff.shape[-1]
        g_out = _f32c(g_out)
        g_xy = torch.zeros_like(fxy) if ctx.needs_input_grad[3] else None
        g_ff = torch.zeros_like(ff) if ctx.needs_input_grad[4] else None
        with torch.cuda.device(pix.device):
            _lib.check(_lib.lib().dtb_sparse_render_backward(_lib.ptr(pix), _lib.ptr(fxy), _lib.ptr(ff), _lib.ptr(idx), _lib.ptr(g_out), B, P, F,
                                                             D, K, ctx.eps, _lib.ptr(g_xy), _lib.ptr(g_ff), _lib.stream_ptr()),
                       "dtb_sparse_render_backward")
        return None, None, None, g_xy, g_ff, None, None, None


def deftet_sparse_render(pixel_coords, render_ranges, face_vertices_z, face_vertices_image, face_features, knum=300, eps=1e-8,
                         grid_res=0):
    """-> (face_features_out (B,P,knum,d), face_idx (B,P,knum) long)."""
    return _SparseRender.apply(pixel_coords, render_ranges, face_vertices_z, face_vertices_image, face_features, int(knum), float(eps),
                               int(grid_res))


class _RenderComposite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pixel_coords, render_ranges, face_vertices_z, face_vertices_image, face_features, knum, eps, grid_res):
        _lib.require_cuda(pixel_coords, face_vertices_image)
        pix, rng = _f32c(pixel_coords), _f32c(render_ranges)
        fz, fxy, ff = _f32c(face_vertices_z), _f32c(face_vertices_image), _f32c(face_features)
        B, P = pix.shape[0], pix.shape[1]
        F, D = fz.shape[1], ff.shape[-1]
        dev = pix.device
        L = _lib.lib()
        color = torch.empty(B, P, D - 1, device=dev)
        mask = torch.empty(B, P, 1, device=dev)
        overflow = torch.zeros(1, device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            cap = _pair_capacity(pix, fxy, grid_res)
            for _ in range(2):
                wsz = L.dtb_sparse_render_workspace(B, P, F, grid_res, cap)
                ws = torch.empty(wsz, device=dev, dtype=torch.uint8)
                _lib.check(L.dtb_render_composite_forward(_lib.ptr(pix), _lib.ptr(rng), _lib.ptr(fz), _lib.ptr(fxy), _lib.ptr(ff), B, P, F, D,
                                                          knum, eps, grid_res, cap, _lib.ptr(color), _lib.ptr(mask), _lib.ptr(overflow),
                                                          _lib.ptr(ws), wsz, _lib.stream_ptr()), "dtb_render_composite_forward")
                if int(overflow.item()) == 0:
                    break
                cap *= 4
            else:
                raise _lib.DeftetB200Error("render_composite: face-binning capacity exceeded")
        ctx.save_for_backward(pix, rng, fz, fxy, ff, ws)
        ctx.cfg = (knum, eps, grid_res, cap)
        return color, mask

    @staticmethod
    def backward(ctx, g_color, g_mask):
        pix, rng, fz, fxy, ff, ws = ctx.saved_tensors
        knum, eps, grid_res, cap = ctx.cfg
        B, P = pix.shape[0], pix.shape[1]
        F, D = fz.shape[1], ff.shape[-1]
        g_color, g_mask = _f32c(g_color), _f32c(g_mask)
        g_xy = torch.zeros_like(fxy) if ctx.needs_input_grad[3] else None
        g_ff = torch.zeros_like(ff) if ctx.needs_input_grad[4] else None
        with torch.cuda.device(pix.device):
            _lib.check(_lib.lib().dtb_render_composite_backward(_lib.ptr(pix), _lib.ptr(rng), _lib.ptr(fz), _lib.ptr(fxy), _lib.ptr(ff),
                                                                _lib.ptr(g_color), _lib.ptr(g_mask), B, P, F, D, knum, eps, grid_res, cap,
                                                                _lib.ptr(g_xy), _lib.ptr(g_ff), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                       "dtb_render_composite_backward")
        return None, None, None, g_xy, g_ff, None, None, None


def render_composite(pixel_coords, render_ranges, face_vertices_z, face_vertices_image, face_features, knum=300, eps=1e-8, grid_res=0):
    """Fused ``deftet_sparse_render`` + ``peel2mask`` (5_rendereq/deftetrneder.py:31-64,97-113): -> (colour (B,P,d-1), mask (B,P,1))
    with white background, never materialising the (B,P,K,d) tensor.  face_features channel 0 is the opacity."""
    return _RenderComposite.apply(pixel_coords, render_ranges, face_vertices_z, face_vertices_image, face_features, int(knum), float(eps),
                                  int(grid_res))


_CS_RES = {}           # (n_verts, n_faces, hash_resolution) -> grid resolution that fitted the first mesh of that size


def check_sign(verts, faces, points, hash_resolution=512):
    """verts (B,n,3), faces (m,3) long, points (B,p,3) -> bool (B,p): True where the point is inside the mesh.

    The first call for a mesh size finds a grid resolution whose (cell, triangle) list fits (blocking, set-up time); later calls
    with meshes of that size reuse it without any host synchronisation (the training loop labels B samples per step,
    layers/DefTet/deftet.py:33-49) -- a mesh that overflows after all is answered by testing every face on the device."""
    _lib.require_cuda(verts, faces, points)
    v, p = _f32c(verts), _f32c(points)
    f = faces.to(torch.int32).contiguous()
    B, n, m, npts = v.shape[0], v.shape[1], f.shape[0], p.shape[1]
    dev = v.device
    L = _lib.lib()
    out = torch.zeros(B, npts, device=dev, dtype=torch.uint8)
    R = min(int(hash_resolution), 1024)
    wsz = L.dtb_check_sign_workspace(B, m, R)
    ws = torch.empty(wsz, device=dev, dtype=torch.uint8)
    key = (n, m, R)
    with torch.cuda.device(dev):
        if key in _CS_RES:
            _lib.check(L.dtb_check_sign_fixed(_lib.ptr(v), _lib.ptr(f), _lib.ptr(p), B, n, m, npts, _CS_RES[key], _lib.ptr(out), _lib.ptr(ws), wsz,
                                              _lib.stream_ptr()), "dtb_check_sign_fixed")
        else:
            r_used = C.c_int(0)
            _lib.check(L.dtb_check_sign_probe(_lib.ptr(v), _lib.ptr(f), _lib.ptr(p), B, n, m, npts, R, _lib.ptr(out), C.byref(r_used), _lib.ptr(ws),
                                              wsz, _lib.stream_ptr()), "dtb_check_sign_probe")
            _CS_RES[key] = int(r_used.value)
    return out.bool()


class _Laplacian(torch.autograd.Function):
    @staticmethod
    def forward(ctx, offset, edges, weight):
        d = _f32c(offset)
        B, V, _ = d.shape
        E = edges.shape[0]
        dev = d.device
        resid = torch.empty_like(d)
        rows = torch.empty(V + 1, device=dev, dtype=torch.int32)
        acc = torch.empty(B, device=dev, dtype=torch.float64)
        loss = torch.empty(B, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().dtb_laplacian_forward(_lib.ptr(d), _lib.ptr(edges), _lib.ptr(weight), B, V, E, _lib.ptr(resid), _lib.ptr(rows),
                                                        _lib.ptr(acc), _lib.ptr(loss), _lib.stream_ptr()), "dtb_laplacian_forward")
        ctx.save_for_backward(resid, edges, weight, rows)
        return loss

    @staticmethod
    def backward(ctx, g):
        resid, edges, weight, rows = ctx.saved_tensors
        B, V, _ = resid.shape
        grad = torch.zeros_like(resid)
        g = _f32c(g)
        with torch.cuda.device(resid.device):
            _lib.check(_lib.lib().dtb_laplacian_backward(_lib.ptr(resid), _lib.ptr(edges), _lib.ptr(weight), _lib.ptr(rows), _lib.ptr(g), B, V,
                                                         _lib.ptr(grad), _lib.stream_ptr()), "dtb_laplacian_backward")
        return grad, None, None


def laplacian_loss(offset_bxvx3, edges_ex2, weight_e):
    """``DefTet.laplacian_sparse`` (deftet.py:340-343) on the (edges, 1/deg) list of builders.tet_point_adj."""
    return _Laplacian.apply(offset_bxvx3, edges_ex2.to(torch.int32).contiguous(), _f32c(weight_e))
