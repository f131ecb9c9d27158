"""Voxel-feature sampling of the reference (layers/pv_module/functional/devoxelization.py:47-53 ``trilinear_devoxelize``,
layers/pc_model.py:182-194 ``sample_f``) on the sm_100a kernels of ``csrc/devox.cu``.

Same names, argument meaning and results as the reference: the live ``trilinear_devoxelize`` there is torch's ``F.grid_sample``
(bilinear, border padding, align_corners=False) behind a normalisation and an axis flip; gradients flow to the volume and to the
coordinates exactly as grid_sample's do.  ``sample_f`` is the fused form of the reference method: one launch per encoder level
writing straight into the concatenated (B, sum C, N) tensor, the ``+0.5 / *r / clamp`` prelude and its gradient done in-kernel."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

FROM_POSITIONS = 1
GLOBAL_GATHER = 2     # the no-staging kernels that serve R > 36
SIMPLE = 4            # one point per thread instead of four (self-tests, A/B timing)
NO_OWNER = 8          # volume gradient: the shared-atomic kernel also where the channel-owner kernel applies (R^3 <= 512)
NO_SORT = 16          # volume gradient: no sorted reduction (which is the default from 8 points per voxel on)
FORCE_SORT = 32       # volume gradient: sorted reduction whatever the point count


@_lib.register_signatures
def _devox_sigs(lib, sig):
    vp, i, ll = C.c_void_p, C.c_int, C.c_longlong
    sig("dtb_trilinear_devoxelize_forward", i, vp, vp, ll, ll, ll, i, i, i, i, i, vp, ll, vp)
    sig("dtb_trilinear_devoxelize_backward", i, vp, vp, ll, ll, ll, vp, ll, i, i, i, i, i, vp, vp, vp)
    sig("dtb_trilinear_devoxelize_backward_workspace", C.c_size_t, i, i, i, i, i)
    sig("dtb_trilinear_devoxelize_backward_ws", i, vp, vp, ll, ll, ll, vp, ll, i, i, i, i, i, vp, vp, vp, C.c_size_t, vp)


def _volume(c):
    if c.dim() != 5 or c.shape[2] != c.shape[3] or c.shape[3] != c.shape[4]:
        raise RuntimeError("features must be (B, C, R, R, R), got %s" % (tuple(c.shape),))
    return c.contiguous().float()


def _forward_level(feat, coords, strides, N, flags, out, out_offset):
    B, Cc, R = feat.shape[0], feat.shape[1], feat.shape[-1]
    dst = C.c_void_p(out.data_ptr() + 4 * out_offset)
    _lib.check(_lib.lib().dtb_trilinear_devoxelize_forward(_lib.ptr(feat), _lib.ptr(coords), strides[0], strides[1], strides[2], B, Cc, N, R,
                                                           flags, dst, out.stride(0), _lib.stream_ptr()), "dtb_trilinear_devoxelize_forward")


def _backward_level(feat, coords, strides, N, flags, grad_out, go_offset, grad_feat, grad_coords):
    B, Cc, R = feat.shape[0], feat.shape[1], feat.shape[-1]
    src = C.c_void_p(grad_out.data_ptr() + 4 * go_offset)
    lib = _lib.lib()
    # temporary memory for the sorted reduction of the volume gradient comes from torch's caching allocator (no cudaMalloc on the path)
    ws_bytes = lib.dtb_trilinear_devoxelize_backward_workspace(B, Cc, N, R, flags) if grad_feat is not None else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=feat.device) if ws_bytes else None
    _lib.check(lib.dtb_trilinear_devoxelize_backward_ws(_lib.ptr(feat), _lib.ptr(coords), strides[0], strides[1], strides[2], src,
                                                        grad_out.stride(0), B, Cc, N, R, flags, _lib.ptr(grad_feat), _lib.ptr(grad_coords),
                                                        _lib.ptr(ws), ws_bytes, _lib.stream_ptr()), "dtb_trilinear_devoxelize_backward_ws")


class _Devoxelize(torch.autograd.Function):
    """coords (B,3,N) voxel coordinates, any strides (the reference passes a permuted view)."""

    @staticmethod
    def forward(ctx, features, coords, flags):
        _lib.require_cuda(features, coords)
        feat = _volume(features)
        co = coords if coords.dtype == torch.float32 else coords.float()
        if not (co.is_contiguous() or co.permute(0, 2, 1).is_contiguous()):      # (B,3,N) or the reference's permuted (B,N,3) view
            co = co.contiguous()
        B, Cc, N = feat.shape[0], feat.shape[1], co.shape[2]
        if co.shape[0] != B or co.shape[1] != 3:
            raise RuntimeError("coords must be (B, 3, N), got %s" % (tuple(co.shape),))
        out = torch.empty(B, Cc, N, device=feat.device)
        with torch.cuda.device(feat.device):
            _forward_level(feat, co, co.stride(), N, flags, out, 0)
        ctx.save_for_backward(feat, co)
        ctx.flags = flags
        return out

    @staticmethod
    def backward(ctx, grad_out):
        feat, co = ctx.saved_tensors
        go = grad_out.contiguous().float()
        need_f, need_c = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grad_feat = torch.empty_like(feat) if need_f else None
        # the kernel addresses the coordinate gradient with the strides of coords
        gc_arg = torch.empty_strided(co.shape, co.stride(), device=co.device).zero_() if need_c else None
        with torch.cuda.device(feat.device):
            _backward_level(feat, co, co.stride(), co.shape[2], ctx.flags, go, 0, grad_feat, gc_arg)
        return grad_feat, gc_arg, None


def trilinear_devoxelize(c, coords, r, training=None):
    """devoxelization.py:47-53: c (B,C,R,R,R), coords (B,3,N) in voxel units -> (B,C,N).  ``r`` must be the volume's resolution (it
    always is at the reference's call sites, pc_model.py:190-192, pvconv.py:37); ``training`` is unused there as well."""
    if int(r) != c.shape[-1]:
        raise RuntimeError("resolution %s does not match the volume %s" % (r, tuple(c.shape)))
    return _Devoxelize.apply(c, coords, 0)


class _SampleF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, point_pos, flags, *c_list):
        _lib.require_cuda(point_pos, *c_list)
        flags = int(flags) | FROM_POSITIONS
        pos = point_pos.contiguous().float()
        feats = [_volume(c) for c in c_list]
        B, N = pos.shape[0], pos.shape[1]
        total = sum(f.shape[1] for f in feats)
        out = torch.empty(B, total, N, device=pos.device)
        strides = (N * 3, 1, 3)
        off = 0
        with torch.cuda.device(pos.device):
            for f in feats:
                if f.shape[0] != B:
                    raise RuntimeError("batch mismatch between positions and features")
                _forward_level(f, pos, strides, N, flags, out, off)
                off += f.shape[1] * N
        ctx.save_for_backward(pos, *feats)
        ctx.flags = flags
        return out

    @staticmethod
    def backward(ctx, grad_out):
        pos, *feats = ctx.saved_tensors
        go = grad_out.contiguous().float()
        B, N = pos.shape[0], pos.shape[1]
        need_pos = ctx.needs_input_grad[0]
        grad_pos = torch.zeros_like(pos) if need_pos else None
        grads = []
        off = 0
        with torch.cuda.device(pos.device):
            for i, f in enumerate(feats):
                need_f = ctx.needs_input_grad[2 + i]
                gf = torch.empty_like(f) if need_f else None
                if need_f or need_pos:
                    _backward_level(f, pos, (N * 3, 1, 3), N, ctx.flags, go, off, gf, grad_pos)
                grads.append(gf)
                off += f.shape[1] * N
        return (grad_pos, None, *grads)


def sample_f(point_pos, c_list, flags=0):
    """pc_model.py:182-194 (point-cloud branch): point_pos (B,N,3) in the unit cube centred at 0, c_list = the encoder's voxel
    features, one (B,C_i,R_i,R_i,R_i) per level -> (B, sum C_i, N)."""
    return _SampleF.apply(point_pos, flags, *c_list)
