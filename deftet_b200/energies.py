"""Per-tet energies as autograd Functions over the fused sm_100a kernels (csrc/energies.cu).

Host-side mirror of ``DefTet.amips_energy / volume_variance / edge_length / tet_inverse_v``
(reference ``layers/DefTet/deftet.py:239-338``).  Two call forms:

* ``tet_energies(pos, tet, inv_v)`` -- the engine form: vertex positions (B,V,3) + shared int32 topology
  (T,4); never materialises the (B,T,4,3) gather, gradient is scattered straight into (B,V,3).
* ``amips_energy_soup / volume_variance_soup / edge_length_soup(tet_bxfx4x3, ...)`` -- the drop-in form
  with exactly the tensors the reference methods receive.
"""
from __future__ import annotations

import torch

from . import _lib

AMIPS, EDGE, VOLUME, ALL = 1, 2, 4, 7


def _f32c(t):
    return _lib.aligned(t if t.dtype == torch.float32 else t.float())


def tet_inverse_v(init_pos: torch.Tensor, tet: torch.Tensor) -> torch.Tensor:
    """(T,3,3) inverse of the 20x-scaled rest offset matrix; singular -> identity (deftet.py:205-233,300-318)."""
    _lib.require_cuda(init_pos, tet)
    pos = _f32c(init_pos)
    tet32 = _lib.aligned(tet.to(torch.int32))
    T = tet32.shape[0]
    out = torch.empty(T, 3, 3, device=pos.device, dtype=torch.float32)
    with torch.cuda.device(pos.device):
        _lib.check(_lib.lib().dtb_tet_inverse_v(_lib.ptr(pos), _lib.ptr(tet32), pos.shape[0], T, _lib.ptr(out),
                                               _lib.stream_ptr()), "dtb_tet_inverse_v")
    return out


class TetTiles:
    """Tile-local re-encoding of a tet list (csrc/energies_tiled.cu): tiles of 256 consecutive tets with their distinct
    vertices, 2-byte local corner ids and per-vertex incidence lists.  Built once per topology on the GPU
    (``dtb_tet_tiles_build``); the energy kernels stage vertices per tile instead of gathering 12 scalars per tet."""

    def __init__(self, tet32: torch.Tensor, n_vert: int):
        _lib.require_cuda(tet32)
        assert tet32.dtype == torch.int32 and tet32.is_contiguous()
        L = _lib.lib()
        T = tet32.shape[0]
        dev = tet32.device
        nbytes = L.dtb_tet_tiles_bytes(T)
        self.buf = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        nmax = torch.zeros(1, device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            _lib.check(L.dtb_tet_tiles_build(_lib.ptr(tet32), T, int(n_vert), _lib.ptr(self.buf), nbytes, _lib.ptr(nmax),
                                             _lib.stream_ptr()), "dtb_tet_tiles_build")
        self.nloc_max = int(nmax.item())            # set-up time synchronisation (once per topology)
        self.n_tet = T
        self.tet = tet32


_TILE_CACHE = {}


def tiles_for(tet32: torch.Tensor, n_vert: int) -> TetTiles:
    """TetTiles of a topology tensor, cached on (storage address, shape, version) -- the last 8 topologies are kept."""
    key = (tet32.data_ptr(), tet32.shape[0], tet32._version, tet32.device.index)
    t = _TILE_CACHE.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.DeftetB200Error("tet_energies: first use of a topology inside CUDA-graph capture; call "
                                       "deftet_b200.energies.tiles_for(tet, n_vert) once before capturing")
        if len(_TILE_CACHE) >= 8:
            _TILE_CACHE.pop(next(iter(_TILE_CACHE)))
        t = _TILE_CACHE[key] = TetTiles(tet32, n_vert)
    return t


def _use_tiled():
    """The tile-local kernels (csrc/energies_tiled.cu) are an OPT-IN alternative: measured on B200 at res 70 b8 they execute
    20-40 % more instructions than the direct-gather kernels for the same issue rate (profiles/r2_ab_energies.md) -- the
    kernels are issue-bound, not load/RED-bound -- so the direct kernels stay the default."""
    import os
    return os.environ.get("DTB_ENERGY_PATH", "") == "tiled"


class _TetEnergiesTiled(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, tet32, inv_v, flags, tiles):
        _lib.require_cuda(pos, tet32)
        pos = _f32c(pos)
        B, V, _ = pos.shape
        T = tet32.shape[0]
        dev = pos.device
        L = _lib.lib()
        out = torch.zeros(3, B, device=dev, dtype=torch.float32)
        stats = torch.empty(B, 8, device=dev, dtype=torch.float64)
        with torch.cuda.device(dev):
            _lib.check(L.dtb_tet_energies_forward_tiled(_lib.ptr(pos), _lib.ptr(tet32), _lib.ptr(inv_v), _lib.ptr(tiles.buf), tiles.nloc_max,
                                                        B, V, T, flags, _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]),
                                                        _lib.ptr(stats), _lib.stream_ptr()), "dtb_tet_energies_forward_tiled")
        ctx.save_for_backward(pos, inv_v, stats)
        ctx.flags, ctx.tiles, ctx.T = flags, tiles, T
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_amips, g_edge, g_vol):
        pos, inv_v, stats = ctx.saved_tensors
        B, V, _ = pos.shape
        tiles = ctx.tiles
        grad4 = torch.zeros(B, V, 4, device=pos.device, dtype=torch.float32)
        gs = [None if g is None else _f32c(g) for g in (g_amips, g_edge, g_vol)]
        with torch.cuda.device(pos.device):
            _lib.check(_lib.lib().dtb_tet_energies_backward_tiled(_lib.ptr(pos), _lib.ptr(inv_v), _lib.ptr(tiles.buf), tiles.nloc_max, B, V,
                                                                  ctx.T, ctx.flags, _lib.ptr(stats), _lib.ptr(gs[0]), _lib.ptr(gs[1]),
                                                                  _lib.ptr(gs[2]), _lib.ptr(grad4), _lib.stream_ptr()),
                       "dtb_tet_energies_backward_tiled")
        return grad4[..., :3], None, None, None, None


class _TetEnergies(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, tet32, inv_v, flags):
        _lib.require_cuda(pos, tet32)
        pos = _f32c(pos)
        B, V, _ = pos.shape
        T = tet32.shape[0]
        dev = pos.device
        L = _lib.lib()
        out = torch.zeros(3, B, device=dev, dtype=torch.float32)
        stats = torch.empty(B, 8, device=dev, dtype=torch.float64)
        wsz = L.dtb_tet_energies_workspace(B, V, T)
        ws = torch.empty(wsz, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _lib.check(L.dtb_tet_energies_forward(_lib.ptr(pos), _lib.ptr(tet32), _lib.ptr(inv_v), B, V, T, flags,
                                                  _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.ptr(stats),
                                                  _lib.ptr(ws), wsz, _lib.stream_ptr()), "dtb_tet_energies_forward")
        ctx.save_for_backward(pos, tet32, inv_v, stats)
        ctx.flags = flags
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_amips, g_edge, g_vol):
        pos, tet32, inv_v, stats = ctx.saved_tensors
        B, V, _ = pos.shape
        T = tet32.shape[0]
        # padded (B,V,4) accumulator: one 16-byte vector reduction per vertex update; the xyz view is returned as the gradient
        grad4 = torch.zeros(B, V, 4, device=pos.device, dtype=torch.float32)
        gs = [None if g is None else _f32c(g) for g in (g_amips, g_edge, g_vol)]
        with torch.cuda.device(pos.device):
            _lib.check(_lib.lib().dtb_tet_energies_backward_v4(_lib.ptr(pos), _lib.ptr(tet32), _lib.ptr(inv_v), B, V, T,
                                                               ctx.flags, _lib.ptr(stats), _lib.ptr(gs[0]), _lib.ptr(gs[1]),
                                                               _lib.ptr(gs[2]), _lib.ptr(grad4), _lib.stream_ptr()),
                       "dtb_tet_energies_backward_v4")
        return grad4[..., :3], None, None, None


def tet_energies(pos, tet32, inv_v, flags=ALL, tiles=None):
    """-> (amips (B,), edge (B,), volume_variance (B,)) for vertex positions (B,V,3).

    ``tiles``: the TetTiles of ``tet32`` for the opt-in tile-local kernels (DTB_ENERGY_PATH=tiled); looked up in / added to a
    small cache when omitted."""
    if tet32.dtype != torch.int32:
        tet32 = tet32.to(torch.int32)
    _lib.require_cuda(pos, tet32)
    tet32 = _lib.aligned(tet32)
    inv = None if inv_v is None else _f32c(inv_v)
    if not _use_tiled():
        return _TetEnergies.apply(pos, tet32, inv, int(flags))
    if tiles is None:
        tiles = tiles_for(tet32, pos.shape[1])
    return _TetEnergiesTiled.apply(pos, tet32, inv, int(flags), tiles)


def tet_energies_direct(pos, tet32, inv_v, flags=ALL):
    """The direct-gather kernels (csrc/energies.cu): no per-topology set-up, for topologies seen only once."""
    if tet32.dtype != torch.int32:
        tet32 = tet32.to(torch.int32)
    _lib.require_cuda(pos, tet32)
    return _TetEnergies.apply(pos, _lib.aligned(tet32), None if inv_v is None else _f32c(inv_v), int(flags))


class _SoupEnergies(torch.autograd.Function):
    @staticmethod
    def forward(ctx, soup, inv_v, flags):
        _lib.require_cuda(soup)
        soup = _f32c(soup)
        B, T = soup.shape[0], soup.shape[1]
        dev = soup.device
        L = _lib.lib()
        out = torch.zeros(3, B, device=dev, dtype=torch.float32)
        stats = torch.empty(B, 8, device=dev, dtype=torch.float64)
        wsz = L.dtb_tet_energies_workspace(B, 0, T)
        ws = torch.empty(wsz, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _lib.check(L.dtb_tet_energies_forward_soup(_lib.ptr(soup), _lib.ptr(inv_v), B, T, flags, _lib.ptr(out[0]),
                                                       _lib.ptr(out[1]), _lib.ptr(out[2]), _lib.ptr(stats), _lib.ptr(ws),
                                                       wsz, _lib.stream_ptr()), "dtb_tet_energies_forward_soup")
        ctx.save_for_backward(soup, inv_v, stats)
        ctx.flags = flags
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_amips, g_edge, g_vol):
        soup, inv_v, stats = ctx.saved_tensors
        B, T = soup.shape[0], soup.shape[1]
        grad = torch.empty_like(soup)
        gs = [None if g is None else _f32c(g) for g in (g_amips, g_edge, g_vol)]
        with torch.cuda.device(soup.device):
            _lib.check(_lib.lib().dtb_tet_energies_backward_soup(_lib.ptr(soup), _lib.ptr(inv_v), B, T, ctx.flags,
                                                                 _lib.ptr(stats), _lib.ptr(gs[0]), _lib.ptr(gs[1]),
                                                                 _lib.ptr(gs[2]), _lib.ptr(grad), _lib.stream_ptr()),
                       "dtb_tet_energies_backward_soup")
        return grad, None, None


def amips_energy_soup(tet_bxfx4x3, inverse_v):
    return _SoupEnergies.apply(tet_bxfx4x3, _f32c(inverse_v), AMIPS)[0]


def edge_length_soup(tet_bxfx4x3):
    return _SoupEnergies.apply(tet_bxfx4x3, None, EDGE)[1]


def volume_variance_soup(tet_bxfx4x3):
    return _SoupEnergies.apply(tet_bxfx4x3, None, VOLUME)[2]
