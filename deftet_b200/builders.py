"""Topology builders (A10-A14) over csrc/builders.cu -- host-side mirror of ``utils/tet_utils.py`` /
``utils/lib/*/interface.py`` of the reference.

Device functions return torch tensors on the input's device (int32, reference ordering); the ``c_*`` /
class wrappers reproduce the reference's numpy/scipy/torch-sparse return types."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


@_lib.register_signatures
def _builder_sigs(lib, sig):
    vp, i, sz, ll = C.c_void_p, C.c_int, C.c_size_t, C.c_longlong
    sig("dtb_tet_point_adj_workspace", sz, i, i)
    sig("dtb_tet_point_adj", i, vp, i, i, vp, vp, vp, vp, sz, vp)
    sig("dtb_tet_to_face_workspace", sz, i)
    sig("dtb_tet_to_face", i, vp, i, i, vp, vp, vp, vp, vp, vp, sz, vp)
    sig("dtb_tet_adj_share_workspace", sz, i)
    sig("dtb_tet_adj_share", i, vp, i, i, vp, vp, vp, sz, vp)
    sig("dtb_tet_face_adj_workspace", sz, i)
    sig("dtb_tet_face_adj", i, vp, i, i, vp, ll, vp, vp, sz, vp)
    sig("dtb_collapse_vertices_workspace", sz, i)
    sig("dtb_collapse_vertices", i, vp, i, vp, vp, vp, vp, sz, vp)
    for name in ("dtb_host_tet_point_adj", "dtb_host_tet_adj_share", "dtb_host_tet_face_adj"):
        sig(name, i, vp, vp, vp, i, i)
    sig("dtb_host_colaps_v", i, vp, vp, vp, vp, i)


def _tet32(tet, device=None):
    if isinstance(tet, np.ndarray):
        tet = torch.from_numpy(np.ascontiguousarray(tet))
    if device is not None:
        tet = tet.to(device)
    _lib.require_cuda(tet)
    return _lib.aligned(tet.to(torch.int32))


def _ws(n, dev):
    return torch.empty(max(int(n), 16), device=dev, dtype=torch.uint8)


def tet_point_adj(tet, n_point, normalize=False):
    """-> edges (E,2) int32 sorted by (a,b) [, weights (E,) f32 = 1/deg(a)]"""
    tet = _tet32(tet)
    T, dev = tet.shape[0], tet.device
    L = _lib.lib()
    edges = torch.empty(T * 12, 2, device=dev, dtype=torch.int32)
    weight = torch.empty(T * 12, device=dev, dtype=torch.float32) if normalize else None
    n = torch.zeros(1, device=dev, dtype=torch.int32)
    wsz = L.dtb_tet_point_adj_workspace(n_point, T)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_tet_point_adj(_lib.ptr(tet), n_point, T, _lib.ptr(edges), _lib.ptr(weight), _lib.ptr(n), _lib.ptr(ws), wsz,
                                       _lib.stream_ptr()), "dtb_tet_point_adj")
    e = int(n.item())
    return (edges[:e], weight[:e]) if normalize else edges[:e]


def tet_to_adj_sparse(n_point, tet, normalize=True):
    """``c_tet_to_adj_sparse`` (utils/tet_utils.py:94-95): torch sparse (V,V) adjacency, row-normalised if asked."""
    if normalize:
        edges, w = tet_point_adj(tet, n_point, True)
    else:
        edges = tet_point_adj(tet, n_point, False)
        w = torch.ones(edges.shape[0], device=edges.device)
    return torch.sparse_coo_tensor(edges.t().long(), w, (n_point, n_point))


def tet_to_face(n_point, tet):
    """GPU ``tet_to_face`` (utils/tet_utils.py:208-256) -> (tet_face_fx3, tet_face_tetidx_fx2, tet_face_tetfaceidx_fx2,
    tet_boundary_face) as int32 device tensors in the reference's first-occurrence order."""
    tet = _tet32(tet)
    T, dev = tet.shape[0], tet.device
    L = _lib.lib()
    face = torch.empty(T * 2, 3, device=dev, dtype=torch.int32)
    ftet = torch.empty(T * 2, 2, device=dev, dtype=torch.int32)
    fslot = torch.empty(T * 2, 2, device=dev, dtype=torch.int32)
    bnd = torch.empty(T * 4, 3, device=dev, dtype=torch.int32)
    counts = torch.zeros(2, device=dev, dtype=torch.int32)
    wsz = L.dtb_tet_to_face_workspace(T)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_tet_to_face(_lib.ptr(tet), n_point, T, _lib.ptr(face), _lib.ptr(ftet), _lib.ptr(fslot), _lib.ptr(bnd),
                                     _lib.ptr(counts), _lib.ptr(ws), wsz, _lib.stream_ptr()), "dtb_tet_to_face")
    nf, nb = counts.tolist()
    return face[:nf], ftet[:nf], fslot[:nf], bnd[:nb]


def tet_adj_share(tet, n_point):
    """-> int32 (2*n_shared, 3) rows (t0,t1,f0),(t1,t0,f1) in the reference's key order."""
    tet = _tet32(tet)
    T, dev = tet.shape[0], tet.device
    L = _lib.lib()
    out = torch.empty(T * 8, 3, device=dev, dtype=torch.int32)
    n = torch.zeros(1, device=dev, dtype=torch.int32)
    wsz = L.dtb_tet_adj_share_workspace(T)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_tet_adj_share(_lib.ptr(tet), n_point, T, _lib.ptr(out), _lib.ptr(n), _lib.ptr(ws), wsz, _lib.stream_ptr()),
                   "dtb_tet_adj_share")
    return out[:2 * int(n.item())]


def tet_face_adj(tet, n_point, capacity_per_face=50):
    """-> int32 (n_pairs, 2) ordered pairs of tet-face ids sharing an edge (reference order)."""
    tet = _tet32(tet)
    T, dev = tet.shape[0], tet.device
    L = _lib.lib()
    cap = T * 4 * capacity_per_face
    pairs = torch.empty(cap, 2, device=dev, dtype=torch.int32)
    n = torch.zeros(1, device=dev, dtype=torch.int32)
    wsz = L.dtb_tet_face_adj_workspace(T)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_tet_face_adj(_lib.ptr(tet), n_point, T, _lib.ptr(pairs), cap, _lib.ptr(n), _lib.ptr(ws), wsz, _lib.stream_ptr()),
                   "dtb_tet_face_adj")
    k = int(n.item())
    if k > cap:
        raise _lib.DeftetB200Error("tet_face_adj: %d pairs exceed the capacity of %d rows" % (k, cap))
    return pairs[:k]


def collapse_vertices(points):
    """-> (map_array (N,) int32, inverse_idx (n_unique,) int32) -- colaps_v semantics."""
    _lib.require_cuda(points)
    pts = points.float().contiguous()
    N, dev = pts.shape[0], pts.device
    L = _lib.lib()
    m = torch.empty(N, device=dev, dtype=torch.int32)
    inv = torch.empty(N, device=dev, dtype=torch.int32)
    n = torch.zeros(1, device=dev, dtype=torch.int32)
    wsz = L.dtb_collapse_vertices_workspace(N)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_collapse_vertices(_lib.ptr(pts), N, _lib.ptr(m), _lib.ptr(inv), _lib.ptr(n), _lib.ptr(ws), wsz, _lib.stream_ptr()),
                   "dtb_collapse_vertices")
    return m, inv[:int(n.item())]


# ---- host-buffer ABI (numpy in / numpy out), the calling convention of utils/lib/*/interface.py ----------
def _np_p(a):
    return a.ctypes.data_as(C.c_void_p)


def host_run(name, tet_list, n_point, out_rows, out_cols):
    tet = np.ascontiguousarray(tet_list, dtype=np.int32)
    out = np.zeros((out_rows, out_cols), dtype=np.int32)
    n = np.zeros(1, dtype=np.int32)
    fn = getattr(_lib.lib(), "dtb_host_" + name)
    _lib.check(fn(_np_p(tet), _np_p(out), _np_p(n), int(n_point), tet.shape[0]), "dtb_host_" + name)
    return out, int(n[0])


def host_colaps_v(point_nx3):
    pts = np.ascontiguousarray(point_nx3, dtype=np.float32)
    n_point = pts.shape[0]
    m = np.zeros(n_point, dtype=np.int32)
    inv = np.zeros(n_point, dtype=np.int32)
    cnt = np.zeros(1, dtype=np.int32)
    _lib.check(_lib.lib().dtb_host_colaps_v(_np_p(pts), _np_p(m), _np_p(inv), _np_p(cnt), n_point), "dtb_host_colaps_v")
    return m, inv[:cnt[0]]
