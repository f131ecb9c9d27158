"""Topology editing and regularisers of the diff_render optimisation loop on the GPU.

Function names and argument meaning follow the reference's host code under diff_render/diftet_6_subdiv/3_model
(prepare_for_wz.py, utils_tetsv.py, deftet.py), which runs in numpy / Python loops; here every step is a CUDA kernel of
``csrc/topology.cu`` / ``csrc/builders.cu`` working on device tensors.  Index-valued results are int32 device tensors in the
reference's order; callers that need the reference's numpy / int64 types convert (``deftet_b200.dropin.diff_render``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .builders import _tet32, _ws, tet_point_adj
from .search import _f32c


@_lib.register_signatures
def _topology_sigs(lib, sig):
    vp, i, sz, f = C.c_void_p, C.c_int, C.c_size_t, C.c_float
    sig("dtb_tet_edges_workspace", sz, i)
    sig("dtb_tet_edges", i, vp, i, i, vp, vp, vp, vp, sz, vp)
    sig("dtb_subdivide_tets_workspace", sz, i)
    sig("dtb_subdivide_tets", i, vp, vp, vp, i, i, vp, vp, vp, sz, vp)
    sig("dtb_edge_midpoints", i, vp, i, vp, i, vp, vp)
    sig("dtb_tet_to_face_idx_workspace", sz, i)
    sig("dtb_tet_to_face_idx", i, vp, i, i, vp, vp, vp, vp, vp, sz, vp)
    sig("dtb_tet_neighbours_workspace", sz, i)
    sig("dtb_tet_neighbours", i, vp, i, i, vp, vp, sz, vp)
    sig("dtb_point_adj_rows", i, vp, i, i, vp, vp, vp, vp, vp)
    sig("dtb_point_adj_table", i, vp, i, i, vp, i, vp, vp)
    sig("dtb_tet_delete_workspace", sz, i)
    sig("dtb_tet_delete", i, vp, vp, vp, i, i, f, vp, vp, vp, vp, sz, vp)
    sig("dtb_featlap_forward", i, vp, vp, vp, i, i, i, vp, vp)
    sig("dtb_featlap_backward", i, vp, vp, vp, vp, i, i, i, vp, vp)
    sig("dtb_tet_volume_deviation_forward", i, vp, vp, i, f, vp, vp, vp)
    sig("dtb_tet_volume_deviation_backward", i, vp, vp, i, f, vp, vp, vp, vp)
    sig("dtb_project_faces_forward", i, vp, vp, vp, vp, vp, vp, i, i, i, f, i, vp, vp, vp, vp)
    sig("dtb_project_faces_backward", i, vp, vp, vp, vp, vp, vp, i, i, i, f, i, vp, vp, vp, vp, vp, vp)


# ---------------------------------------------------------------------------------------------------------------- T1 / T2
def tet_edges(tet, n_point):
    """``generate_edge`` + ``generate_tet_edge_idx`` (prepare_for_wz.py:186-238) -> (edges (E,2) i32 sorted, tet_edge (T,6) i32)."""
    tet = _tet32(tet)
    T, dev = tet.shape[0], tet.device
    L = _lib.lib()
    edges = torch.empty(T * 6, 2, device=dev, dtype=torch.int32)
    tet_edge = torch.empty(T, 6, device=dev, dtype=torch.int32)
    n = torch.zeros(1, device=dev, dtype=torch.int32)
    wsz = L.dtb_tet_edges_workspace(T)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_tet_edges(_lib.ptr(tet), int(n_point), T, _lib.ptr(edges), _lib.ptr(tet_edge), _lib.ptr(n), _lib.ptr(ws), wsz,
                                   _lib.stream_ptr()), "dtb_tet_edges")
    return edges[:int(n.item())], tet_edge


def edge_midpoints(values_pxk, edges_ex2):
    """``generate_edge_points`` (prepare_for_wz.py:241-254) for one (P,K) array -> (E,K)."""
    _lib.require_cuda(values_pxk, edges_ex2)
    x = _f32c(values_pxk)
    e = edges_ex2.to(torch.int32).contiguous()
    E, K = e.shape[0], x.shape[1]
    out = torch.empty(E, K, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().dtb_edge_midpoints(_lib.ptr(x), K, _lib.ptr(e), E, _lib.ptr(out), _lib.stream_ptr()), "dtb_edge_midpoints")
    return out


def generate_subdivision(tet_list_tx4, tet_points_px3, tet_feat_pxk, tet_list_subdiv_sig=None):
    """``generate_subdivision`` (prepare_for_wz.py:257-301) on device tensors ->
    (points (P+E,3) f32, feat (P+E,K) f32, tets (T',4) i32); ``tet_list_subdiv_sig`` is a bool (T,) tensor or None (= all)."""
    _lib.require_cuda(tet_list_tx4, tet_points_px3, tet_feat_pxk)
    tet = _tet32(tet_list_tx4)
    pts, feat = _f32c(tet_points_px3), _f32c(tet_feat_pxk)
    T, P, dev = tet.shape[0], pts.shape[0], tet.device
    edges, tet_edge = tet_edges(tet, P)
    new_pts = torch.cat([pts, edge_midpoints(pts, edges)], dim=0)
    new_feat = torch.cat([feat, edge_midpoints(feat, edges)], dim=0)
    L = _lib.lib()
    sig = None if tet_list_subdiv_sig is None else tet_list_subdiv_sig.to(device=dev, dtype=torch.uint8).contiguous()
    out = torch.empty(T * 8, 4, device=dev, dtype=torch.int32)
    n = torch.zeros(1, device=dev, dtype=torch.int32)
    wsz = L.dtb_subdivide_tets_workspace(T)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_subdivide_tets(_lib.ptr(tet), _lib.ptr(tet_edge), _lib.ptr(sig), P, T, _lib.ptr(out), _lib.ptr(n), _lib.ptr(ws), wsz,
                                        _lib.stream_ptr()), "dtb_subdivide_tets")
    return new_pts, new_feat, out[:int(n.item())]


# ---------------------------------------------------------------------------------------------------------------- geometry tables
def tet_to_face_idx(n_point, tet_list, with_boundary=True):
    """``tet_to_face_idx`` (prepare_for_wz.py:49-108) -> (faces (F,3), face_tet (F,2), face_slot (F,2)) i32, first-occurrence order;
    boundary faces (second column -1) are included iff ``with_boundary``."""
    tet = _tet32(tet_list)
    T, dev = tet.shape[0], tet.device
    L = _lib.lib()
    face = torch.empty(T * 4, 3, device=dev, dtype=torch.int32)
    ftet = torch.empty(T * 4, 2, device=dev, dtype=torch.int32)
    fslot = torch.empty(T * 4, 2, device=dev, dtype=torch.int32)
    n = torch.zeros(1, device=dev, dtype=torch.int32)
    wsz = L.dtb_tet_to_face_idx_workspace(T)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_tet_to_face_idx(_lib.ptr(tet), int(n_point), T, _lib.ptr(face), _lib.ptr(ftet), _lib.ptr(fslot), _lib.ptr(n),
                                         _lib.ptr(ws), wsz, _lib.stream_ptr()), "dtb_tet_to_face_idx")
    k = int(n.item())
    face, ftet, fslot = face[:k], ftet[:k], fslot[:k]
    if not with_boundary:
        keep = ftet[:, 1] >= 0
        face, ftet, fslot = face[keep], ftet[keep], fslot[keep]
    return face, ftet, fslot


def tet_neighbours(tet_list, n_point):
    """``tet_neighbour_idx`` of utils_tetsv.tet_adj_share (utils_tetsv.py:16-62): (T,4) i32, -1 = no neighbour across that face."""
    tet = _tet32(tet_list)
    T, dev = tet.shape[0], tet.device
    L = _lib.lib()
    nbr = torch.empty(T, 4, device=dev, dtype=torch.int32)
    wsz = L.dtb_tet_neighbours_workspace(T)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_tet_neighbours(_lib.ptr(tet), int(n_point), T, _lib.ptr(nbr), _lib.ptr(ws), wsz, _lib.stream_ptr()),
                   "dtb_tet_neighbours")
    return nbr


def generate_point_adj_idx(n_point, tet_list):
    """``generate_point_adj_idx`` (prepare_for_wz.py:121-137) -> (pointadj_idx (P,M) i32 with -1 padding, adjsum (P,1) f32) without
    the reference's dense (P,P) matrix."""
    tet = _tet32(tet_list)
    dev = tet.device
    P = int(n_point)
    edges = tet_point_adj(tet, P).contiguous()
    E = edges.shape[0]
    L = _lib.lib()
    row_start = torch.empty(P, device=dev, dtype=torch.int32)
    row_end = torch.empty(P, device=dev, dtype=torch.int32)
    degree = torch.empty(P, device=dev)
    mx = torch.zeros(1, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_point_adj_rows(_lib.ptr(edges), E, P, _lib.ptr(row_start), _lib.ptr(row_end), _lib.ptr(degree), _lib.ptr(mx),
                                        _lib.stream_ptr()), "dtb_point_adj_rows")
        M = int(mx.item())
        table = torch.empty(P, M, device=dev, dtype=torch.int32)
        _lib.check(L.dtb_point_adj_table(_lib.ptr(edges), E, P, _lib.ptr(row_start), M, _lib.ptr(table), _lib.stream_ptr()),
                   "dtb_point_adj_table")
    return table, degree.reshape(-1, 1)


# ---------------------------------------------------------------------------------------------------------------- T4
def delete_tet_by_weight(tet_list, point_weights_px1, tet_neighbour_idx, thres=0.01, neilevel=3):
    """``Deftet.deletetet`` minus the bookkeeping (3_model/deftet.py:311-326; prepare_for_wz.py:171-181): keep a tet iff the largest
    vertex weight within ``neilevel`` neighbour steps exceeds ``thres`` -> (kept tets (K,4) i32, keep (T,) bool).  If nothing
    would be kept the input list is returned unchanged, like the reference."""
    _lib.require_cuda(point_weights_px1)
    tet = _tet32(tet_list)
    T, dev = tet.shape[0], tet.device
    w = _f32c(point_weights_px1).reshape(-1)
    nbr = tet_neighbour_idx.to(device=dev, dtype=torch.int32).contiguous()
    L = _lib.lib()
    out = torch.empty(T, 4, device=dev, dtype=torch.int32)
    keep = torch.empty(T, device=dev, dtype=torch.uint8)
    n = torch.zeros(1, device=dev, dtype=torch.int32)
    wsz = L.dtb_tet_delete_workspace(T)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_tet_delete(_lib.ptr(tet), _lib.ptr(w), _lib.ptr(nbr), T, int(neilevel), float(thres), _lib.ptr(out), _lib.ptr(n),
                                    _lib.ptr(keep), _lib.ptr(ws), wsz, _lib.stream_ptr()), "dtb_tet_delete")
    k = int(n.item())
    if k == 0:
        return tet, torch.ones(T, device=dev, dtype=torch.bool)
    return out[:k], keep.bool()


# ---------------------------------------------------------------------------------------------------------------- T5 / T6
class _FeatLap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, weight):
        x = _f32c(x)
        P, C_ = x.shape
        M = table.shape[1]
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().dtb_featlap_forward(_lib.ptr(x), _lib.ptr(table), _lib.ptr(weight), P, M, C_, _lib.ptr(out),
                                                      _lib.stream_ptr()), "dtb_featlap_forward")
        ctx.save_for_backward(x, table, weight)
        return out

    @staticmethod
    def backward(ctx, g):
        x, table, weight = ctx.saved_tensors
        P, C_ = x.shape
        g = _f32c(g)
        gx = torch.zeros_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().dtb_featlap_backward(_lib.ptr(x), _lib.ptr(table), _lib.ptr(weight), _lib.ptr(g), P, table.shape[1], C_,
                                                       _lib.ptr(gx), _lib.stream_ptr()), "dtb_featlap_backward")
        return gx, None, None


def featlap(pointfeat_pxc, point_adj_idx_pxm, point_adj_weights_px1):
    """``Deftet.get_featlap`` (3_model/deftet.py:227-250): per-element squared difference between each vertex feature and the
    degree-normalised sum over its neighbours.  ``point_adj_idx_pxm`` holds neighbour ids with -1 padding (NOT the reference's
    +1-shifted copy), ``point_adj_weights_px1`` = degree + 1e-10."""
    _lib.require_cuda(pointfeat_pxc, point_adj_idx_pxm, point_adj_weights_px1)
    table = point_adj_idx_pxm.to(torch.int32).contiguous()
    return _FeatLap.apply(pointfeat_pxc, table, _f32c(point_adj_weights_px1).reshape(-1))


class _VolumeDeviation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, tet, scale):
        pos = _f32c(pos)
        T = tet.shape[0]
        out = torch.empty(T, device=pos.device)
        acc = torch.empty(1, device=pos.device, dtype=torch.float64)
        with torch.cuda.device(pos.device):
            _lib.check(_lib.lib().dtb_tet_volume_deviation_forward(_lib.ptr(pos), _lib.ptr(tet), T, scale, _lib.ptr(out), _lib.ptr(acc),
                                                                   _lib.stream_ptr()), "dtb_tet_volume_deviation_forward")
        ctx.save_for_backward(pos, tet)
        ctx.scale = scale
        return out

    @staticmethod
    def backward(ctx, g):
        pos, tet = ctx.saved_tensors
        g = _f32c(g)
        gp = torch.zeros_like(pos)
        acc = torch.empty(1, device=pos.device, dtype=torch.float64)
        with torch.cuda.device(pos.device):
            _lib.check(_lib.lib().dtb_tet_volume_deviation_backward(_lib.ptr(pos), _lib.ptr(tet), tet.shape[0], ctx.scale, _lib.ptr(g),
                                                                    _lib.ptr(acc), _lib.ptr(gp), _lib.stream_ptr()),
                       "dtb_tet_volume_deviation_backward")
        return gp, None, None


def volume_deviation(points_px3, tet_tx4, scale=2.0):
    """``Deftet.get_volume_variance`` (3_model/deftet.py:252-309): (T,) signed tet volume of ``scale * points`` minus its mean."""
    _lib.require_cuda(points_px3, tet_tx4)
    return _VolumeDeviation.apply(points_px3, _tet32(tet_tx4), float(scale))


# ---------------------------------------------------------------------------------------------------------------- T7
class _ProjectFaces(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, feat, faces, rot, cam_pos, proj, multiplier, sigmoid):
        pos, feat = _f32c(pos), _f32c(feat)
        rot, cam_pos, proj = _f32c(rot), _f32c(cam_pos), _f32c(proj).reshape(-1)
        B, F, D = rot.shape[0], faces.shape[0], feat.shape[1]
        dev = pos.device
        fz = torch.empty(B, F, 3, device=dev)
        fxy = torch.empty(B, F, 3, 2, device=dev)
        ff = torch.empty(B, F, 3, D, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().dtb_project_faces_forward(_lib.ptr(pos), _lib.ptr(feat), _lib.ptr(faces), _lib.ptr(rot), _lib.ptr(cam_pos),
                                                            _lib.ptr(proj), B, F, D, multiplier, int(sigmoid), _lib.ptr(fz), _lib.ptr(fxy),
                                                            _lib.ptr(ff), _lib.stream_ptr()), "dtb_project_faces_forward")
        ctx.save_for_backward(pos, feat, faces, rot, cam_pos, proj)
        ctx.cfg = (multiplier, int(sigmoid))
        return fz, fxy, ff

    @staticmethod
    def backward(ctx, g_z, g_xy, g_ff):
        pos, feat, faces, rot, cam_pos, proj = ctx.saved_tensors
        multiplier, sigmoid = ctx.cfg
        B, F, D = rot.shape[0], faces.shape[0], feat.shape[1]
        g_z = _f32c(g_z) if g_z is not None else None
        g_xy = _f32c(g_xy) if g_xy is not None else None
        g_ff = _f32c(g_ff) if g_ff is not None else None
        gp = torch.zeros_like(pos) if ctx.needs_input_grad[0] else None
        gf = torch.zeros_like(feat) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(pos.device):
            _lib.check(_lib.lib().dtb_project_faces_backward(_lib.ptr(pos), _lib.ptr(feat), _lib.ptr(faces), _lib.ptr(rot), _lib.ptr(cam_pos),
                                                             _lib.ptr(proj), B, F, D, multiplier, sigmoid, _lib.ptr(g_xy), _lib.ptr(g_ff),
                                                             _lib.ptr(g_z), _lib.ptr(gp), _lib.ptr(gf), _lib.stream_ptr()),
                       "dtb_project_faces_backward")
        return gp, gf, None, None, None, None, None, None


def project_faces(points_px3, pointfeat_pxd, faces_fx3, camrot_bx3x3, campos_bx3, camproj_3x1, multiplier=1.0, sigmoid=True):
    """Fused ``perspective`` (3_model/cameraop.py:14-33) + ``vertex2face`` (4_render/vertex2face.py:14-28) + the feature sigmoid of
    ``rendermeshcolor`` (5_rendereq/deftetrneder.py:84) for B views of ONE vertex set ->
    (face_vertices_z (B,F,3), face_vertices_image (B,F,3,2) * multiplier, face_features (B,F,3,D)), differentiable w.r.t. the
    vertex positions and features."""
    _lib.require_cuda(points_px3, pointfeat_pxd, faces_fx3, camrot_bx3x3)
    faces = faces_fx3.to(torch.int32).contiguous()
    return _ProjectFaces.apply(points_px3, pointfeat_pxd, faces, camrot_bx3x3, campos_bx3, camproj_3x1, float(multiplier), bool(sigmoid))
