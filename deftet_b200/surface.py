"""Predicted-surface stage: boundary faces (A9), surface sampling + one-sided chamfer (A3/A2), point->surface
distance (A4), boundary-face adjacency + normal consistency (A5) -- over csrc/surface.cu, tridist.cu, faceadj.cu.

Two layers:
* drop-in functions with the reference's tensor contracts (``tet_analytic_distance_f_batch``,
  ``tet_face_adj_m_f_idx``, ``get_boundary_index``);
* the batched engine ops on the padded-ragged layout (faces (B,Fmax,3) int32 + counts (B,) int32) that remove
  the reference's per-sample Python loop (layers/DefTet/deftet.py:89-103).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .search import _f32c


@_lib.register_signatures
def _surface_sigs(lib, sig):
    vp, i, sz, f = C.c_void_p, C.c_int, C.c_size_t, C.c_float
    sig("dtb_nearest_neighbor_ragged", i, vp, vp, i, vp, vp, i, i, i, i, vp, sz, vp)
    sig("dtb_boundary_faces_workspace", sz, i, i)
    sig("dtb_boundary_faces", i, vp, vp, vp, i, i, i, i, vp, vp, vp, vp, sz, vp)
    sig("dtb_surface_sample", i, vp, vp, vp, vp, vp, i, i, i, i, vp, vp)
    sig("dtb_chamfer_forward", i, vp, vp, vp, vp, i, i, i, i, vp, vp, vp)
    sig("dtb_chamfer_backward", i, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp, i, vp)
    sig("dtb_face_soup", i, vp, vp, vp, i, i, i, vp, vp)
    sig("dtb_point_face_distance_grid_res", i, i)
    sig("dtb_point_face_distance_workspace", sz, i, i, i, i)
    sig("dtb_point_face_distance_forward", i, vp, vp, vp, i, i, i, i, vp, vp, vp, sz, vp)
    sig("dtb_point_face_distance_backward", i, vp, vp, vp, vp, i, i, i, vp, vp)
    sig("dtb_point_face_distance_backward_indexed", i, vp, vp, vp, vp, vp, vp, i, i, i, i, vp, i, vp)
    sig("dtb_sqrt_mean", i, vp, vp, i, i, f, vp, vp, vp)
    sig("dtb_face_adjacency_workspace", sz, i, i, i)
    sig("dtb_face_adjacency", i, vp, vp, vp, i, i, i, vp, vp, vp, vp, sz, vp)
    sig("dtb_normal_loss_forward", i, vp, vp, vp, vp, i, i, i, vp, vp, vp, vp)
    sig("dtb_normal_loss_backward", i, vp, vp, vp, vp, vp, vp, vp, i, i, i, vp, vp, i, vp)


def _i32c(t):
    return _lib.aligned(t if t.dtype == torch.int32 else t.to(torch.int32))


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 16), device=dev, dtype=torch.uint8)


# ====================================================================================================
# drop-in layer
# ====================================================================================================
class _AnalyticDistance(torch.autograd.Function):
    """``VarianceFunc`` of layers/DefTet/tet_analytic_distance_batch/utils.py:35-79."""

    @staticmethod
    def forward(ctx, gt_point_clouds_bxpx3, face_bxfx3x3, n_face_b):
        _lib.require_cuda(gt_point_clouds_bxpx3, face_bxfx3x3)
        pts = _f32c(gt_point_clouds_bxpx3)
        faces = _f32c(face_bxfx3x3)
        B, S = pts.shape[0], pts.shape[1]
        F = faces.shape[1]
        dev = pts.device
        counts = n_face_b.to(device=dev).to(torch.int32).clamp(max=F).contiguous()      # static_cast<int>(n_face_b[b])
        closest_f = torch.zeros(B, S, 1, device=dev)
        closest_d = torch.zeros(B, S, 1, device=dev)
        L = _lib.lib()
        wsz = L.dtb_point_face_distance_workspace(B, S, F, 0)
        ws = _ws(wsz, dev)
        with torch.cuda.device(dev):
            _lib.check(L.dtb_point_face_distance_forward(_lib.ptr(pts), _lib.ptr(faces), _lib.ptr(counts), B, S, F, 0,
                                                         _lib.ptr(closest_d), _lib.ptr(closest_f), _lib.ptr(ws), wsz,
                                                         _lib.stream_ptr()), "dtb_point_face_distance_forward")
        ctx.save_for_backward(pts, faces, closest_f)
        ctx.mark_non_differentiable(closest_f)
        return closest_d, closest_f

    @staticmethod
    def backward(ctx, dl_dclosest_d, dl_dclosest_f):
        pts, faces, closest_f = ctx.saved_tensors
        B, S, F = pts.shape[0], pts.shape[1], faces.shape[1]
        g = _f32c(dl_dclosest_d)
        dldface = torch.zeros(B, F, 3, 3, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib().dtb_point_face_distance_backward(_lib.ptr(pts), _lib.ptr(faces), _lib.ptr(closest_f), _lib.ptr(g),
                                                                   B, S, F, _lib.ptr(dldface), _lib.stream_ptr()),
                       "dtb_point_face_distance_backward")
        return None, dldface, None


tet_analytic_distance_f_batch = _AnalyticDistance.apply


def face_adjacency_table(face_fx3x3=None, faces=None, counts=None, n_vert=0, want="f32"):
    """(B?,F,30) table of edge-neighbours (ascending, -1 padded).  Give a coordinate soup or vertex ids."""
    L = _lib.lib()
    if face_fx3x3 is not None:
        soup = _f32c(face_fx3x3)
        batched = soup.dim() == 4
        B = soup.shape[0] if batched else 1
        F = soup.shape[-3]
        dev = soup.device
        fidx = None
    else:
        fidx = _i32c(faces)
        batched = fidx.dim() == 3
        B = fidx.shape[0] if batched else 1
        F = fidx.shape[-2]
        dev = fidx.device
        soup = None
    _lib.require_cuda(soup, fidx)
    adj_f = torch.empty(B, F, 30, device=dev, dtype=torch.float32) if want == "f32" else None
    adj_i = torch.empty(B, F, 30, device=dev, dtype=torch.int32) if want == "i32" else None
    if F > 0:
        wsz = L.dtb_face_adjacency_workspace(B, F, 0 if soup is not None else n_vert)
        ws = _ws(wsz, dev)
        with torch.cuda.device(dev):
            _lib.check(L.dtb_face_adjacency(_lib.ptr(soup), _lib.ptr(fidx), _lib.ptr(counts), B, F, 0 if soup is not None else n_vert,
                                            _lib.ptr(adj_f), _lib.ptr(adj_i), None, _lib.ptr(ws), wsz, _lib.stream_ptr()),
                       "dtb_face_adjacency")
    out = adj_f if want == "f32" else adj_i
    return out if batched else out[0]


def tet_face_adj_m_f_idx(face_fx3x3):
    """``VarianceFunc.apply`` of layers/DefTet/tet_face_adj_m_idx/utils.py:37-70: (F,3,3) -> long (2,E)."""
    n_face = face_fx3x3.shape[0]
    if n_face == 0:
        return torch.zeros(0, device=face_fx3x3.device)
    n_max_nei = 30
    with torch.no_grad():
        adj_idx = face_adjacency_table(face_fx3x3=face_fx3x3.contiguous().float())
        idx = torch.arange(0, n_face, device=face_fx3x3.device, dtype=torch.int32)
        idx = idx.unsqueeze(-1).unsqueeze(-1).expand(-1, n_max_nei, 1)
        mask = (adj_idx >= 0)
        all_adj_idx = torch.cat([idx, adj_idx.int().unsqueeze(-1)], dim=-1)[mask]
        return all_adj_idx.permute(1, 0).long()


# ====================================================================================================
# engine layer (padded-ragged batch)
# ====================================================================================================
class FaceTable:
    """Interior-face table of a grid on the device, int32 (tet_face_fx3, tet_face_tetidx_fx2 of train_multigpu.py:77-82)."""

    def __init__(self, tet_face_fx3, tet_face_tetidx_fx2):
        self.face = _i32c(tet_face_fx3)
        self.face_tet = _i32c(tet_face_tetidx_fx2)
        self.n_face = self.face.shape[0]


def boundary_faces(table: FaceTable, occ_bxt, max_faces):
    """-> faces (B,max_faces,3) i32, counts (B,) i32, overflow (1,) i32 -- get_boundary_index without host sync."""
    _lib.require_cuda(occ_bxt)
    occ = _f32c(occ_bxt)
    B, T = occ.shape
    dev = occ.device
    L = _lib.lib()
    faces = torch.zeros(B, max_faces, 3, device=dev, dtype=torch.int32)
    counts = torch.zeros(B, device=dev, dtype=torch.int32)
    overflow = torch.zeros(1, device=dev, dtype=torch.int32)
    wsz = L.dtb_boundary_faces_workspace(B, table.n_face)
    ws = _ws(wsz, dev)
    with torch.cuda.device(dev):
        _lib.check(L.dtb_boundary_faces(_lib.ptr(table.face), _lib.ptr(table.face_tet), _lib.ptr(occ), B, T, table.n_face, max_faces,
                                        _lib.ptr(faces), _lib.ptr(counts), _lib.ptr(overflow), _lib.ptr(ws), wsz, _lib.stream_ptr()),
                   "dtb_boundary_faces")
    return faces, counts, overflow


def get_boundary_index(tet_face_fx3, tet_idx_fx2, occ_bxn):
    """Drop-in ``DefTet.get_boundary_index`` (deftet.py:186-195): list of B (F_b,3) long tensors."""
    table = FaceTable(tet_face_fx3, tet_idx_fx2)
    faces, counts, _ = boundary_faces(table, occ_bxn, table.n_face)
    n = counts.tolist()
    return [faces[b, :n[b]].long() for b in range(len(n))]


def sample_and_match(pos, faces, counts, u, v, gt, grid_res=0):
    """The index-valued half of the chamfer op, exposed for inspection: -> q (B,Fmax*S,3) sampled surface points and
    nn (B,Fmax*S) int32 index of the nearest gt point of each (only the first S*counts[b] entries of a sample are defined)."""
    pos, gt, u, v = _f32c(pos.detach()), _f32c(gt), _f32c(u), _f32c(v)
    B, V, _ = pos.shape
    Fmax, S, M = faces.shape[1], u.shape[2], gt.shape[1]
    dev = pos.device
    L = _lib.lib()
    q = torch.zeros(B, Fmax * S, 3, device=dev)
    nn = torch.zeros(B, Fmax * S, device=dev, dtype=torch.int32)
    wsz = L.dtb_nearest_neighbor_workspace(B, Fmax * S, M, grid_res)
    ws = _ws(wsz, dev)
    st = _lib.stream_ptr()
    with torch.cuda.device(dev):
        _lib.check(L.dtb_surface_sample(_lib.ptr(pos), _lib.ptr(faces), _lib.ptr(counts), _lib.ptr(u), _lib.ptr(v), B, V, Fmax, S,
                                        _lib.ptr(q), st), "dtb_surface_sample")
        _lib.check(L.dtb_nearest_neighbor_ragged(_lib.ptr(q), _lib.ptr(counts), S, _lib.ptr(gt), _lib.ptr(nn), B, Fmax * S, M,
                                                 grid_res, _lib.ptr(ws), wsz, st), "dtb_nearest_neighbor_ragged")
    return q, nn


def closest_faces(pos, faces, counts, gt, grid_res=0):
    """The index-valued half of the surface-distance op: -> soup (B,Fmax,3,3), squared distance (B,S), closest face id (B,S) f32."""
    pos, gt = _f32c(pos.detach()), _f32c(gt)
    B, V, _ = pos.shape
    Fmax, S = faces.shape[1], gt.shape[1]
    dev = pos.device
    L = _lib.lib()
    soup = torch.zeros(B, Fmax, 3, 3, device=dev)
    cd = torch.empty(B, S, device=dev)
    cf = torch.empty(B, S, device=dev)
    wsz = L.dtb_point_face_distance_workspace(B, S, Fmax, grid_res)
    ws = _ws(wsz, dev)
    st = _lib.stream_ptr()
    with torch.cuda.device(dev):
        _lib.check(L.dtb_face_soup(_lib.ptr(pos), _lib.ptr(faces), _lib.ptr(counts), B, V, Fmax, _lib.ptr(soup), st), "dtb_face_soup")
        _lib.check(L.dtb_point_face_distance_forward(_lib.ptr(gt), _lib.ptr(soup), _lib.ptr(counts), B, S, Fmax, grid_res,
                                                     _lib.ptr(cd), _lib.ptr(cf), _lib.ptr(ws), wsz, st), "dtb_point_face_distance_forward")
    return soup, cd, cf


class _SurfaceChamfer(torch.autograd.Function):
    """pos (B,V,3) -> chamfer (B,): sample S points per boundary face, 1-NN into gt (B,M,3), mean distance."""

    @staticmethod
    def forward(ctx, pos, faces, counts, u, v, gt, grid_res):
        pos, gt, u, v = _f32c(pos), _f32c(gt), _f32c(u), _f32c(v)
        B, V, _ = pos.shape
        Fmax, S, M = faces.shape[1], u.shape[2], gt.shape[1]
        dev = pos.device
        L = _lib.lib()
        q = torch.empty(B, Fmax * S, 3, device=dev)
        nn = torch.empty(B, Fmax * S, device=dev, dtype=torch.int32)
        acc = torch.empty(B, device=dev, dtype=torch.float64)
        loss = torch.empty(B, device=dev)
        wsz = L.dtb_nearest_neighbor_workspace(B, Fmax * S, M, grid_res)
        ws = _ws(wsz, dev)
        st = _lib.stream_ptr()
        with torch.cuda.device(dev):
            _lib.check(L.dtb_surface_sample(_lib.ptr(pos), _lib.ptr(faces), _lib.ptr(counts), _lib.ptr(u), _lib.ptr(v), B, V, Fmax, S,
                                            _lib.ptr(q), st), "dtb_surface_sample")
            _lib.check(L.dtb_nearest_neighbor_ragged(_lib.ptr(q), _lib.ptr(counts), S, _lib.ptr(gt), _lib.ptr(nn), B, Fmax * S, M,
                                                     grid_res, _lib.ptr(ws), wsz, st), "dtb_nearest_neighbor_ragged")
            _lib.check(L.dtb_chamfer_forward(_lib.ptr(q), _lib.ptr(nn), _lib.ptr(gt), _lib.ptr(counts), B, Fmax, S, M, _lib.ptr(acc),
                                             _lib.ptr(loss), st), "dtb_chamfer_forward")
        ctx.save_for_backward(pos, faces, counts, u, v, gt, q, nn)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        pos, faces, counts, u, v, gt, q, nn = ctx.saved_tensors
        B, V, _ = pos.shape
        Fmax, S, M = faces.shape[1], u.shape[2], gt.shape[1]
        grad = torch.zeros(B, V, 4, device=pos.device)            # padded: vector reductions (grad_stride = 4)
        g = _f32c(g_loss)
        with torch.cuda.device(pos.device):
            _lib.check(_lib.lib().dtb_chamfer_backward(_lib.ptr(q), _lib.ptr(nn), _lib.ptr(gt), _lib.ptr(faces), _lib.ptr(counts),
                                                       _lib.ptr(u), _lib.ptr(v), _lib.ptr(g), B, V, Fmax, S, M, _lib.ptr(grad), 4,
                                                       _lib.stream_ptr()), "dtb_chamfer_backward")
        return grad[..., :3], None, None, None, None, None, None


def surface_chamfer(pos, faces, counts, u, v, gt, grid_res=0):
    return _SurfaceChamfer.apply(pos, faces, counts, u, v, gt, int(grid_res))


class _SurfaceDistance(torch.autograd.Function):
    """pos (B,V,3) -> mean_i sqrt(d(gt_i, surface) + 1e-10) (B,)  (point_mesh_distance + means, deftet.py:179-181)."""

    @staticmethod
    def forward(ctx, pos, faces, counts, gt, grid_res):
        pos, gt = _f32c(pos), _f32c(gt)
        B, V, _ = pos.shape
        Fmax, S = faces.shape[1], gt.shape[1]
        dev = pos.device
        L = _lib.lib()
        soup = torch.empty(B, Fmax, 3, 3, device=dev)
        cd = torch.empty(B, S, device=dev)
        cf = torch.empty(B, S, device=dev)
        acc = torch.empty(B, device=dev, dtype=torch.float64)
        loss = torch.empty(B, device=dev)
        wsz = L.dtb_point_face_distance_workspace(B, S, Fmax, grid_res)
        ws = _ws(wsz, dev)
        st = _lib.stream_ptr()
        with torch.cuda.device(dev):
            _lib.check(L.dtb_face_soup(_lib.ptr(pos), _lib.ptr(faces), _lib.ptr(counts), B, V, Fmax, _lib.ptr(soup), st), "dtb_face_soup")
            _lib.check(L.dtb_point_face_distance_forward(_lib.ptr(gt), _lib.ptr(soup), _lib.ptr(counts), B, S, Fmax, grid_res,
                                                         _lib.ptr(cd), _lib.ptr(cf), _lib.ptr(ws), wsz, st),
                       "dtb_point_face_distance_forward")
            _lib.check(L.dtb_sqrt_mean(_lib.ptr(cd), _lib.ptr(counts), B, S, 1e-10, _lib.ptr(acc), _lib.ptr(loss), st), "dtb_sqrt_mean")
        ctx.save_for_backward(pos, faces, gt, soup, cd, cf)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        pos, faces, gt, soup, cd, cf = ctx.saved_tensors
        B, V, _ = pos.shape
        Fmax, S = faces.shape[1], gt.shape[1]
        grad = torch.zeros(B, V, 4, device=pos.device)            # padded: vector reductions (grad_stride = 4)
        g = _f32c(g_loss)
        with torch.cuda.device(pos.device):
            _lib.check(_lib.lib().dtb_point_face_distance_backward_indexed(_lib.ptr(gt), _lib.ptr(soup), _lib.ptr(faces), _lib.ptr(cf),
                                                                           _lib.ptr(cd), _lib.ptr(g), B, S, Fmax, V, _lib.ptr(grad), 4,
                                                                           _lib.stream_ptr()),
                       "dtb_point_face_distance_backward_indexed")
        return grad[..., :3], None, None, None, None


def surface_distance(pos, faces, counts, gt, grid_res=0):
    return _SurfaceDistance.apply(pos, faces, counts, gt, int(grid_res))


class _NormalLoss(torch.autograd.Function):
    """pos (B,V,3) -> normal-consistency loss (B,) over the boundary faces' edge adjacency."""

    @staticmethod
    def forward(ctx, pos, faces, counts):
        pos = _f32c(pos)
        B, V, _ = pos.shape
        Fmax = faces.shape[1]
        dev = pos.device
        L = _lib.lib()
        adj = torch.empty(B, Fmax, 30, device=dev, dtype=torch.int32)
        normals = torch.empty(B, Fmax, 3, device=dev)
        acc = torch.empty(B, 2, device=dev, dtype=torch.float64)
        loss = torch.empty(B, device=dev)
        wsz = L.dtb_face_adjacency_workspace(B, Fmax, V)
        ws = _ws(wsz, dev)
        st = _lib.stream_ptr()
        with torch.cuda.device(dev):
            if Fmax > 0:
                _lib.check(L.dtb_face_adjacency(None, _lib.ptr(faces), _lib.ptr(counts), B, Fmax, V, None, _lib.ptr(adj), None, _lib.ptr(ws),
                                                wsz, st), "dtb_face_adjacency")
            _lib.check(L.dtb_normal_loss_forward(_lib.ptr(pos), _lib.ptr(faces), _lib.ptr(counts), _lib.ptr(adj), B, V, Fmax,
                                                 _lib.ptr(normals), _lib.ptr(acc), _lib.ptr(loss), st), "dtb_normal_loss_forward")
        ctx.save_for_backward(pos, faces, counts, adj, normals, acc)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        pos, faces, counts, adj, normals, acc = ctx.saved_tensors
        B, V, _ = pos.shape
        Fmax = faces.shape[1]
        grad = torch.zeros(B, V, 4, device=pos.device)            # padded: vector reductions (grad_stride = 4)
        gn = torch.empty(B, Fmax, 3, device=pos.device)
        g = _f32c(g_loss)
        with torch.cuda.device(pos.device):
            _lib.check(_lib.lib().dtb_normal_loss_backward(_lib.ptr(pos), _lib.ptr(faces), _lib.ptr(counts), _lib.ptr(adj), _lib.ptr(normals),
                                                           _lib.ptr(acc), _lib.ptr(g), B, V, Fmax, _lib.ptr(gn), _lib.ptr(grad), 4,
                                                           _lib.stream_ptr()), "dtb_normal_loss_backward")
        return grad[..., :3], None, None


def surface_normal_loss(pos, faces, counts):
    return _NormalLoss.apply(pos, faces, counts)


# ====================================================================================================
# host-level mirrors of utils/mesh_utils.py (dense tensors in, same return conventions)
# ====================================================================================================
def face_unit_normals(corner_a, corner_b, corner_c):
    """Unit normals with the reference's regulariser: cross / sqrt(|cross|^2 + 1e-12) (utils/mesh_utils.py:42-53)."""
    cr = torch.linalg.cross(corner_b - corner_a, corner_c - corner_a, dim=-1)
    return cr * torch.rsqrt((cr * cr).sum(dim=-1, keepdim=True) + 1e-12)


def surface_normal_loss_dense(vertices_bxnx3, faces_bxfx3):
    """utils/mesh_utils.py:16-39: mean over edge-adjacent face pairs of 1 - n_i.n_j (adjacency from sample 0)."""
    B, F = faces_bxfx3.shape[0], faces_bxfx3.shape[1]
    tri = vertices_bxnx3[torch.arange(B, device=faces_bxfx3.device).reshape(B, 1, 1), faces_bxfx3.long()]          # (B,F,3,3)
    n = face_unit_normals(tri[:, :, 0], tri[:, :, 1], tri[:, :, 2])
    with torch.no_grad():
        pairs = tet_face_adj_m_f_idx(tri[0].float())
    if pairs.sum() == 0:
        return torch.zeros(B, device=faces_bxfx3.device).float()
    return (1 - (n[:, pairs[0]] * n[:, pairs[1]]).sum(dim=-1)).mean(dim=-1)


def sample_faces_uniform(face_bxfx3x3, each_face_num=20):
    """utils/mesh_utils.py:290-299; the two torch.rand calls keep the reference's order (sqrt'ed u first, then v)."""
    B, F = face_bxfx3x3.shape[0], face_bxfx3x3.shape[1]
    u = torch.rand(size=(B, F, each_face_num, 1), device=face_bxfx3x3.device).sqrt()
    v = torch.rand(size=(B, F, each_face_num, 1), device=face_bxfx3x3.device)
    w0, w1, w2 = 1 - u, u * (1 - v), u * v
    return w0 * face_bxfx3x3[:, :, 0:1] + w1 * face_bxfx3x3[:, :, 1:2] + w2 * face_bxfx3x3[:, :, 2:3]


def one_sided_chamfer_dense(a_bxnx3, b_bxmx3, eps=1e-10):
    """utils/mesh_utils.py:360-366: distance of every a to its nearest b (gradient flows to a only through the difference)."""
    from .search import nearest_neighbor_index
    idx = nearest_neighbor_index(a_bxnx3, b_bxmx3).long()
    nearest = torch.gather(b_bxmx3, 1, idx.unsqueeze(-1).expand(-1, -1, 3))
    return ((a_bxnx3 - nearest).pow(2).sum(dim=-1) + eps).sqrt()


def point_to_faces_distance_dense(a_bxnx3, mesh_bxfx3x3, eps=1e-10):
    """utils/mesh_utils.py:368-374."""
    n_face = torch.full((mesh_bxfx3x3.shape[0],), float(mesh_bxfx3x3.shape[1]), device=mesh_bxfx3x3.device)
    d, _ = tet_analytic_distance_f_batch(a_bxnx3, mesh_bxfx3x3, n_face)
    return (d + eps).sqrt()
