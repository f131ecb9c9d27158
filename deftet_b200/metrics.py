"""Evaluation metrics of the reference (utils/point_cloud_utils.py) on the sm_100a kernels.

Same function names, arguments and return values as the reference module; the two Kaolin primitives underneath are
``sided_distance`` (1-NN through the binned A2 kernel, ``search.nearest_neighbor_index``) and ``point_to_mesh_distance``
(``csrc/metrics.cu``).  Kaolin is un-vendored / un-pinned in the reference: parity for these two is defined by oracle/metrics.py
("parity unpinned"); everything above them is the reference's own arithmetic restated."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, search
from .search import _f32c

esp = 1e-15


@_lib.register_signatures
def _metric_sigs(lib, sig):
    vp, i = C.c_void_p, C.c_int
    sig("dtb_point_to_mesh_distance", i, vp, vp, i, i, i, vp, vp, vp, vp)


def sided_distance(p1, p2):
    """``kal.metrics.pointcloud.sided_distance`` -> (squared distance (B,N) from each point of p1 to its nearest point of p2,
    index (B,N) long of that point; lowest index on exact ties)."""
    _lib.require_cuda(p1, p2)
    p1, p2 = _f32c(p1), _f32c(p2)
    idx = search.nearest_neighbor_index(p1, p2).long()
    closest = torch.gather(p2, 1, idx.unsqueeze(-1).expand(-1, -1, 3))
    return ((p1 - closest) ** 2).sum(dim=-1), idx


def index_vertices_by_faces(vertices_features, faces):
    """``kal.ops.mesh.index_vertices_by_faces``: (B,V,C), (F,3) -> (B,F,3,C)."""
    B, C_ = vertices_features.shape[0], vertices_features.shape[-1]
    return vertices_features[:, faces.reshape(-1).long()].reshape(B, faces.shape[0], 3, C_)


def face_normals(face_vertices, unit=False):
    """``kal.ops.mesh.face_normals``: (B,F,3,3) -> (B,F,3), (v1 - v0) x (v2 - v0), optionally normalised (dataloader.py:80-81)."""
    n = torch.cross(face_vertices[:, :, 1] - face_vertices[:, :, 0], face_vertices[:, :, 2] - face_vertices[:, :, 0], dim=-1)
    if unit:
        n = n / (n.norm(dim=-1, keepdim=True) + 1e-10)
    return n


def face_areas(vertices, faces):
    """``kal.ops.mesh.face_areas``: (B,V,3), (F,3) -> (B,F)."""
    return 0.5 * face_normals(index_vertices_by_faces(vertices, faces)).norm(dim=-1)


def sample_points(vertices, faces, num_samples, areas=None, face_features=None):
    """``kal.ops.mesh.sample_points`` as eval.py:244-245 and dataloader.py:76-77 call it: ``num_samples`` points uniformly on the surface
    -> (points (B,S,3), face_choices (B,S) long[, interpolated face_features (B,S,C)]).  Faces are drawn proportionally to their area,
    the position inside a face with the recipe of the reference's own sampler (utils/mesh_utils.py:290-299: u = sqrt(rand), v = rand,
    (1-u) a + u (1-v) b + u v c).  Plain tensor glue on the caller's device (gather + elementwise, like index_vertices_by_faces); the
    random stream is torch's, so individual samples differ from Kaolin's while the distribution is the same (parity unpinned)."""
    faces = faces.long()
    B, F = vertices.shape[0], faces.shape[0]
    if areas is None:
        areas = face_areas(vertices, faces)
    if F == 0 or num_samples == 0:
        raise RuntimeError("sample_points needs at least one face and one sample")
    w = areas.reshape(B, F).clamp(min=0) + 1e-30                     # a mesh of degenerate faces still samples (uniformly over faces)
    face_choices = torch.multinomial(w, num_samples, replacement=True)
    fv = index_vertices_by_faces(vertices, faces)                    # (B,F,3,3)
    sel = torch.gather(fv, 1, face_choices.reshape(B, num_samples, 1, 1).expand(-1, -1, 3, 3))
    u = torch.sqrt(torch.rand(B, num_samples, 1, device=vertices.device, dtype=vertices.dtype))
    v = torch.rand(B, num_samples, 1, device=vertices.device, dtype=vertices.dtype)
    w0, w1, w2 = 1 - u, u * (1 - v), u * v
    points = w0 * sel[:, :, 0] + w1 * sel[:, :, 1] + w2 * sel[:, :, 2]
    if face_features is None:
        return points, face_choices
    ff = torch.gather(face_features, 1, face_choices.reshape(B, num_samples, 1, 1).expand(-1, -1, 3, face_features.shape[-1]))
    return points, face_choices, w0 * ff[:, :, 0] + w1 * ff[:, :, 1] + w2 * ff[:, :, 2]


def point_to_mesh_distance(pointclouds, face_vertices):
    """``kal.metrics.trianglemesh.point_to_mesh_distance``: (B,P,3), (B,F,3,3) -> (squared distance (B,P), face index (B,P) long,
    distance type (B,P) int32: 0 face interior, 1-3 vertex, 4-6 edge)."""
    _lib.require_cuda(pointclouds, face_vertices)
    pts, fv = _f32c(pointclouds), _f32c(face_vertices)
    B, P, F = pts.shape[0], pts.shape[1], fv.shape[1]
    dev = pts.device
    dist = torch.empty(B, P, device=dev)
    fidx = torch.empty(B, P, device=dev, dtype=torch.int64)
    dtype_ = torch.empty(B, P, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().dtb_point_to_mesh_distance(_lib.ptr(pts), _lib.ptr(fv), B, P, F, _lib.ptr(dist), _lib.ptr(fidx), _lib.ptr(dtype_),
                                                         _lib.stream_ptr()), "dtb_point_to_mesh_distance")
    return dist, fidx, dtype_


# ---- utils/point_cloud_utils.py ---------------------------------------------------------------------------------------------
def iou(points1, points2, thresh=.5):
    """point_cloud_utils.py:12-42."""
    a, b = (points1 > thresh).reshape(-1), (points2 > thresh).reshape(-1)
    assert a.shape == b.shape, 'points1 and points2 must have the same shape'
    return (a & b).float().sum() / (a | b).float().sum()


def hausdorff_distance(mesh_a_v, mesh_a_f, mesh_b_v, mesh_b_f, pts_a, pts_b):
    """point_cloud_utils.py:46-61 -> (average symmetric distance, symmetric max distance)."""
    d_a, _, _ = point_to_mesh_distance(pts_b.unsqueeze(0), index_vertices_by_faces(mesh_a_v.unsqueeze(0), mesh_a_f))
    d_b, _, _ = point_to_mesh_distance(pts_a.unsqueeze(0), index_vertices_by_faces(mesh_b_v.unsqueeze(0), mesh_b_f))
    ra, rb = torch.sqrt(d_a + esp), torch.sqrt(d_b + esp)
    return torch.mean((ra + rb) / 2), (torch.max(ra) + torch.max(rb)) / 2


def f_score(gt_points, pred_points, radius=0.01, extend=False):
    """point_cloud_utils.py:66-108."""
    pred_distances = torch.sqrt(sided_distance(gt_points, pred_points)[0] + esp)
    gt_distances = torch.sqrt(sided_distance(pred_points, gt_points)[0] + esp)
    if extend:
        fp = (gt_distances > radius).float().sum()
        tp = (gt_distances <= radius).float().sum()
        precision = tp / (tp + fp)
        tp = (pred_distances <= radius).float().sum()
        fn = (pred_distances > radius).float().sum()
        recall = tp / (tp + fn)
    else:
        fn = torch.sum(pred_distances > radius)
        fp = torch.sum(gt_distances > radius).float()
        tp = torch.sum(gt_distances <= radius).float()
        precision = tp / (tp + fp)
        recall = tp / (tp + fn)
    return 2 * (precision * recall) / (precision + recall + 1e-8)


def chamfer_distance(S1, S2):
    """point_cloud_utils.py:110-115."""
    return (torch.sqrt(sided_distance(S1, S2)[0] + esp).mean() + torch.sqrt(sided_distance(S2, S1)[0] + esp).mean()) / 2


def chamfer_distance_l1(S1, S2):
    """point_cloud_utils.py:118-130 (batch element 0, like the reference)."""
    _, idx1 = sided_distance(S1, S2)
    d12 = torch.abs(S1 - torch.index_select(S2[0], 0, idx1[0]).unsqueeze(0)).sum(dim=-1)
    _, idx2 = sided_distance(S2, S1)
    d21 = torch.abs(S2 - torch.index_select(S1[0], 0, idx2[0]).unsqueeze(0)).sum(dim=-1)
    return d12.mean() + d21.mean()
