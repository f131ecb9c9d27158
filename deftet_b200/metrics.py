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
