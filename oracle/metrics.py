"""Oracle for the evaluation metrics -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the two primitives are Kaolin's (kal.metrics.pointcloud.sided_distance,
kal.metrics.trianglemesh.point_to_mesh_distance; call sites utils/point_cloud_utils.py:48-130); Kaolin is un-vendored and
un-pinned in the reference (README.md:30) and absent here, and the reference has no test holding their outputs.  Restated from the
call-site contract: squared Euclidean distance to the nearest point / to the closest point of the closest triangle, fp64,
first strict minimum in index order.  The functions above them (f_score, chamfer_distance, ...) are the reference's arithmetic.
"""
import numpy as np


def sided_distance(p1, p2):
    p1, p2 = np.asarray(p1, dtype=np.float64), np.asarray(p2, dtype=np.float64)
    d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
    return d.min(-1), d.argmin(-1)


def _closest_point_on_triangles(p, a, b, c):
    """p (P,1,3), a/b/c (1,F,3) -> squared distance (P,F) (region classification, vectorised)."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = (ab * ap).sum(-1), (ac * ap).sum(-1)
    bp = p - b
    d3, d4 = (ab * bp).sum(-1), (ac * bp).sum(-1)
    cp = p - c
    d5, d6 = (ab * cp).sum(-1), (ac * cp).sum(-1)
    vc, vb, va = d1 * d4 - d3 * d2, d5 * d2 - d1 * d6, d3 * d6 - d5 * d4
    with np.errstate(divide="ignore", invalid="ignore"):
        denom = 1.0 / (va + vb + vc)
        q = a + ab * (vb * denom)[..., None] + ac * (vc * denom)[..., None]                       # interior
        w = ((d4 - d3) / ((d4 - d3) + (d5 - d6)))[..., None]
        q = np.where(((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0))[..., None], b + w * (c - b), q)
        w = (d2 / (d2 - d6))[..., None]
        q = np.where(((vb <= 0) & (d2 >= 0) & (d6 <= 0))[..., None], a + w * ac, q)
        q = np.where(((d6 >= 0) & (d5 <= d6))[..., None], np.broadcast_to(c, q.shape), q)
        v = (d1 / (d1 - d3))[..., None]
        q = np.where(((vc <= 0) & (d1 >= 0) & (d3 <= 0))[..., None], a + v * ab, q)
        q = np.where(((d3 >= 0) & (d4 <= d3))[..., None], np.broadcast_to(b, q.shape), q)
        q = np.where(((d1 <= 0) & (d2 <= 0))[..., None], np.broadcast_to(a, q.shape), q)
    return ((p - q) ** 2).sum(-1)


def point_to_mesh_distance(points, face_vertices):
    """(B,P,3), (B,F,3,3) -> (squared distance (B,P), face index (B,P)) in fp64."""
    pts, fv = np.asarray(points, dtype=np.float64), np.asarray(face_vertices, dtype=np.float64)
    out_d, out_f = [], []
    for b in range(pts.shape[0]):
        d = _closest_point_on_triangles(pts[b][:, None, :], fv[b][None, :, 0], fv[b][None, :, 1], fv[b][None, :, 2])
        out_d.append(d.min(-1)); out_f.append(d.argmin(-1))
    return np.stack(out_d), np.stack(out_f)
