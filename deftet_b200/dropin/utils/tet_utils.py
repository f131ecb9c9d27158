"""Drop-in for the hot-path functions of reference utils/tet_utils.py; everything else falls through to the
reference module when DEFTET_REFERENCE_ROOT is set (see _fallthrough.py)."""
import numpy as np
import torch

from deftet_b200 import builders

from utils.lib.tet_adj_share.interface import Tet_adj_share
from utils.lib.tet_face_adj.interface import Tet_face_adj
from utils.lib.tet_point_adj.interface import Tet_point_adj

c_tet_point_adj = Tet_point_adj()
c_tet_face_adj = Tet_face_adj()
c_obj_tet_adj_share = Tet_adj_share()


def scaler_triplet_produt(a, b, c):
    return (a * torch.linalg.cross(b, c, dim=-1)).sum(dim=-1)


def bary_centric_tet(a, b, c, d, p):
    """Barycentric weights of p in tet (a,b,c,d) as ratios of signed volumes (reference utils/tet_utils.py:28-45)."""
    inv6v = scaler_triplet_produt(b - a, c - a, d - a).reciprocal()
    wa = scaler_triplet_produt(p - b, d - b, c - b)
    wb = scaler_triplet_produt(p - a, c - a, d - a)
    wc = scaler_triplet_produt(p - a, d - a, b - a)
    wd = scaler_triplet_produt(p - a, b - a, c - a)
    return wa * inv6v, wb * inv6v, wc * inv6v, wd * inv6v


def c_tet_to_adj_sparse(points, tet_list, normalize=True):
    return c_tet_point_adj.run(points.shape[0], tet_list.astype(np.int32), normalize)


def tet_to_adj_sparse(points, tet_list, normalize=False):
    return c_tet_point_adj.run(points.shape[0], np.asarray(tet_list).astype(np.int32), normalize)


def c_tet_to_face_adj_sparse(points, tet_list):
    return c_tet_face_adj.run(points.shape[0], tet_list.astype(np.int32))


def c_tet_adj_share(points, tet_list):
    return c_obj_tet_adj_share.run(tet_list.astype(np.int32), points.shape[0])


def tet_to_face(n_point, tet_list):
    """GPU replacement of the dict loop (utils/tet_utils.py:208-256); same four numpy arrays, same order."""
    f3, ft2, fs2, bnd = builders.tet_to_face(int(n_point), torch.from_numpy(np.ascontiguousarray(tet_list)).cuda())
    print('Cnt neighbor tet: ', [int(bnd.shape[0]), int(f3.shape[0]), 0])
    return tuple(x.cpu().numpy().astype(np.int64) for x in (f3, ft2, fs2, bnd))
