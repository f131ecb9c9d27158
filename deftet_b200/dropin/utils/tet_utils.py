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
    return torch.sum(a * torch.cross(b, c, dim=-1), dim=-1)


def bary_centric_tet(a, b, c, d, p):          # utils/tet_utils.py:28-45 (pure torch in the reference too)
    vap, vbp = p - a, p - b
    vab, vac, vad = b - a, c - a, d - a
    vbc, vbd = c - b, d - b
    v6 = 1 / scaler_triplet_produt(vab, vac, vad)
    return (scaler_triplet_produt(vbp, vbd, vbc) * v6, scaler_triplet_produt(vap, vac, vad) * v6,
            scaler_triplet_produt(vap, vad, vab) * v6, scaler_triplet_produt(vap, vab, vac) * v6)


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
