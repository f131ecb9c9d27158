"""Drop-in for reference utils/tet_utils.py: the builders (A10-A13) run on the GPU through the ``run.so`` shims / C ABI with the
reference's signatures and return types; every other name of the reference module (``read_tet``, ``save_tet``, ``get_tet_adj``,
``get_face_use_occ`` ...) is taken from the checkout at DEFTET_REFERENCE_ROOT, with the replaced functions injected into it."""
import numpy as np
import torch

from deftet_b200 import builders

from utils.lib.tet_adj_share.interface import Tet_adj_share
from utils.lib.tet_face_adj.interface import Tet_face_adj
from utils.lib.tet_point_adj.interface import Tet_point_adj

from _fallthrough import adopt_reference_module as _adopt, missing_attribute as _missing

c_tet_point_adj = Tet_point_adj()
c_tet_face_adj = Tet_face_adj()
c_obj_tet_adj_share = Tet_adj_share()


def convert_torch_sparse(adj):
    """scipy COO -> torch sparse float (reference utils/matrix_utils.py:14-20)."""
    idx = np.stack([adj.row, adj.col], axis=0)
    return torch.sparse_coo_tensor(torch.from_numpy(idx).long(), torch.from_numpy(adj.data).float(), tuple(adj.shape))


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
    """reference utils/tet_utils.py:94-95"""
    return c_tet_point_adj.run(points.shape[0], tet_list.astype(np.int32), normalize)


def tet_to_adj_sparse(points, tet_list, normalize=False):
    """reference utils/tet_utils.py:47-92 (pure-Python twin of the C builder: same sparse matrix)"""
    return c_tet_point_adj.run(points.shape[0], np.asarray(tet_list).astype(np.int32), normalize)


def c_tet_to_face_adj_sparse(points, tet_list):
    """reference utils/tet_utils.py:203-205"""
    return c_tet_face_adj.run(points.shape[0], tet_list.astype(np.int32))


def c_tet_adj_share(tet_list, n_point, torch_t=True):
    """reference utils/tet_utils.py:371-375: 4 sparse (T,T) matrices, one per local face slot; torch sparse unless torch_t=False."""
    adj_list = c_obj_tet_adj_share.run(np.asarray(tet_list).astype(np.int32), n_point)
    if torch_t:
        adj_list = [convert_torch_sparse(adj) for adj in adj_list]
    return adj_list


def tet_adj_share(tet_list, n_point):
    """reference utils/tet_utils.py:318-367 (pure-Python twin; measured identical to the C builder)"""
    return c_tet_adj_share(tet_list, n_point, True)


def tet_to_face(n_point, tet_list):
    """GPU replacement of the dict loop (utils/tet_utils.py:208-256); same four numpy arrays, same order."""
    f3, ft2, fs2, bnd = builders.tet_to_face(int(n_point), torch.from_numpy(np.ascontiguousarray(tet_list)).cuda())
    print('Cnt neighbor tet: ', [int(bnd.shape[0]), int(f3.shape[0]), 0])
    return tuple(x.cpu().numpy().astype(np.int64) for x in (f3, ft2, fs2, bnd))


_REPLACED = ("c_tet_point_adj", "c_tet_face_adj", "c_obj_tet_adj_share", "scaler_triplet_produt", "bary_centric_tet",
             "c_tet_to_adj_sparse", "tet_to_adj_sparse", "c_tet_to_face_adj_sparse", "c_tet_adj_share", "tet_adj_share", "tet_to_face")
_reference = _adopt(globals(), "utils/tet_utils.py", _REPLACED)


def __getattr__(name):
    raise _missing(__name__, name)
