"""Drop-in for the hot function of diff_render/diftet_6_subdiv/3_model/utils_tetsv.py."""
import numpy as np
import torch

from deftet_b200 import topology


def tet_adj_share(tet_list_tx4, n_point):
    """utils_tetsv.py:16-77 -> (adj_list: four scipy (T,T) matrices, one per local face; tet_neighbour_idx (T,4) int64, -1 = none).
    Column i of tet_neighbour_idx is the tet across local face i (the reference fills its columns in Python dict order; both are
    only ever reduced over the row)."""
    from scipy.sparse import coo_matrix
    nbr = topology.tet_neighbours(torch.from_numpy(np.ascontiguousarray(tet_list_tx4)).cuda(), int(n_point)).cpu().numpy().astype(np.int64)
    n_tet = nbr.shape[0]
    adj_list = []
    for i in range(4):
        rows = np.flatnonzero(nbr[:, i] >= 0)
        adj_list.append(coo_matrix((np.ones(rows.shape[0]), (rows, nbr[rows, i])), shape=(n_tet, n_tet)))
    return adj_list, nbr


# every other name of the reference module comes from the checkout at DEFTET_REFERENCE_ROOT (see dropin/_fallthrough.py)
from _fallthrough import adopt_reference_module as _adopt, missing_attribute as _missing  # noqa: E402

_REPLACED = ('tet_adj_share',)
_reference = _adopt(globals(), 'diff_render/diftet_6_subdiv/3_model/utils_tetsv.py', _REPLACED)


def __getattr__(name):
    raise _missing(__name__, name)
