"""Drop-in for reference utils/lib/tet_adj_share/interface.py: list of 4 scipy COO matrices (one per local face slot)."""
import numpy as np
from scipy.sparse import coo_matrix

from deftet_b200 import builders


class Tet_adj_share:
    def run(self, tet_list, n_point):
        assert tet_list.dtype == np.int32
        index_list, n = builders.host_run("tet_adj_share", tet_list, n_point, tet_list.shape[0] * 8, 3)
        index_list = index_list[:n * 2]
        n_tet = tet_list.shape[0]
        value = np.ones(index_list.shape[0])
        return [coo_matrix((value[index_list[:, 2] == i], (index_list[:, 0][index_list[:, 2] == i], index_list[:, 1][index_list[:, 2] == i])),
                           shape=(n_tet, n_tet)) for i in range(4)]
