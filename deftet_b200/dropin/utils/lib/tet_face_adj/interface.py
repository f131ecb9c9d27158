"""Drop-in for reference utils/lib/tet_face_adj/interface.py: scipy CSR (4T x 4T) face-face adjacency."""
import numpy as np
from scipy.sparse import coo_matrix

from deftet_b200 import builders


class Tet_face_adj:
    def run(self, n_point, tet_list):
        assert tet_list.dtype == np.int32
        n_face = tet_list.shape[0] * 4
        face_edge, n = builders.host_run("tet_face_adj", tet_list, n_point, n_face * 50, 2)
        v = np.ones(n)
        return coo_matrix((v, (face_edge[:n, 0], face_edge[:n, 1])), shape=(n_face, n_face)).tocsr()
