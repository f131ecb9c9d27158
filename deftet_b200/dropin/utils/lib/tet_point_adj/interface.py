"""Drop-in for reference utils/lib/tet_point_adj/interface.py (same class, same return types)."""
import numpy as np
import torch

from deftet_b200 import builders


class Tet_point_adj:
    def run(self, n_point, tet_list, normalize=False):
        assert tet_list.dtype == np.int32
        dev = torch.device("cuda")
        adj = builders.tet_to_adj_sparse(n_point, torch.from_numpy(np.ascontiguousarray(tet_list)).to(dev), normalize)
        return adj.cpu()
