"""Drop-in for diff_render/diftet_6_subdiv/3_model/prepare_for_wz.py: the reference's numpy-in / numpy-out signatures and dtypes, with
the Python dict / dense-matrix / O(E*T) loops replaced by the GPU builders of deftet_b200.topology."""
import numpy as np
import torch

from deftet_b200 import topology
from deftet_b200.grid import read_tet_file


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def _np64(t):
    return t.cpu().numpy().astype(np.int64)


def read_tetrahedron(file_name, res=0.02):
    """prepare_for_wz.py:18-45."""
    vertices, tets = read_tet_file(file_name)
    vertices = vertices.astype(np.float32)
    vertices[vertices <= (0 + res / 4.0)] = 0
    vertices[vertices >= (1 - res / 4.0)] = 1
    mask = np.logical_and(vertices < 1, vertices > 0)
    return vertices, tets.astype(np.int64), mask


def tet_to_face_idx(n_point, tet_list, with_boundary=False):
    """prepare_for_wz.py:49-108."""
    f3, ft2, fs2 = topology.tet_to_face_idx(int(n_point), _dev(tet_list), with_boundary=True)
    interior = ft2[:, 1] >= 0
    n_int = int(interior.sum())
    print('Cnt neighbor tet: ', [int(f3.shape[0]) - n_int, n_int, 0])
    if not with_boundary:
        f3, ft2, fs2 = f3[interior], ft2[interior], fs2[interior]
    return _np64(f3), _np64(ft2), _np64(fs2)


def generate_point_adj_idx(n_point, tet_list):
    """prepare_for_wz.py:121-137."""
    table, adjsum = topology.generate_point_adj_idx(int(n_point), _dev(tet_list))
    return _np64(table), adjsum.cpu().numpy()


def delete_tet(tet_list_tx4, tet_weights_tx4, thres=0.01):
    """prepare_for_wz.py:171-181 (already a vectorised numpy one-liner in the reference; kept on the host for callers that hold the
    gathered (T, 4^(L+1)) weight table -- Deftet.deletetet itself uses topology.delete_tet_by_weight and never builds it)."""
    return tet_list_tx4[np.max(tet_weights_tx4, axis=1) > thres]


def generate_edge(tet_list_tx4):
    """prepare_for_wz.py:186-205."""
    n_point = int(np.max(tet_list_tx4)) + 1
    return _np64(topology.tet_edges(_dev(tet_list_tx4), n_point)[0])


def generate_tet_edge_idx(tet_list_tx4, edges_all_ex2):
    """prepare_for_wz.py:225-238 (edges_all_ex2 must be generate_edge's output, as in the reference)."""
    n_point = int(np.max(tet_list_tx4)) + 1
    return _np64(topology.tet_edges(_dev(tet_list_tx4), n_point)[1])


def generate_subdivision(tet_list_tx4, tet_points_px3, tet_feat_pxk, tet_list_subdiv_sig=None):
    """prepare_for_wz.py:257-301."""
    sig = None if tet_list_subdiv_sig is None else _dev(np.asarray(tet_list_subdiv_sig, dtype=bool))
    p, f, t = topology.generate_subdivision(_dev(tet_list_tx4), _dev(tet_points_px3, torch.float32), _dev(tet_feat_pxk, torch.float32), sig)
    return p.cpu().numpy().astype(tet_points_px3.dtype), f.cpu().numpy().astype(tet_feat_pxk.dtype), _np64(t)


# every other name of the reference module comes from the checkout at DEFTET_REFERENCE_ROOT (see dropin/_fallthrough.py)
from _fallthrough import adopt_reference_module as _adopt, missing_attribute as _missing  # noqa: E402

_REPLACED = ('read_tetrahedron', 'tet_to_face_idx', 'generate_point_adj_idx', 'delete_tet', 'generate_edge', 'generate_tet_edge_idx', 'generate_subdivision')
_reference = _adopt(globals(), 'diff_render/diftet_6_subdiv/3_model/prepare_for_wz.py', _REPLACED)


def __getattr__(name):
    raise _missing(__name__, name)
