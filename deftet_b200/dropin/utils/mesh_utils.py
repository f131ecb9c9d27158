"""Drop-in for reference utils/mesh_utils.py.  The hot-path functions (:16-53 normal loss, :290-299 surface sampling,
:360-374 point-point / point-mesh distance) are re-implemented on the sm_100a kernels (deftet_b200/surface.py); every other
name of the reference module (``save_mesh``, ``save_tet_face``, ``loadobj`` ...; dataloader.py:15, eval.py:305) is taken from the
checkout at DEFTET_REFERENCE_ROOT, so the shadow is complete."""
from deftet_b200.search import NearestNeighbor  # noqa: F401
from deftet_b200.surface import face_unit_normals as get_normal  # noqa: F401
from deftet_b200.surface import one_sided_chamfer_dense as point_point_distance  # noqa: F401
from deftet_b200.surface import point_to_faces_distance_dense as point_mesh_distance  # noqa: F401
from deftet_b200.surface import sample_faces_uniform as sample_surf_point_batch  # noqa: F401
from deftet_b200.surface import surface_normal_loss_dense as get_surface_normal_loss  # noqa: F401
from deftet_b200.surface import tet_analytic_distance_f_batch, tet_face_adj_m_f_idx  # noqa: F401

from _fallthrough import adopt_reference_module as _adopt, missing_attribute as _missing

EPS = 1e-10

_REPLACED = ("NearestNeighbor", "get_normal", "point_point_distance", "point_mesh_distance", "sample_surf_point_batch",
             "get_surface_normal_loss", "tet_analytic_distance_f_batch", "tet_face_adj_m_f_idx")
_reference = _adopt(globals(), "utils/mesh_utils.py", _REPLACED)


def __getattr__(name):
    raise _missing(__name__, name)
