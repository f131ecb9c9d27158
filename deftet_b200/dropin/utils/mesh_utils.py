"""Drop-in for the hot-path functions of reference utils/mesh_utils.py (:16-53, :290-299, :360-374); the
implementations live in deftet_b200/surface.py."""
from deftet_b200.search import NearestNeighbor  # noqa: F401
from deftet_b200.surface import face_unit_normals as get_normal  # noqa: F401
from deftet_b200.surface import one_sided_chamfer_dense as point_point_distance  # noqa: F401
from deftet_b200.surface import point_to_faces_distance_dense as point_mesh_distance  # noqa: F401
from deftet_b200.surface import sample_faces_uniform as sample_surf_point_batch  # noqa: F401
from deftet_b200.surface import surface_normal_loss_dense as get_surface_normal_loss  # noqa: F401
from deftet_b200.surface import tet_analytic_distance_f_batch, tet_face_adj_m_f_idx  # noqa: F401

EPS = 1e-10
