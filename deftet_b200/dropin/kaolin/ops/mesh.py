from deftet_b200.render import check_sign  # noqa: F401
from deftet_b200.metrics import index_vertices_by_faces  # noqa: F401,E402
from deftet_b200.metrics import face_areas, face_normals, sample_points  # noqa: F401,E402
