from deftet_b200.render import check_sign  # noqa: F401
from deftet_b200.metrics import index_vertices_by_faces  # noqa: F401,E402
