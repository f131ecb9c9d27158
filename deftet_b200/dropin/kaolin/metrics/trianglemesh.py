"""kal.metrics.trianglemesh shim (utils/point_cloud_utils.py:49-53): parity unpinned, see deftet_b200/metrics.py."""
from deftet_b200.metrics import point_to_mesh_distance  # noqa: F401
