"""kal.metrics.pointcloud shim (utils/point_cloud_utils.py:86-87,111-112,119,125): parity unpinned, see deftet_b200/metrics.py."""
from deftet_b200.metrics import sided_distance  # noqa: F401
