"""Minimal ``kaolin`` shim exposing the operations DefTet calls (check_sign, deftet_sparse_render; the metrics of utils/point_cloud_utils.py) (parity unpinned).
Only used when the real Kaolin is absent: put deftet_b200/dropin AFTER site-packages to prefer a real install."""
from . import metrics, ops, render  # noqa: F401
