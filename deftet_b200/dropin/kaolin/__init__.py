"""Minimal ``kaolin`` shim exposing the two operations the DefTet hot path calls (parity unpinned).
Only used when the real Kaolin is absent: put deftet_b200/dropin AFTER site-packages to prefer a real install."""
from . import ops, render  # noqa: F401
