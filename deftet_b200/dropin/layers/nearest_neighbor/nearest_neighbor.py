"""Drop-in for reference layers/nearest_neighbor/nearest_neighbor.py."""
from deftet_b200.search import NearestNeighbor, NearestNeighborFunction  # noqa: F401
