from .nearest_neighbor import NearestNeighbor, NearestNeighborFunction  # noqa: F401
