"""Drop-in for reference layers/nearest_neighbor/nearest_neighbor.py."""
from deftet_b200.search import NearestNeighbor, NearestNeighborFunction, nearest_neighbor_index  # noqa: F401


class _Native:
    """Stands in for the pybind module ``nearest_neighbor_cuda`` (nearest_neighbor.cpp:34-52): in-place int32 result."""

    @staticmethod
    def forward(queries, points, result, batch_size, num_queries, num_points, dim):
        if dim != 3:
            raise RuntimeError("Currently only 3D points are supported")
        result.copy_(nearest_neighbor_index(queries.reshape(batch_size, num_queries, 3), points.reshape(batch_size, num_points, 3)))


native = _Native()
