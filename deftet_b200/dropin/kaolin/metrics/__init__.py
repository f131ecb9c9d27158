from . import pointcloud, trianglemesh  # noqa: F401
