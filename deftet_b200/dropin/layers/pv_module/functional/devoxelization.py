"""Drop-in for reference layers/pv_module/functional/devoxelization.py."""
from deftet_b200.devox import trilinear_devoxelize  # noqa: F401

__all__ = ['trilinear_devoxelize', 'trilinear_devoxelize_ori']


def trilinear_devoxelize_ori(features, coords, resolution, is_training=True):
    """devoxelization.py:9-45 (`TrilinearDevoxelization.apply`): PVCNN's original kernel.  Dead in the reference (every caller uses the
    grid_sample definition at :47-53, which re-binds the public name); inside [0, R-1]^3 the two agree, so it maps to the same kernel."""
    return trilinear_devoxelize(features, coords, resolution, is_training)
