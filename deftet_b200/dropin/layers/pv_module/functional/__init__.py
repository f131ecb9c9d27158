"""Drop-in shadow of layers/pv_module/functional: `trilinear_devoxelize` comes from deftet_b200, the rest (ball_query, voxelization,
...: PVCNN's own CUDA extension, out of scope) from the reference checkout when DEFTET_REFERENCE_ROOT is set."""
from _fallthrough import extend as _extend, exec_reference_init as _exec_init
_extend(__path__, "layers/pv_module/functional")
if not _exec_init(globals(), "layers/pv_module/functional"):
    from layers.pv_module.functional.devoxelization import trilinear_devoxelize, trilinear_devoxelize_ori  # noqa: F401
