"""Drop-in shadow of the reference package layers/pv_module: only functional/devoxelization.py is replaced; with
DEFTET_REFERENCE_ROOT set every other sub-module (pvconv, pointnet, ...) and the package's own exports come from the checkout."""
from _fallthrough import extend as _extend, exec_reference_init as _exec_init
_extend(__path__, "layers/pv_module")
_exec_init(globals(), "layers/pv_module")
