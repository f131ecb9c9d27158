"""Drop-in for reference layers/pv_module/functional/backend.py (:1-27), which JIT-compiles PVCNN's own CUDA extension
(`_pvcnn_backend`: ball query, grouping, voxelisation, ...) at import time.  That extension belongs to the point-cloud ENCODER
(out of scope) -- but importing train_multigpu.py pulls it in through layers/pc_model.py:11, so the
import must not fail where the extension cannot be built.  Behaviour:
  * sources present next to the checkout's backend.py and DEFTET_BUILD_PVCNN=1  -> build and use the reference's own extension;
  * otherwise                                                                   -> a placeholder that raises on first USE.
`trilinear_devoxelize` never goes through this object: deftet_b200 replaces it (devoxelization.py)."""
import os

__all__ = ['_backend']


class _MissingBackend:
    def __getattr__(self, name):
        raise RuntimeError("PVCNN backend function %r was called, but the reference's _pvcnn_backend extension (point-cloud encoder, "
                           "outside the deftet_b200 scope) is not built; set DEFTET_BUILD_PVCNN=1 with a full DefTet checkout at "
                           "DEFTET_REFERENCE_ROOT to JIT-build it" % name)


def _load():
    root = os.environ.get("DEFTET_REFERENCE_ROOT")
    if not root or os.environ.get("DEFTET_BUILD_PVCNN") != "1":
        return _MissingBackend()
    src = os.path.join(root, "layers", "pv_module", "functional", "src")
    files = ['ball_query/ball_query.cpp', 'ball_query/ball_query.cu', 'grouping/grouping.cpp', 'grouping/grouping.cu',
             'interpolate/neighbor_interpolate.cpp', 'interpolate/neighbor_interpolate.cu', 'interpolate/trilinear_devox.cpp',
             'interpolate/trilinear_devox.cu', 'sampling/sampling.cpp', 'sampling/sampling.cu', 'voxelization/vox.cpp',
             'voxelization/vox.cu', 'bindings.cpp']
    if not all(os.path.isfile(os.path.join(src, f)) for f in files):
        return _MissingBackend()
    from torch.utils.cpp_extension import load
    return load(name='_pvcnn_backend', extra_cflags=['-O3', '-std=c++17'], sources=[os.path.join(src, f) for f in files])


_backend = _load()
