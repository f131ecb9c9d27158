"""deftet_b200 -- Blackwell-native (sm_100a) engine for DefTet's per-tetrahedron hot path.

The CUDA kernels live in ``csrc/`` and are reached only through the C ABI of ``libdeftet_b200.so``
(``include/deftet_b200.h``); this package is the Python host side that mirrors the reference's
``autograd.Function`` / extension-module surface.  Importing the package does not touch CUDA; the first
kernel call loads the library and fails loudly if it is missing.
"""
from . import grid  # noqa: F401

__version__ = "0.1.0"
