"""ctypes binding of ``libdeftet_b200.so`` (the C ABI declared in ``include/deftet_b200.h``).

There is deliberately no fallback: if the shared library is missing or a call is made without a CUDA
device the product path raises.  ``oracle/`` is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEFTET_B200_LIB") or os.path.join(_HERE, "libdeftet_b200.so")      # override: A/B builds of the same ABI

_lib = None
_lock = threading.Lock()

c_f32p = C.c_void_p
c_i32p = C.c_void_p
c_f64p = C.c_void_p
c_vp = C.c_void_p
c_sz = C.c_size_t
c_int = C.c_int


class DeftetB200Error(RuntimeError):
    pass


def _declare(lib):
    def sig(name, restype, *argtypes):
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = list(argtypes)

    sig("dtb_last_error", C.c_char_p)
    sig("dtb_version", c_int)
    sig("dtb_device_is_sm100", c_int, c_int)
    # energies
    sig("dtb_tet_energies_workspace", c_sz, c_int, c_int, c_int)
    sig("dtb_tet_energies_forward", c_int, c_f32p, c_i32p, c_f32p, c_int, c_int, c_int, c_int, c_f32p, c_f32p, c_f32p,
        c_f64p, c_vp, c_sz, c_vp)
    sig("dtb_tet_energies_backward", c_int, c_f32p, c_i32p, c_f32p, c_int, c_int, c_int, c_int, c_f64p, c_f32p, c_f32p,
        c_f32p, c_f32p, c_vp)
    sig("dtb_tet_energies_backward_v4", c_int, c_f32p, c_i32p, c_f32p, c_int, c_int, c_int, c_int, c_f64p, c_f32p, c_f32p,
        c_f32p, c_f32p, c_vp)
    sig("dtb_tet_energies_forward_soup", c_int, c_f32p, c_f32p, c_int, c_int, c_int, c_f32p, c_f32p, c_f32p, c_f64p, c_vp,
        c_sz, c_vp)
    sig("dtb_tet_energies_backward_soup", c_int, c_f32p, c_f32p, c_int, c_int, c_int, c_f64p, c_f32p, c_f32p, c_f32p,
        c_f32p, c_vp)
    sig("dtb_tet_inverse_v", c_int, c_f32p, c_i32p, c_int, c_int, c_f32p, c_vp)
    sig("dtb_tet_tiles_bytes", c_sz, c_int)
    sig("dtb_tet_tiles_build", c_int, c_i32p, c_int, c_int, c_vp, c_sz, c_i32p, c_vp)
    sig("dtb_tet_energies_forward_tiled", c_int, c_f32p, c_i32p, c_f32p, c_vp, c_int, c_int, c_int, c_int, c_int, c_f32p, c_f32p,
        c_f32p, c_f64p, c_vp)
    sig("dtb_tet_energies_backward_tiled", c_int, c_f32p, c_f32p, c_vp, c_int, c_int, c_int, c_int, c_int, c_f64p, c_f32p, c_f32p,
        c_f32p, c_f32p, c_vp)
    for name, fn in _EXTRA_SIGS:
        fn(lib, sig)


_EXTRA_SIGS = []


def register_signatures(fn):
    """Other modules of the package append their own prototypes before first use."""
    _EXTRA_SIGS.append((fn.__name__, fn))
    if _lib is not None:            # library already loaded: declare right away
        def sig(name, restype, *argtypes):
            f = getattr(_lib, name)
            f.restype = restype
            f.argtypes = list(argtypes)
        fn(_lib, sig)
    return fn


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise DeftetB200Error(
                        "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(or `make`). deftet_b200 has no CPU fallback." % LIB_PATH)
                l = C.CDLL(LIB_PATH)
                _declare(l)
                _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().dtb_last_error()
        raise DeftetB200Error("%s failed (code %d): %s" % (what or "deftet_b200 call", rc, (msg or b"").decode()))


def aligned(t):
    """Contiguous tensor at a 16-byte aligned address.  The kernels read index / matrix / coordinate arrays with 16-byte vector
    loads and TMA bulk copies; tensors that are views into a packed buffer -- nn.DataParallel hands every replica its parameters
    as slices of ONE coalesced broadcast buffer (train_multigpu.py:105-110,136-140 makes inverse_v such a parameter) -- are only
    element-aligned and get a private copy here."""
    t = t.contiguous()
    return t.clone() if (t.data_ptr() & 15) else t


def ptr(t):
    """Device (or host) address of a torch tensor / None -> NULL."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DeftetB200Error("deftet_b200 kernels need CUDA tensors (got %s); there is no CPU fallback" % t.device)
