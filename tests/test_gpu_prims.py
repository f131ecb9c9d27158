"""Hand-written device scan / radix sort against numpy."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from deftet_b200 import _lib as L
    lib = L.lib()
    lib.dtb_prim_scan_workspace.restype = C.c_size_t
    lib.dtb_prim_scan_workspace.argtypes = [C.c_size_t]
    lib.dtb_prim_exclusive_scan_u32.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.dtb_prim_sort_workspace.restype = C.c_size_t
    lib.dtb_prim_sort_workspace.argtypes = [C.c_size_t]
    lib.dtb_prim_radix_sort_pairs_u64.argtypes = [C.c_void_p] * 4 + [C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    return L, lib


@pytest.mark.parametrize("n", [1, 5, 2048, 2049, 100003, 5_000_011])
def test_exclusive_scan(n):
    L, lib = _lib()
    rng = np.random.default_rng(n)
    a = rng.integers(0, 7, size=n, dtype=np.uint32)
    d = torch.from_numpy(a.view(np.int32)).cuda()
    out = torch.empty_like(d)
    tot = torch.zeros(1, dtype=torch.int32, device="cuda")
    wsz = lib.dtb_prim_scan_workspace(n)
    ws = torch.empty(wsz, dtype=torch.uint8, device="cuda")
    L.check(lib.dtb_prim_exclusive_scan_u32(L.ptr(d), L.ptr(out), n, L.ptr(tot), L.ptr(ws), wsz, L.stream_ptr()))
    ref = np.concatenate([[0], np.cumsum(a.astype(np.int64))[:-1]])
    assert np.array_equal(out.cpu().numpy().view(np.uint32).astype(np.int64), ref)
    assert int(tot.item()) == int(a.sum())


@pytest.mark.parametrize("n,bits", [(1, 8), (777, 16), (2048 * 3 + 5, 40), (300001, 64), (1_000_003, 24)])
def test_radix_sort_stable(n, bits):
    L, lib = _lib()
    rng = np.random.default_rng(n)
    hi = (1 << bits) - 1 if bits < 64 else (1 << 63) - 1
    keys = rng.integers(0, min(hi, 1 << 62), size=n, dtype=np.int64)
    if n > 100:
        keys[::3] = keys[0]                      # many duplicates: stability matters
    vals = np.arange(n, dtype=np.int32)
    dk = torch.from_numpy(keys).cuda()
    dv = torch.from_numpy(vals).cuda()
    ok = torch.empty_like(dk)
    ov = torch.empty_like(dv)
    wsz = lib.dtb_prim_sort_workspace(n)
    ws = torch.empty(wsz, dtype=torch.uint8, device="cuda")
    L.check(lib.dtb_prim_radix_sort_pairs_u64(L.ptr(dk), L.ptr(dv), L.ptr(ok), L.ptr(ov), n, bits, L.ptr(ws), wsz, L.stream_ptr()))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(ok.cpu().numpy(), keys[order])
    assert np.array_equal(ov.cpu().numpy(), vals[order])
