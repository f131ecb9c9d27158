"""Graph-convolution neighbourhood product on the vertex adjacency.

``sparse_batch_matmul`` keeps the reference's name and arguments (utils/matrix_utils.py:22-33): a torch sparse (n,n) matrix and a
dense (b,n,p) batch.  The matrix is converted once to CSR (and CSR of its transpose, for the backward pass) by
``csrc/graph.cu`` and cached on the sparse tensor's identity; the product itself is one HBM-bound kernel per direction."""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import _lib
from .search import _f32c


@_lib.register_signatures
def _graph_sigs(lib, sig):
    vp, i, sz, ll = C.c_void_p, C.c_int, C.c_size_t, C.c_longlong
    sig("dtb_coo_to_csr_workspace", sz, ll)
    sig("dtb_coo_to_csr", i, vp, vp, vp, ll, i, i, i, vp, vp, vp, vp, sz, vp)
    sig("dtb_spmm_csr", i, vp, vp, vp, vp, i, i, i, i, vp, vp)


class CsrMatrix:
    """CSR of a sparse matrix and of its transpose, on the device (built once per adjacency)."""

    def __init__(self, indices_2xnnz, values_nnz, shape):
        _lib.require_cuda(indices_2xnnz, values_nnz)
        idx = indices_2xnnz.to(torch.int64).contiguous()
        val = _f32c(values_nnz)
        self.shape = (int(shape[0]), int(shape[1]))
        self.nnz = int(val.shape[0])
        self.device = val.device
        self.fwd = self._build(idx, val, 0)
        self.bwd = self._build(idx, val, 1)

    def _build(self, idx, val, transpose):
        n_rows, n_cols = self.shape
        n_major = n_cols if transpose else n_rows
        dev = self.device
        L = _lib.lib()
        row_ptr = torch.empty(n_major + 1, device=dev, dtype=torch.int32)
        col = torch.empty(max(self.nnz, 1), device=dev, dtype=torch.int32)
        out_val = torch.empty(max(self.nnz, 1), device=dev, dtype=torch.float32)
        wsz = L.dtb_coo_to_csr_workspace(self.nnz)
        ws = torch.empty(wsz, device=dev, dtype=torch.uint8)
        rows, cols = idx[0].contiguous(), idx[1].contiguous()
        with torch.cuda.device(dev):
            _lib.check(L.dtb_coo_to_csr(_lib.ptr(rows), _lib.ptr(cols), _lib.ptr(val), self.nnz, n_rows, n_cols, transpose, _lib.ptr(row_ptr),
                                        _lib.ptr(col), _lib.ptr(out_val), _lib.ptr(ws), wsz, _lib.stream_ptr()), "dtb_coo_to_csr")
        return row_ptr, col, out_val

    @classmethod
    def from_sparse(cls, sparse_matrix):
        m = sparse_matrix.coalesce() if not sparse_matrix.is_coalesced() else sparse_matrix
        return cls(m.indices(), m.values(), m.shape)


def _spmm(csr_triplet, n_rows, n_cols, x):
    row_ptr, col, val = csr_triplet
    B, _, p = x.shape
    out = torch.empty(B, n_rows, p, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().dtb_spmm_csr(_lib.ptr(row_ptr), _lib.ptr(col), _lib.ptr(val), _lib.ptr(x), B, n_rows, n_cols, p, _lib.ptr(out),
                                           _lib.stream_ptr()), "dtb_spmm_csr")
    return out


class _SpmmFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, csr):
        x = _f32c(x)
        ctx.csr = csr
        return _spmm(csr.fwd, csr.shape[0], csr.shape[1], x)

    @staticmethod
    def backward(ctx, g):
        csr = ctx.csr
        return _spmm(csr.bwd, csr.shape[1], csr.shape[0], _f32c(g)), None


_CACHE = {}          # id(sparse tensor) -> (weak reference to it, CsrMatrix); an entry is valid only while that very tensor is alive


def csr_of(sparse_matrix):
    """CSR cache keyed by the identity of the torch sparse tensor (the reference keeps one normalised adjacency per device,
    layers/gcn_decoder.py:197-201)."""
    key = id(sparse_matrix)
    hit = _CACHE.get(key)
    if hit is not None and hit[0]() is sparse_matrix:
        return hit[1]
    csr = CsrMatrix.from_sparse(sparse_matrix)
    for k in [k for k, (r, _) in _CACHE.items() if r() is None]:
        del _CACHE[k]
    _CACHE[key] = (weakref.ref(sparse_matrix), csr)
    return csr


def sparse_batch_matmul(sparse_matrix, dense_matrix_batch):
    """``sparse_batch_matmul`` (utils/matrix_utils.py:22-33): (m,n) sparse @ (b,n,p) dense -> (b,m,p); differentiable w.r.t. the dense
    operand (the adjacency is a constant in the reference: MySparse sets requires_grad = False, matrix_utils.py:49-60).
    ``sparse_matrix`` may be a torch sparse COO tensor (converted once, cached) or a ``CsrMatrix``."""
    _lib.require_cuda(dense_matrix_batch)
    csr = sparse_matrix if isinstance(sparse_matrix, CsrMatrix) else csr_of(sparse_matrix)
    assert dense_matrix_batch.shape[1] == csr.shape[1], "inner dimensions differ"
    return _SpmmFunction.apply(dense_matrix_batch, csr)


def adjacency_csr(tet, n_point, normalize=True):
    """CSR of the (row-normalised) vertex adjacency straight from the tet list (A10 edges are already sorted by row)."""
    from .builders import tet_point_adj
    if normalize:
        edges, w = tet_point_adj(tet, n_point, True)
    else:
        edges = tet_point_adj(tet, n_point, False)
        w = torch.ones(edges.shape[0], device=edges.device)
    return CsrMatrix(edges.t().long(), w, (n_point, n_point))
