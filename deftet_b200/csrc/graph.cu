// N2: the graph-convolution neighbourhood product on the vertex adjacency that A10 builds.
//   utils/matrix_utils.py:22-33 sparse_batch_matmul(sparse (n,n), dense (b,n,p)) -> (b,n,p)   [torch.sparse.mm on a reshaped copy]
//   layers/gcn_decoder.py:44-56  GraphConv.forward: filter(x) + filter(A x), 256-wide features
// Here: CSR SpMM, one warp per output row (b, i): the lanes cover the feature dimension with 16-byte loads, the <= ~14 neighbour
// rows of a vertex stream through L2 (mesh numbering keeps them close), every x row is read from HBM once and every output row is
// written once -> HBM bound, algorithmic bytes 2 * 4 * B * n * p (+ 8 nnz + 4 n for the matrix).  No transposed copy of the dense
// operand (the reference materialises (n, b*p) and back).  Backward w.r.t. the dense operand is the same kernel on the transposed
// matrix, built once per adjacency by dtb_coo_to_csr(transpose = 1) -- gather form, no atomics, deterministic.
#include "prims.cuh"
#include "deftet_b200.h"

namespace dtb {

typedef unsigned long long u64;

__global__ void __launch_bounds__(256) coo_keys_kernel(const long long* __restrict__ rows, const long long* __restrict__ cols, long long nnz,
                                                       u64 n_minor, int transpose, u64* __restrict__ keys, unsigned* __restrict__ vals) {
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    u64 r = (u64)(transpose ? cols[e] : rows[e]), c = (u64)(transpose ? rows[e] : cols[e]);
    keys[e] = r * n_minor + c;
    vals[e] = (unsigned)e;
}
// sorted entry p: column + value in place; the first entry of each run of equal rows fills row_ptr for its row and for the empty
// rows before it; the last entry fills the tail
__global__ void __launch_bounds__(256) csr_emit_kernel(const u64* __restrict__ keys, const unsigned* __restrict__ perm, const float* __restrict__ val,
                                                       long long nnz, u64 n_minor, int n_major, int32_t* __restrict__ row_ptr,
                                                       int32_t* __restrict__ col, float* __restrict__ out_val) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    u64 k = keys[p];
    long long r = (long long)(k / n_minor);
    col[p] = (int)(k % n_minor);
    out_val[p] = val[perm[p]];
    long long prev = p == 0 ? -1 : (long long)(keys[p - 1] / n_minor);
    for (long long q = prev + 1; q <= r; ++q) row_ptr[q] = (int)p;
    if (p == nnz - 1) for (long long q = r + 1; q <= n_major; ++q) row_ptr[q] = (int)nnz;
}
__global__ void fill_i32_kernel(int32_t* p, int n, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// VEC = floats per lane and step (4: 16-byte loads, needs p % 4 == 0 and 16-byte aligned rows; 1: any p)
template <int VEC>
__global__ void __launch_bounds__(256) spmm_csr_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                                       const float* __restrict__ val, const float* __restrict__ x, int n_rows, int n_cols,
                                                       int p, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);       // consecutive rows of one sample share a CTA
    const int b = blockIdx.y;
    if (row >= n_rows) return;
    const int e0 = row_ptr[row], e1 = row_ptr[row + 1];
    const float* xb = x + (size_t)b * n_cols * p;
    float* ob = out + ((size_t)b * n_rows + row) * p;
    for (int c0 = lane * VEC; c0 < p; c0 += 32 * VEC) {
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
#pragma unroll 4
        for (int e = e0; e < e1; ++e) {
            const int j = __ldg(col + e);
            const float w = __ldg(val + e);
            const float* src = xb + (size_t)j * p + c0;
            if (VEC == 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]); acc[2] = fmaf(w, v.z, acc[2]); acc[3] = fmaf(w, v.w, acc[3]);
            } else {
                acc[0] = fmaf(w, __ldg(src), acc[0]);
            }
        }
        if (VEC == 4) *reinterpret_cast<float4*>(ob + c0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else ob[c0] = acc[0];
    }
}

}  // namespace dtb

using namespace dtb;

extern "C" size_t dtb_coo_to_csr_workspace(long long nnz) {
    size_t n = (size_t)(nnz > 0 ? nnz : 1);
    Workspace ws(nullptr, 0);
    ws.take<u64>(n); ws.take<u64>(n); ws.take<unsigned>(n); ws.take<unsigned>(n);
    ws.take<char>(sort_workspace_bytes(n));
    return ws.off + 1024;
}
// COO (rows, cols int64 as torch sparse tensors hold them; vals f32; any order, no duplicates expected) -> CSR of the matrix
// (transpose = 0) or of its transpose (transpose = 1): row_ptr (n_major + 1) i32, col (nnz) i32 ascending within a row, val (nnz).
extern "C" int dtb_coo_to_csr(const long long* rows, const long long* cols, const float* vals, long long nnz, int n_rows, int n_cols,
                              int transpose, int32_t* row_ptr, int32_t* col, float* val, void* workspace, size_t workspace_bytes,
                              void* stream) {
    DTB_REQUIRE(row_ptr && n_rows >= 0 && n_cols >= 0 && nnz >= 0, "coo_to_csr: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_major = transpose ? n_cols : n_rows;
    const u64 n_minor = (u64)(transpose ? n_rows : n_cols);
    if (nnz == 0) {
        fill_i32_kernel<<<cdiv(n_major + 1, 256), 256, 0, st>>>(row_ptr, n_major + 1, 0);
        DTB_LAUNCH_CHECK("fill_i32");
        return DTB_OK;
    }
    DTB_REQUIRE(rows && cols && vals && col && val, "coo_to_csr: null argument");
    DTB_REQUIRE(nnz < (1ll << 31), "coo_to_csr: nnz %lld exceeds int32", nnz);
    size_t n = (size_t)nnz;
    Workspace ws(workspace, workspace_bytes);
    u64* k0 = ws.take<u64>(n); u64* k1 = ws.take<u64>(n);
    unsigned* v0 = ws.take<unsigned>(n); unsigned* v1 = ws.take<unsigned>(n);
    size_t sob = sort_workspace_bytes(n);
    void* sows = ws.take<char>(sob);
    if (!ws.ok || !workspace) { set_error("coo_to_csr: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    int blocks = cdiv(nnz, 256);
    coo_keys_kernel<<<blocks, 256, 0, st>>>(rows, cols, nnz, n_minor, transpose, k0, v0);
    DTB_LAUNCH_CHECK("coo_keys");
    int bits = 1;
    u64 mx = (u64)n_major * n_minor;
    while (bits < 64 && (mx >> bits)) ++bits;
    int rc = radix_sort_pairs_u64(k0, v0, k1, v1, n, bits, sows, sob, st);
    if (rc) return rc;
    csr_emit_kernel<<<blocks, 256, 0, st>>>(k1, v1, vals, nnz, n_minor, n_major, row_ptr, col, val);
    DTB_LAUNCH_CHECK("csr_emit");
    return DTB_OK;
}

// out (B, n_rows, p) = A (n_rows, n_cols, CSR) @ x (B, n_cols, p) per sample.  out must not alias x.
extern "C" int dtb_spmm_csr(const int32_t* row_ptr, const int32_t* col, const float* val, const float* x, int B, int n_rows, int n_cols,
                            int p, float* out, void* stream) {
    DTB_REQUIRE(B >= 0 && n_rows >= 0 && n_cols >= 0 && p >= 0, "spmm_csr: bad sizes");
    if ((long long)B * n_rows * p == 0) return DTB_OK;
    DTB_REQUIRE(row_ptr && x && out, "spmm_csr: null argument");
    DTB_REQUIRE(B <= 65535, "spmm_csr: batch %d exceeds the grid limit", B);
    dim3 grid(cdiv(n_rows, 8), B);
    bool vec = (p % 4 == 0) && (((uintptr_t)x | (uintptr_t)out) % 16 == 0);
    if (vec) spmm_csr_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(row_ptr, col, val, x, n_rows, n_cols, p, out);
    else spmm_csr_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(row_ptr, col, val, x, n_rows, n_cols, p, out);
    DTB_LAUNCH_CHECK("spmm_csr");
    return DTB_OK;
}
