// Device-wide primitives shared by the binned search kernels and the adjacency builders:
// exclusive scan (uint32), LSD radix sort (64-bit keys + 32-bit payload), order-preserving float encoding.
#pragma once
#include "common.cuh"

namespace dtb {

// ---- order-preserving float <-> uint (for atomicMin/Max bounding boxes) ---------------------------
__host__ __device__ __forceinline__ unsigned f2ord(float f) {
#ifdef __CUDA_ARCH__
    unsigned u = __float_as_uint(f);
#else
    unsigned u; memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---- exclusive scan --------------------------------------------------------------------------------
// out[i] = sum_{j<i} in[j]; out may alias in.  If total != nullptr the grand total is written there.
// Workspace: scan_workspace_bytes(n).  Three launches per level (reduce / scan block sums / downsweep).
size_t scan_workspace_bytes(size_t n);
int exclusive_scan_u32(const unsigned* in, unsigned* out, size_t n, unsigned* total, void* ws, size_t ws_bytes, cudaStream_t st);
// same, additionally writing a copy of the result to out2 (saves the cell_end = cell_start memcpy of the counting sorts)
int exclusive_scan_u32_dup(const unsigned* in, unsigned* out, unsigned* out2, size_t n, unsigned* total, void* ws, size_t ws_bytes,
                           cudaStream_t st);

// ---- radix sort ------------------------------------------------------------------------------------
// Stable LSD sort of (key64, val32) pairs on key bits [0, key_bits).  Result ends in keys_out/vals_out.
// Workspace: sort_workspace_bytes(n).  keys_in/vals_in are clobbered (used as ping-pong buffers).
size_t sort_workspace_bytes(size_t n);
int radix_sort_pairs_u64(unsigned long long* keys_in, unsigned* vals_in, unsigned long long* keys_out, unsigned* vals_out,
                         size_t n, int key_bits, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace dtb
