// Hand-written device-wide scan and radix sort (no CUB/Thrust on the product path).
#include "prims.cuh"
#include <string.h>

namespace dtb {

// =====================================================================================================
// exclusive scan: tiles of SCAN_TILE = 512 threads x 4 items; `out` may alias `in` (a tile reads its items before writing them)
// =====================================================================================================
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* s_warp, unsigned& block_total) {
    // inclusive warp scan
    unsigned x = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned w = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0u;
        unsigned xi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) xi += y;
        }
        s_warp[lane] = xi - w;               // exclusive prefix of warp totals
        if (lane == 31) s_warp[32] = xi;     // block total
    }
    __syncthreads();
    block_total = s_warp[32];
    return x - v + s_warp[warp];
}

// scans one tile per block; adds block_offsets[blockIdx.x] when given
__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(const unsigned* in, unsigned* out, size_t n,
                                                                 const unsigned* __restrict__ block_offsets, unsigned* __restrict__ total,
                                                                 unsigned* out2) {
    __shared__ unsigned s_warp[33];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS];
    unsigned tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        tsum += v[k];
    }
    unsigned btot;
    unsigned ex = block_exclusive_scan(tsum, s_warp, btot);
    unsigned off = block_offsets ? block_offsets[blockIdx.x] : 0u;
    ex += off;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) { out[base + k] = ex; if (out2) out2[base + k] = ex; }
        ex += v[k];
    }
    if (total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total = off + btot;
}

// Single-pass chained scan with decoupled look-back: tiles take a ticket (scheduling order, so a tile only ever waits
// for tiles that are already running), publish (flag, value) as ONE 64-bit word -- flag 1 = tile aggregate, 2 = inclusive
// prefix -- and warp 0 sums the predecessors' words backwards, 32 at a time, until it meets an inclusive prefix.
__global__ void __launch_bounds__(SCAN_THREADS) scan_onepass_kernel(const unsigned* in, unsigned* out,
                                                                   unsigned* out2, size_t n, unsigned long long* state,
                                                                   unsigned* ticket, unsigned* __restrict__ total) {
    __shared__ unsigned s_warp[33];
    __shared__ unsigned s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    size_t base = (size_t)tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS];
    unsigned tsum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        tsum += v[k];
    }
    unsigned btot;
    unsigned ex = block_exclusive_scan(tsum, s_warp, btot);
    if (threadIdx.x < 32) {                       // warp 0 publishes and looks back, 32 predecessors per step
        volatile unsigned long long* st = state;
        const int lane = threadIdx.x;
        unsigned prefix = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = (2ull << 32) | btot;
        } else {
            if (lane == 0) { st[tile] = (1ull << 32) | btot; }
            __threadfence();
            int hi = (int)tile - 1;               // window = tiles hi, hi-1, ..., hi-31
            for (;;) {
                int p = hi - lane;
                unsigned long long w = (p >= 0) ? st[p] : (2ull << 32);      // before tile 0: an inclusive prefix of 0
                unsigned flag = (unsigned)(w >> 32);
                unsigned ready = __ballot_sync(0xffffffffu, flag != 0u);
                unsigned incl = __ballot_sync(0xffffffffu, flag == 2u);
                // lanes 0..first-1 must all be published, where first = nearest inclusive prefix in the window (or 32)
                int first = incl ? (__ffs(incl) - 1) : 32;
                unsigned need = (first >= 32) ? 0xffffffffu : ((2u << first) - 1u);
                if ((ready & need) != need) continue;                       // somebody in front has not published yet: re-read
                unsigned val = (lane <= first) ? (unsigned)w : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                prefix += val;
                if (first < 32) break;
                hi -= 32;
            }
            if (lane == 0) st[tile] = (2ull << 32) | (prefix + btot);
        }
        __threadfence();
        if (lane == 0) {
            s_prefix = prefix;
            if (total && (size_t)(tile + 1) * SCAN_TILE >= n) *total = prefix + btot;
        }
    }
    __syncthreads();
    ex += s_prefix;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) { out[base + k] = ex; if (out2) out2[base + k] = ex; }
        ex += v[k];
    }
}

size_t scan_workspace_bytes(size_t n) {
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    return align_up((nb + 2) * sizeof(unsigned long long), 256) + 256;
}

int exclusive_scan_u32(const unsigned* in, unsigned* out, size_t n, unsigned* total, void* ws, size_t ws_bytes, cudaStream_t st) {
    return exclusive_scan_u32_dup(in, out, nullptr, n, total, ws, ws_bytes, st);
}

int exclusive_scan_u32_dup(const unsigned* in, unsigned* out, unsigned* out2, size_t n, unsigned* total, void* ws, size_t ws_bytes,
                           cudaStream_t st) {
    if (n == 0) {
        if (total) DTB_CUDA(cudaMemsetAsync(total, 0, sizeof(unsigned), st));
        return DTB_OK;
    }
    if (n <= (size_t)SCAN_TILE) {
        scan_down_kernel<<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr, total, out2);
        DTB_LAUNCH_CHECK("scan_down");
        return DTB_OK;
    }
    size_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    size_t need = align_up((nb + 2) * sizeof(unsigned long long), 256);
    if (ws_bytes < need || !ws) { set_error("exclusive_scan: workspace too small"); return DTB_EWORKSPACE; }
    if (nb >= (1ull << 31)) { set_error("exclusive_scan: too many tiles"); return DTB_EOVERFLOW; }
    unsigned long long* state = (unsigned long long*)ws;
    unsigned* ticket = (unsigned*)(state + nb);
    DTB_CUDA(cudaMemsetAsync(ws, 0, (nb + 1) * sizeof(unsigned long long), st));
    scan_onepass_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, out2, n, state, ticket, total);
    DTB_LAUNCH_CHECK("scan_onepass");
    return DTB_OK;
}

// =====================================================================================================
// LSD radix sort, 8 bits per pass.  Per pass: (1) per-block digit histogram, (2) scan of the
// digit-major [256][blocks] table, (3) stable scatter (warp-level match ranking inside the block).
// =====================================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // keys per block; items are blocked per warp so that order is preserved
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const unsigned long long* __restrict__ keys, size_t n, int shift,
                                                             unsigned* __restrict__ table, unsigned nblocks) {
    __shared__ unsigned h[RS_RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        size_t i = base + (size_t)k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    table[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter.  Each warp owns a contiguous run of 32*RS_ITEMS keys of the tile (warp w: items
// [w*256, (w+1)*256)), processed 32 at a time in order; ranks come from __match_any_sync.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals,
                                                                unsigned long long* __restrict__ keys_out, unsigned* __restrict__ vals_out,
                                                                size_t n, int shift, const unsigned* __restrict__ table, unsigned nblocks) {
    constexpr int WARPS = RS_THREADS / 32;
    __shared__ unsigned s_cnt[WARPS][RS_RADIX];     // per-warp digit counts -> exclusive offsets across warps
    __shared__ unsigned s_base[RS_RADIX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < WARPS * RS_RADIX; i += RS_THREADS) (&s_cnt[0][0])[i] = 0;
    s_base[threadIdx.x] = table[(size_t)threadIdx.x * nblocks + blockIdx.x];
    __syncthreads();
    size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)warp * (32 * RS_ITEMS);
    unsigned long long k[RS_ITEMS];
    unsigned digit[RS_ITEMS];
    unsigned rank[RS_ITEMS];
    // pass A: count digits per warp, remember each key's rank among equal digits seen so far in this warp
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        size_t i = wbase + (size_t)it * 32 + lane;
        bool valid = i < n;
        k[it] = valid ? keys[i] : ~0ull;
        digit[it] = valid ? ((unsigned)(k[it] >> shift) & 0xffu) : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, digit[it]);
        unsigned before = __popc(peers & ((1u << lane) - 1u));
        unsigned prev = valid ? s_cnt[warp][digit[it]] : 0u;
        rank[it] = prev + before;
        __syncwarp();
        if (valid && before == 0) s_cnt[warp][digit[it]] = prev + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive scan of the per-warp counts over warps, per digit (thread d handles digit d)
    {
        unsigned run = s_base[threadIdx.x];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            unsigned c = s_cnt[w][threadIdx.x];
            s_cnt[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        size_t i = wbase + (size_t)it * 32 + lane;
        if (i < n) {
            unsigned dst = s_cnt[warp][digit[it]] + rank[it];
            keys_out[dst] = k[it];
            vals_out[dst] = vals[i];
        }
    }
}

size_t sort_workspace_bytes(size_t n) {
    size_t nb = (n + RS_TILE - 1) / RS_TILE;
    size_t tbl = (size_t)RS_RADIX * (nb ? nb : 1);
    return align_up(tbl * sizeof(unsigned), 256) + scan_workspace_bytes(tbl) + 256;
}

int radix_sort_pairs_u64(unsigned long long* keys_in, unsigned* vals_in, unsigned long long* keys_out, unsigned* vals_out,
                         size_t n, int key_bits, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (n == 0) return DTB_OK;
    if (n >= (1ull << 32)) { set_error("radix_sort: n too large"); return DTB_EOVERFLOW; }
    int passes = (key_bits + 7) / 8;
    if (passes < 1) passes = 1;
    size_t nb = (n + RS_TILE - 1) / RS_TILE;
    size_t tbl = (size_t)RS_RADIX * nb;
    size_t tbl_bytes = align_up(tbl * sizeof(unsigned), 256);
    if (ws_bytes < sort_workspace_bytes(n) || !ws) { set_error("radix_sort: workspace too small"); return DTB_EWORKSPACE; }
    unsigned* table = (unsigned*)ws;
    void* scan_ws = (char*)ws + tbl_bytes;
    size_t scan_ws_bytes = ws_bytes - tbl_bytes;
    unsigned long long* ka = keys_in; unsigned* va = vals_in;
    unsigned long long* kb = keys_out; unsigned* vb = vals_out;
    // make the final pass land in *_out: with an even number of passes start by copying in -> out
    if (passes % 2 == 0) {
        DTB_CUDA(cudaMemcpyAsync(keys_out, keys_in, n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        DTB_CUDA(cudaMemcpyAsync(vals_out, vals_in, n * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
        ka = keys_out; va = vals_out; kb = keys_in; vb = vals_in;
    }
    for (int p = 0; p < passes; ++p) {
        int shift = p * 8;
        rs_hist_kernel<<<(unsigned)nb, RS_THREADS, 0, st>>>(ka, n, shift, table, (unsigned)nb);
        DTB_LAUNCH_CHECK("rs_hist");
        int rc = exclusive_scan_u32(table, table, tbl, nullptr, scan_ws, scan_ws_bytes, st);
        if (rc) return rc;
        rs_scatter_kernel<<<(unsigned)nb, RS_THREADS, 0, st>>>(ka, va, kb, vb, n, shift, table, (unsigned)nb);
        DTB_LAUNCH_CHECK("rs_scatter");
        unsigned long long* tk = ka; ka = kb; kb = tk;
        unsigned* tv = va; va = vb; vb = tv;
    }
    return DTB_OK;
}

}  // namespace dtb
