// Tile-local form of the fused per-tetrahedron energies (A6 AMIPS, A7 volume variance, A8 edge length; forward + backward).
//
// Same math as energies.cu (energies_math.cuh; reference layers/DefTet/deftet.py:239-338), different data movement.  The
// direct-gather kernels of energies.cu are bound by the load/store pipe: 12 scattered 4-byte loads per tet-sample forward and
// 4 vector reductions per tet-sample backward (ncu r1: 7.7 M REDs, issue slots 54 % busy, DRAM 5 %).  Consecutive tets of a grid
// share vertices -- 256 tets of the res-70 lattice (and of the shipped QuarTet grids) touch ~120 distinct vertices, not 1024 --
// so the topology is re-encoded ONCE per grid (dtb_tet_tiles_build) into tiles of 256 tets with
//   vid      the tile's distinct vertex ids, ascending                      (local -> global)
//   corner   per (tet, corner) slot the LOCAL vertex id, 2 bytes
//   inc_*    per local vertex the list of slots that reference it         (gather form of the scatter)
// and a CTA then
//   forward   stages corner ids / inverse rest matrices / vertex ids with 1-D TMA bulk copies, and per sample copies the tile's
//             ~120 vertices into shared memory with cp.async (one coalesced-ish pass instead of 1024 scattered loads), double
//             buffered against the math; per-sample sums are reduced in fp32 inside the warp, in fp64 across warps and CTAs;
//             the volume variance needs no second pass: power sums of (V - c) about a per-sample pilot value c (the volume of
//             tet 0) are re-centred to the mean in fp64 by the last CTA, which also writes the outputs (no finalize launch,
//             no 4-byte-per-tet-sample volume round trip);
//   backward  writes the 12 gradient components of each tet to shared memory, then ONE thread per local vertex sums its
//             incident slots and issues ONE 16-byte vector reduction per (vertex, tile, sample): ~8.7x fewer global REDs.
// Grid = (tiles, sample groups): many more CTAs than SM slots, so the hardware scheduler balances the tail.
//
// Algorithmic HBM bytes per step are unchanged (topology once, positions once, gradient once); what changes is L1/L2 traffic
// and the instruction count.
#include <stdlib.h>
#include "energies_math.cuh"
#include "deftet_b200.h"

namespace dtb {

constexpr int TT = 256;             // tets per tile
constexpr int TCAP = 4 * TT;        // worst-case distinct vertices per tile (= slots per tile)
constexpr int TOFF = TCAP + 16;     // inc_off entries per tile (n_loc + 1 used)

// byte offsets of the arrays inside the tile buffer
struct TileLayout {
    size_t vid, corner, inc_off, inc_slot, nloc, total;
    __host__ __device__ explicit TileLayout(int n_tiles) {
        size_t o = 0;
        vid = o;      o += align_up_c((size_t)n_tiles * TCAP * 4);
        corner = o;   o += align_up_c((size_t)n_tiles * TCAP * 2);
        inc_off = o;  o += align_up_c((size_t)n_tiles * TOFF * 2);
        inc_slot = o; o += align_up_c((size_t)n_tiles * TCAP * 2);
        nloc = o;     o += align_up_c((size_t)n_tiles * 4);
        total = o;
    }
    __host__ __device__ static size_t align_up_c(size_t x) { return (x + 255) / 256 * 256; }
};

// ---- builder: one CTA per tile, bitonic sort of the 1024 (vertex id, slot) keys -------------------------------------------
__global__ void __launch_bounds__(256) tet_tiles_build_kernel(const int32_t* __restrict__ tet, int T, unsigned char* __restrict__ buf,
                                                              int n_tiles, int* __restrict__ nloc_max) {
    __shared__ unsigned long long key[TCAP];
    __shared__ int head_scan[TCAP];
    __shared__ int warp_tot[8];
    const TileLayout L(n_tiles);
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int t0 = tile * TT;
    const int n = min(TT, T - t0);
    int32_t* vid = reinterpret_cast<int32_t*>(buf + L.vid) + (size_t)tile * TCAP;
    uint16_t* corner = reinterpret_cast<uint16_t*>(buf + L.corner) + (size_t)tile * TCAP;
    uint16_t* inc_off = reinterpret_cast<uint16_t*>(buf + L.inc_off) + (size_t)tile * TOFF;
    uint16_t* inc_slot = reinterpret_cast<uint16_t*>(buf + L.inc_slot) + (size_t)tile * TCAP;
    for (int s = tid; s < TCAP; s += 256) {
        int tl = s >> 2;
        key[s] = (tl < n) ? (((unsigned long long)(unsigned)tet[(size_t)(t0 + tl) * 4 + (s & 3)] << 10) | (unsigned)s) : ~0ull;
        corner[s] = 0;
        vid[s] = 0;
        inc_slot[s] = 0;
    }
    for (int s = tid; s < TOFF; s += 256) inc_off[s] = 0;
    __syncthreads();
    for (int k = 2; k <= TCAP; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < TCAP; i += 256) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = key[i], b = key[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { key[i] = b; key[ixj] = a; }
                }
            }
            __syncthreads();
        }
    // heads of runs of equal vertex id -> local ids by an inclusive scan (4 consecutive elements per thread)
    int h[4], run = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int p = tid * 4 + e;
        unsigned long long kp = key[p];
        bool valid = kp != ~0ull;
        bool head = valid && (p == 0 || (key[p - 1] >> 10) != (kp >> 10));
        run += head ? 1 : 0;
        h[e] = run;
    }
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += y;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    int before = incl - run;
    for (int w = 0; w < (tid >> 5); ++w) before += warp_tot[w];
#pragma unroll
    for (int e = 0; e < 4; ++e) head_scan[tid * 4 + e] = before + h[e];
    __syncthreads();
    const int n_loc = head_scan[TCAP - 1];
    for (int p = tid; p < TCAP; p += 256) {
        unsigned long long kp = key[p];
        if (kp == ~0ull) continue;
        int loc = head_scan[p] - 1;
        int slot = (int)(kp & 1023u);
        bool head = (p == 0) || (key[p - 1] >> 10) != (kp >> 10);
        if (head) { vid[loc] = (int32_t)(kp >> 10); inc_off[loc] = (uint16_t)p; }
        corner[slot] = (uint16_t)loc;
        inc_slot[p] = (uint16_t)slot;
    }
    if (tid == 0) {
        inc_off[n_loc] = (uint16_t)(4 * n);
        reinterpret_cast<int32_t*>(buf + L.nloc)[tile] = n_loc;
        atomicMax(nloc_max, n_loc);
    }
}

// ---- shared-memory carve (same for forward and backward) -------------------------------------------------------------------
struct TileSmem {
    float* inv;            // [TT*9]
    uint16_t* corner;      // [TCAP]
    int32_t* vid;          // [nstage]
    float4* pos;           // [2][nstage]
    uint16_t* inc_off;     // [nstage + 8]          (backward)
    uint16_t* inc_slot;    // [TCAP]                (backward)
    float4* g;             // [TCAP]                (backward)
    float* part;           // [8][SG][6]            (forward)
    float* shift;          // [SG]                  (forward)
    uint64_t* bar;
};
__host__ __device__ inline size_t tile_smem_carve(unsigned char* base, int nstage, int sg, bool backward, TileSmem* s) {
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 127) / 128 * 128; return base ? base + at : (unsigned char*)nullptr; };
    unsigned char* p;
    p = take((size_t)TT * 9 * 4);        if (s) s->inv = (float*)p;
    p = take((size_t)TCAP * 2);          if (s) s->corner = (uint16_t*)p;
    p = take((size_t)nstage * 4);        if (s) s->vid = (int32_t*)p;
    p = take((size_t)2 * nstage * 16);   if (s) s->pos = (float4*)p;
    if (backward) {
        p = take((size_t)(nstage + 8) * 2);  if (s) s->inc_off = (uint16_t*)p;
        p = take((size_t)TCAP * 2);          if (s) s->inc_slot = (uint16_t*)p;
        p = take((size_t)TCAP * 16);         if (s) s->g = (float4*)p;
    } else {
        p = take((size_t)8 * sg * 6 * 4);    if (s) s->part = (float*)p;
        p = take((size_t)sg * 4);            if (s) s->shift = (float*)p;
    }
    p = take(16);                        if (s) s->bar = (uint64_t*)p;
    return o;
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// copy the tile's vertices of sample b into s_pos (float4 slots, w unused)
__device__ __forceinline__ void stage_vertices(const float* __restrict__ pos, int V, int b, const int32_t* s_vid, int n_loc, float4* dst) {
    const float* pb = pos + (size_t)b * V * 3;
    for (int i = threadIdx.x; i < n_loc; i += blockDim.x) {
        const float* src = pb + (size_t)s_vid[i] * 3;
        float* d = reinterpret_cast<float*>(dst + i);
        cp_async4(d, src); cp_async4(d + 1, src + 1); cp_async4(d + 2, src + 2);
    }
    cp_async_commit();
}

// TMA-stage the topology of one tile (full tiles: bulk copies; the ragged last tile / unaligned inverse matrices: plain loads)
__device__ __forceinline__ void stage_topology(const TileSmem& S, const unsigned char* __restrict__ buf, const TileLayout& L, int tile, int n,
                                               int nstage, const float* __restrict__ inv_v, bool want_inv, bool inv_aligned, bool backward) {
    const int tid = threadIdx.x;
    const int32_t* g_vid = reinterpret_cast<const int32_t*>(buf + L.vid) + (size_t)tile * TCAP;
    const uint16_t* g_corner = reinterpret_cast<const uint16_t*>(buf + L.corner) + (size_t)tile * TCAP;
    const uint16_t* g_off = reinterpret_cast<const uint16_t*>(buf + L.inc_off) + (size_t)tile * TOFF;
    const uint16_t* g_slot = reinterpret_cast<const uint16_t*>(buf + L.inc_slot) + (size_t)tile * TCAP;
    const bool inv_tma = want_inv && inv_aligned && n == TT;
    if (tid == 0) {
        unsigned bytes = TCAP * 2u + (unsigned)nstage * 4u + (inv_tma ? TT * 36u : 0u) + (backward ? ((unsigned)nstage + 8u) * 2u + TCAP * 2u : 0u);
        mbar_expect_tx(S.bar, bytes);
        tma_load_1d(S.corner, g_corner, TCAP * 2u, S.bar);
        tma_load_1d(S.vid, g_vid, (unsigned)nstage * 4u, S.bar);
        if (inv_tma) tma_load_1d(S.inv, inv_v + (size_t)tile * TT * 9, TT * 36u, S.bar);
        if (backward) {
            tma_load_1d(S.inc_off, g_off, ((unsigned)nstage + 8u) * 2u, S.bar);
            tma_load_1d(S.inc_slot, g_slot, TCAP * 2u, S.bar);
        }
    }
    if (want_inv && !inv_tma)
        for (int i = tid; i < n * 9; i += blockDim.x) S.inv[i] = inv_v[(size_t)tile * TT * 9 + i];
    mbar_wait(S.bar, 0);
    __syncthreads();
}

__device__ __forceinline__ void load_tet(const float4* sp, const ushort4& c, Tet12& t) {
    float4 A = sp[c.x], B = sp[c.y], C = sp[c.z], D = sp[c.w];
    t.a[0] = A.x; t.a[1] = A.y; t.a[2] = A.z; t.b[0] = B.x; t.b[1] = B.y; t.b[2] = B.z;
    t.c[0] = C.x; t.c[1] = C.y; t.c[2] = C.z; t.d[0] = D.x; t.d[1] = D.y; t.d[2] = D.z;
}

// ---- forward ----------------------------------------------------------------------------------------------------------------
// stats (double, per sample, 8 slots): during the kernel 0 amips_sum 1 edge_sum 2..5 power sums S1..S4 of (V - c);
// after the last CTA: 3 = m4, 4 = m3 (centred), 5 = mean volume as fp32 -- the layout energies_bwd kernels read;
// slot 6 = c, slot 7 of sample 0 = CTA ticket counter (zeroed by the host memset, reset by the last CTA).
template <int SG>
__global__ void __launch_bounds__(TT) energies_tiled_fwd_kernel(const float* __restrict__ pos, const int32_t* __restrict__ tet,
                                                                const float* __restrict__ inv_v, const unsigned char* __restrict__ buf,
                                                                int n_tiles, int nstage, int B, int V, int T, int flags, int inv_aligned,
                                                                double* __restrict__ stats, float* __restrict__ amips,
                                                                float* __restrict__ edge, float* __restrict__ volvar) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_last;
    TileSmem S;
    tile_smem_carve(smem_raw, nstage, SG, false, &S);
    const TileLayout L(n_tiles);
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int b0 = blockIdx.y * SG;
    const int nb = min(SG, B - b0);
    const int n = min(TT, T - tile * TT);
    if (tid == 0) { mbar_init(S.bar, 1); mbar_fence_init(); }
    __syncthreads();
    const bool want_inv = (flags & DTB_ENERGY_AMIPS) != 0;
    // pilot volume of each sample of the group (tet 0): identical in every CTA, so the power sums add up across CTAs
    if ((flags & DTB_ENERGY_VOLUME) && tid < nb) {
        int4 id0 = reinterpret_cast<const int4*>(tet)[0];
        const float* p = pos + (size_t)(b0 + tid) * V * 3;
        Tet12 t;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            t.a[k] = p[(size_t)id0.x * 3 + k]; t.b[k] = p[(size_t)id0.y * 3 + k];
            t.c[k] = p[(size_t)id0.z * 3 + k]; t.d[k] = p[(size_t)id0.w * 3 + k];
        }
        float a[3], bb[3], c[3];
        S.shift[tid] = tet_volume(t, a, bb, c);
    }
    stage_topology(S, buf, L, tile, n, nstage, inv_v, want_inv, inv_aligned != 0, false);
    const int n_loc = reinterpret_cast<const int32_t*>(buf + L.nloc)[tile];
    stage_vertices(pos, V, b0, S.vid, n_loc, S.pos);
    const bool active = tid < n;
    ushort4 c = make_ushort4(0, 0, 0, 0);
    float M[9];
    if (active) {
        c = reinterpret_cast<const ushort4*>(S.corner)[tid];
        if (want_inv) {
#pragma unroll
            for (int i = 0; i < 9; ++i) M[i] = S.inv[tid * 9 + i];
        }
    }
    const int lane = tid & 31, warp = tid >> 5;
    for (int s = 0; s < nb; ++s) {
        cp_async_wait_all();
        __syncthreads();
        if (s + 1 < nb) stage_vertices(pos, V, b0 + s + 1, S.vid, n_loc, S.pos + (size_t)((s + 1) & 1) * nstage);
        float e_am = 0.f, e_ed = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f, x4 = 0.f;
        if (active) {
            Tet12 t;
            load_tet(S.pos + (size_t)(s & 1) * nstage, c, t);
            if (flags & DTB_ENERGY_AMIPS) {
                float J[9], det, tr, g;
                e_am = amips_energy(t, M, J, det, tr, g);
            }
            if (flags & DTB_ENERGY_EDGE) e_ed = edge_energy(t);
            if (flags & DTB_ENERGY_VOLUME) {
                float a[3], bb[3], cc[3];
                x1 = tet_volume(t, a, bb, cc) - S.shift[s];
                x2 = x1 * x1; x3 = x2 * x1; x4 = x2 * x2;
            }
        }
        e_am = warp_sum(e_am); e_ed = warp_sum(e_ed);
        x1 = warp_sum(x1); x2 = warp_sum(x2); x3 = warp_sum(x3); x4 = warp_sum(x4);
        if (lane == 0) {
            float* q = S.part + ((size_t)warp * SG + s) * 6;
            q[0] = e_am; q[1] = e_ed; q[2] = x1; q[3] = x2; q[4] = x3; q[5] = x4;
        }
    }
    __syncthreads();
    for (int i = tid; i < nb * 6; i += blockDim.x) {
        int s = i / 6, q = i % 6;
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < TT / 32; ++w) acc += (double)S.part[((size_t)w * SG + s) * 6 + q];
        atomicAdd(&stats[(size_t)(b0 + s) * 8 + q], acc);
    }
    // ---- the last CTA to arrive turns the sums into the outputs (fused finalize) ------------------------------------------
    __threadfence();
    __syncthreads();
    unsigned* counter = reinterpret_cast<unsigned*>(stats + 7);
    if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1u) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int b = tid; b < B; b += blockDim.x) {
        volatile double* a = stats + (size_t)b * 8;
        const double Td = (double)T;
        if ((flags & DTB_ENERGY_AMIPS) && amips) amips[b] = (float)(a[0] / Td);
        if ((flags & DTB_ENERGY_EDGE) && edge) edge[b] = (float)(a[1] / (6.0 * Td));
        if (flags & DTB_ENERGY_VOLUME) {
            int4 id0 = reinterpret_cast<const int4*>(tet)[0];
            const float* p = pos + (size_t)b * V * 3;
            Tet12 t;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                t.a[k] = p[(size_t)id0.x * 3 + k]; t.b[k] = p[(size_t)id0.y * 3 + k];
                t.c[k] = p[(size_t)id0.z * 3 + k]; t.d[k] = p[(size_t)id0.w * 3 + k];
            }
            float ta[3], tb[3], tc[3];
            const double cs = (double)tet_volume(t, ta, tb, tc);
            const double S1 = a[2], S2 = a[3], S3 = a[4], S4 = a[5];
            const double mu = (double)(float)(cs + S1 / Td);              // torch.mean in fp32 (deftet.py:258)
            const double d = mu - cs;
            const double m4 = S4 - 4.0 * d * S3 + 6.0 * d * d * S2 - 4.0 * d * d * d * S1 + Td * d * d * d * d;
            const double m3 = S3 - 3.0 * d * S2 + 3.0 * d * d * S1 - Td * d * d * d;
            if (volvar) volvar[b] = (float)m4;
            a[3] = m4; a[4] = m3; a[5] = mu; a[6] = cs;
        }
    }
    __syncthreads();
    if (tid == 0) *counter = 0u;
}

// ---- backward ---------------------------------------------------------------------------------------------------------------
template <int SG>
__global__ void __launch_bounds__(TT) energies_tiled_bwd_kernel(const float* __restrict__ pos, const float* __restrict__ inv_v,
                                                                const unsigned char* __restrict__ buf, int n_tiles, int nstage, int B, int V,
                                                                int T, int flags, int inv_aligned, const double* __restrict__ stats,
                                                                const float* __restrict__ g_amips, const float* __restrict__ g_edge,
                                                                const float* __restrict__ g_vol, float4* __restrict__ grad4) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileSmem S;
    tile_smem_carve(smem_raw, nstage, SG, true, &S);
    const TileLayout L(n_tiles);
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int b0 = blockIdx.y * SG;
    const int nb = min(SG, B - b0);
    const int n = min(TT, T - tile * TT);
    if (tid == 0) { mbar_init(S.bar, 1); mbar_fence_init(); }
    __syncthreads();
    const bool want_inv = (flags & DTB_ENERGY_AMIPS) != 0 && g_amips != nullptr;
    stage_topology(S, buf, L, tile, n, nstage, inv_v, want_inv, inv_aligned != 0, true);
    const int n_loc = reinterpret_cast<const int32_t*>(buf + L.nloc)[tile];
    stage_vertices(pos, V, b0, S.vid, n_loc, S.pos);
    const bool active = tid < n;
    ushort4 c = make_ushort4(0, 0, 0, 0);
    float M[9];
    if (active) {
        c = reinterpret_cast<const ushort4*>(S.corner)[tid];
        if (want_inv) {
#pragma unroll
            for (int i = 0; i < 9; ++i) M[i] = S.inv[tid * 9 + i];
        }
    }
    const float invT = 1.0f / (float)T;
    for (int s = 0; s < nb; ++s) {
        const int b = b0 + s;
        cp_async_wait_all();
        __syncthreads();                 // vertices of sample s are in place; the gather of sample s-1 has finished reading S.g
        if (s + 1 < nb) stage_vertices(pos, V, b + 1, S.vid, n_loc, S.pos + (size_t)((s + 1) & 1) * nstage);
        if (active) {
            Tet12 t;
            load_tet(S.pos + (size_t)(s & 1) * nstage, c, t);
            float ga[3] = {0, 0, 0}, gb[3] = {0, 0, 0}, gc[3] = {0, 0, 0}, gd[3] = {0, 0, 0};
            if (want_inv) {
                float J[9], det, tr, g;
                amips_energy(t, M, J, det, tr, g);
                amips_grad(M, J, det, tr, g, g_amips[b] * invT, ga, gb, gc, gd);
            }
            if ((flags & DTB_ENERGY_EDGE) && g_edge) edge_grad(t, g_edge[b] * invT * (1.0f / 6.0f), ga, gb, gc, gd);
            if ((flags & DTB_ENERGY_VOLUME) && g_vol) {
                float a[3], bb[3], cc[3];
                float v = tet_volume(t, a, bb, cc);
                float mu = (float)stats[(size_t)b * 8 + 5];
                float s3 = (float)stats[(size_t)b * 8 + 4];
                float d = v - mu;
                float dLdV = 4.f * d * d * d - 4.f * invT * s3;     // mean term is not detached (deftet.py:258-262)
                volume_grad(a, bb, cc, g_vol[b] * dLdV, ga, gb, gc, gd);
            }
            float4* gs = S.g + tid * 4;
            gs[0] = make_float4(ga[0], ga[1], ga[2], 0.f);
            gs[1] = make_float4(gb[0], gb[1], gb[2], 0.f);
            gs[2] = make_float4(gc[0], gc[1], gc[2], 0.f);
            gs[3] = make_float4(gd[0], gd[1], gd[2], 0.f);
        }
        __syncthreads();
        // gather form of the scatter: one thread per local vertex sums its incident slots, one vector reduction per vertex
        for (int i = tid; i < n_loc; i += blockDim.x) {
            const int p0 = S.inc_off[i], p1 = S.inc_off[i + 1];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = p0; p < p1; ++p) {
                float4 x = S.g[S.inc_slot[p]];
                acc.x += x.x; acc.y += x.y; acc.z += x.z;
            }
            atomicAdd(grad4 + (size_t)b * V + S.vid[i], acc);
        }
    }
}

}  // namespace dtb

using namespace dtb;

extern "C" size_t dtb_tet_tiles_bytes(int T) {
    if (T <= 0) return 256;
    return TileLayout(cdiv(T, TT)).total;
}

extern "C" int dtb_tet_tiles_build(const int32_t* tet, int T, int V, void* tiles, size_t tiles_bytes, int32_t* nloc_max, void* stream) {
    DTB_REQUIRE(tet && tiles && nloc_max, "tet_tiles_build: null argument");
    DTB_REQUIRE(T > 0 && V > 0 && V < (1 << 30), "tet_tiles_build: bad sizes T=%d V=%d", T, V);
    DTB_REQUIRE(tiles_bytes >= dtb_tet_tiles_bytes(T), "tet_tiles_build: buffer too small (%zu < %zu)", tiles_bytes, dtb_tet_tiles_bytes(T));
    DTB_REQUIRE((((size_t)tiles) & 255) == 0, "tet_tiles_build: buffer must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    DTB_CUDA(cudaMemsetAsync(nloc_max, 0, sizeof(int32_t), st));
    int n_tiles = cdiv(T, TT);
    tet_tiles_build_kernel<<<n_tiles, 256, 0, st>>>(tet, T, (unsigned char*)tiles, n_tiles, nloc_max);
    DTB_LAUNCH_CHECK("tet_tiles_build");
    return DTB_OK;
}

static int tiled_nstage(int nloc_max) {
    int n = (nloc_max + 7) / 8 * 8;          // multiples of 8 entries: every staged array stays a multiple of 16 bytes
    if (n < 8) n = 8;
    if (n > TCAP) n = TCAP;
    return n;
}

template <int SG>
static int launch_fwd(const float* pos, const int32_t* tet, const float* inv_v, const void* tiles, int nstage, int B, int V, int T, int flags,
                      float* amips, float* edge, float* volvar, double* stats, cudaStream_t st) {
    int n_tiles = cdiv(T, TT);
    size_t smem = tile_smem_carve(nullptr, nstage, SG, false, nullptr);
    auto kern = energies_tiled_fwd_kernel<SG>;
    DTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(n_tiles, cdiv(B, SG));
    prof_begin(PROF_ENERGIES_FWD, st);
    kern<<<grid, TT, smem, st>>>(pos, tet, inv_v, (const unsigned char*)tiles, n_tiles, nstage, B, V, T, flags, (((size_t)inv_v) & 15) == 0 ? 1 : 0,
                                 stats, amips, edge, volvar);
    DTB_LAUNCH_CHECK("energies_tiled_fwd");
    prof_end(PROF_ENERGIES_FWD, st);
    return DTB_OK;
}

template <int SG>
static int launch_bwd(const float* pos, const float* inv_v, const void* tiles, int nstage, int B, int V, int T, int flags, const double* stats,
                      const float* g_amips, const float* g_edge, const float* g_vol, float* grad4, cudaStream_t st) {
    int n_tiles = cdiv(T, TT);
    size_t smem = tile_smem_carve(nullptr, nstage, SG, true, nullptr);
    auto kern = energies_tiled_bwd_kernel<SG>;
    DTB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(n_tiles, cdiv(B, SG));
    prof_begin(PROF_ENERGIES_BWD, st);
    kern<<<grid, TT, smem, st>>>(pos, inv_v, (const unsigned char*)tiles, n_tiles, nstage, B, V, T, flags, (((size_t)inv_v) & 15) == 0 ? 1 : 0, stats,
                                 g_amips, g_edge, g_vol, reinterpret_cast<float4*>(grad4));
    DTB_LAUNCH_CHECK("energies_tiled_bwd");
    prof_end(PROF_ENERGIES_BWD, st);
    return DTB_OK;
}

// samples per CTA: enough CTAs for several waves over the 148 SMs, as few re-stagings of the topology tile as possible
static int pick_group(int n_tiles, int B) {
    const char* e = getenv("DTB_ENERGY_GROUP");
    if (e) { int g = atoi(e); if (g == 1 || g == 2 || g == 4 || g == 8) return g; }
    const long long want = 6LL * 4 * DTB_SM_COUNT;            // >= 6 waves at 4 resident CTAs per SM
    for (int g = 8; g > 1; g >>= 1)
        if ((long long)n_tiles * cdiv(B, g) >= want) return g;
    return 1;
}

extern "C" int dtb_tet_energies_forward_tiled(const float* pos, const int32_t* tet, const float* inv_v, const void* tiles, int nloc_max, int B,
                                              int V, int T, int flags, float* amips, float* edge, float* volvar, double* stats, void* stream) {
    DTB_REQUIRE(pos && tet && tiles && stats, "tet_energies_forward_tiled: null argument");
    DTB_REQUIRE(B > 0 && T > 0, "tet_energies_forward_tiled: empty batch or grid (B=%d T=%d)", B, T);
    DTB_REQUIRE(!(flags & DTB_ENERGY_AMIPS) || inv_v, "tet_energies_forward_tiled: AMIPS requested without inverse_v");
    DTB_REQUIRE(nloc_max > 0 && nloc_max <= TCAP, "tet_energies_forward_tiled: bad nloc_max %d (from dtb_tet_tiles_build)", nloc_max);
    cudaStream_t st = (cudaStream_t)stream;
    DTB_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * 8 * sizeof(double), st));
    const int nstage = tiled_nstage(nloc_max);
    switch (pick_group(cdiv(T, TT), B)) {
        case 8: return launch_fwd<8>(pos, tet, inv_v, tiles, nstage, B, V, T, flags, amips, edge, volvar, stats, st);
        case 4: return launch_fwd<4>(pos, tet, inv_v, tiles, nstage, B, V, T, flags, amips, edge, volvar, stats, st);
        case 2: return launch_fwd<2>(pos, tet, inv_v, tiles, nstage, B, V, T, flags, amips, edge, volvar, stats, st);
        default: return launch_fwd<1>(pos, tet, inv_v, tiles, nstage, B, V, T, flags, amips, edge, volvar, stats, st);
    }
}

extern "C" int dtb_tet_energies_backward_tiled(const float* pos, const float* inv_v, const void* tiles, int nloc_max, int B, int V, int T,
                                               int flags, const double* stats, const float* g_amips, const float* g_edge,
                                               const float* g_volvar, float* grad_pos4, void* stream) {
    DTB_REQUIRE(pos && tiles && stats && grad_pos4, "tet_energies_backward_tiled: null argument");
    DTB_REQUIRE(B > 0 && T > 0, "tet_energies_backward_tiled: empty batch or grid");
    DTB_REQUIRE((((size_t)grad_pos4) & 15) == 0, "tet_energies_backward_tiled: grad_pos4 must be 16-byte aligned");
    DTB_REQUIRE(nloc_max > 0 && nloc_max <= TCAP, "tet_energies_backward_tiled: bad nloc_max %d", nloc_max);
    DTB_REQUIRE(!((flags & DTB_ENERGY_AMIPS) && g_amips) || inv_v, "tet_energies_backward_tiled: AMIPS gradient requested without inverse_v");
    cudaStream_t st = (cudaStream_t)stream;
    const int nstage = tiled_nstage(nloc_max);
    switch (pick_group(cdiv(T, TT), B)) {
        case 8: return launch_bwd<8>(pos, inv_v, tiles, nstage, B, V, T, flags, stats, g_amips, g_edge, g_volvar, grad_pos4, st);
        case 4: return launch_bwd<4>(pos, inv_v, tiles, nstage, B, V, T, flags, stats, g_amips, g_edge, g_volvar, grad_pos4, st);
        case 2: return launch_bwd<2>(pos, inv_v, tiles, nstage, B, V, T, flags, stats, g_amips, g_edge, g_volvar, grad_pos4, st);
        default: return launch_bwd<1>(pos, inv_v, tiles, nstage, B, V, T, flags, stats, g_amips, g_edge, g_volvar, grad_pos4, st);
    }
}
