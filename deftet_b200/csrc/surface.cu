// Predicted-surface stage of the training step (SURVEY.md rows A9, A3 and the consumer half of A2):
//   boundary_faces   DefTet.get_boundary_index              (reference layers/DefTet/deftet.py:186-195)
//   surface_sample   mesh_utils.sample_surf_point_batch     (reference utils/mesh_utils.py:290-299)
//   chamfer          mesh_utils.point_point_distance + mean (reference utils/mesh_utils.py:360-366, deftet.py:177,180)
// The reference runs these per sample in a Python loop over ragged lists (deftet.py:89-103) with a
// boolean-mask compaction (host sync) per sample.  Here the batch is processed at once in a padded-ragged
// layout: faces (B, Fmax, 3) i32 + counts (B,), no host synchronisation anywhere.
#include "prims.cuh"
#include "deftet_b200.h"

namespace dtb {

// ---- A9 -------------------------------------------------------------------------------------------
// Single pass: a tile of BF_TILE faces per CTA (ticketed per sample, so a tile only ever waits for tiles that already run);
// every thread flags its BF_ITEMS consecutive faces (occupancy of the two tets sums to exactly 1, deftet.py:189), the CTA scans the
// flags, warp 0 publishes the tile's count and sums its predecessors' (chained look-back, 32 tiles per step, the scheme of
// prims.cu's scan), and the flagged faces are written straight to their slots in face-id order -- no flag array, no separate scan
// (round 1: three launches and ~80 MB of traffic for 18 MB of algorithmic bytes).
constexpr int BF_THREADS = 256;
#ifndef BF_ITEMS_N
#define BF_ITEMS_N 4
#endif
constexpr int BF_ITEMS = BF_ITEMS_N;
constexpr int BF_TILE = BF_THREADS * BF_ITEMS;

__global__ void __launch_bounds__(BF_THREADS) bf_fused_kernel(const int32_t* __restrict__ face, const int32_t* __restrict__ face_tet,
                                                              const float* __restrict__ occ, int T, int F, int Fmax, int n_tiles,
                                                              unsigned long long* state, unsigned* ticket, int32_t* __restrict__ out,
                                                              int32_t* __restrict__ counts, int* __restrict__ overflow) {
    __shared__ unsigned s_warp[BF_THREADS / 32];
    __shared__ unsigned s_tile, s_prefix, s_total;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket + b, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* ob = occ + (size_t)b * T;
    const int f0 = (int)tile * BF_TILE + threadIdx.x * BF_ITEMS;
    unsigned bits = 0, flips = 0;
#pragma unroll
    for (int k = 0; k < BF_ITEMS; ++k) {
        const int f = f0 + k;
        if (f < F) {
            const int2 t = reinterpret_cast<const int2*>(face_tet)[f];
            const float o0 = ob[t.x], o1 = ob[t.y];
            if (o0 + o1 == 1.0f) bits |= 1u << k;                         // tet_face_occ_bxf (deftet.py:189)
            if (o0 == 1.0f) flips |= 1u << k;                             // change_idx: tet-0 side occupied -> reversed winding (:191-194)
        }
    }
    const unsigned cnt = __popc(bits);
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned w = lane < BF_THREADS / 32 ? s_warp[lane] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
        if (lane < BF_THREADS / 32) s_warp[lane] = wi - w;                // exclusive prefix of the warp totals
        const unsigned btot = __shfl_sync(0xffffffffu, wi, 31);
        volatile unsigned long long* st = state + (size_t)b * n_tiles;
        unsigned prefix = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = (2ull << 32) | btot;
        } else {
            if (lane == 0) st[tile] = (1ull << 32) | btot;
            __threadfence();
            int hi = (int)tile - 1;
            for (;;) {
                const int p = hi - lane;
                const unsigned long long wv = (p >= 0) ? st[p] : (2ull << 32);
                const unsigned fl = (unsigned)(wv >> 32);
                const unsigned ready = __ballot_sync(0xffffffffu, fl != 0u), inc = __ballot_sync(0xffffffffu, fl == 2u);
                const int first = inc ? (__ffs(inc) - 1) : 32;
                const unsigned need = (first >= 32) ? 0xffffffffu : ((2u << first) - 1u);
                if ((ready & need) != need) continue;
                unsigned val = (lane <= first) ? (unsigned)wv : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                prefix += val;
                if (first < 32) break;
                hi -= 32;
            }
            if (lane == 0) st[tile] = (2ull << 32) | (prefix + btot);
        }
        __threadfence();
        if (lane == 0) { s_prefix = prefix; s_total = prefix + btot; }
    }
    __syncthreads();
    unsigned slot = s_prefix + s_warp[warp] + (incl - cnt);
    if ((int)(tile + 1) * BF_TILE >= F && threadIdx.x == 0) {             // the sample's last tile knows the count
        unsigned n = s_total;
        if (n > (unsigned)Fmax) { atomicExch(overflow, 1); n = (unsigned)Fmax; }
        counts[b] = (int)n;
    }
#pragma unroll
    for (int k = 0; k < BF_ITEMS; ++k) {
        if (!(bits & (1u << k))) continue;
        const unsigned dst = slot++;
        if (dst >= (unsigned)Fmax) continue;
        const int f = f0 + k;
        const int a = face[f * 3], bb = face[f * 3 + 1], c = face[f * 3 + 2];
        const bool flip = (flips >> k) & 1u;
        int32_t* o = out + ((size_t)b * Fmax + dst) * 3;
        o[0] = flip ? c : a; o[1] = bb; o[2] = flip ? a : c;
    }
}

// ---- A3: sample points on the predicted surface ------------------------------------------------------
// q = (1-u) a + u (1-v) b + u v c   with u = sqrt(rand), v = rand supplied by the caller (mesh_utils.py:295-298)
__global__ void __launch_bounds__(256) sample_fwd_kernel(const float* __restrict__ pos, int V, const int32_t* __restrict__ faces,
                                                         const int32_t* __restrict__ counts, int Fmax, int S,
                                                         const float* __restrict__ u, const float* __restrict__ v,
                                                         float* __restrict__ q) {
    int b = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;      // sample index in [0, Fmax*S)
    if (i >= counts[b] * S) return;
    int f = i / S;
    const int32_t* fi = faces + ((size_t)b * Fmax + f) * 3;
    const float* p = pos + (size_t)b * V * 3;
    size_t o = (size_t)b * Fmax * S + i;
    float uu = u[o], vv = v[o];
    float wa = 1.f - uu, wb = uu * (1.f - vv), wc = uu * vv;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        q[o * 3 + k] = wa * p[(size_t)fi[0] * 3 + k] + wb * p[(size_t)fi[1] * 3 + k] + wc * p[(size_t)fi[2] * 3 + k];
}

// chamfer_b = mean_i sqrt(|q_i - p_nn(i)|^2 + 1e-10)     (mesh_utils.py:364-365, deftet.py:180)
__global__ void __launch_bounds__(256) chamfer_fwd_kernel(const float* __restrict__ q, const int32_t* __restrict__ nn,
                                                          const float* __restrict__ gt, int M, const int32_t* __restrict__ counts,
                                                          int Qmax, int S, double* __restrict__ acc) {
    int b = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = counts[b] * S;
    double d = 0.0;
    if (i < n) {
        size_t o = (size_t)b * Qmax + i;
        const float* p = gt + ((size_t)b * M + nn[o]) * 3;
        float dx = q[o * 3] - p[0], dy = q[o * 3 + 1] - p[1], dz = q[o * 3 + 2] - p[2];
        d = (double)sqrtf(dx * dx + dy * dy + dz * dz + 1e-10f);
    }
    d = warp_sum(d);
    __shared__ double sw[8];
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += sw[w];
        if (s != 0.0) atomicAdd(acc + b, s);
    }
}
__global__ void chamfer_finalize_kernel(const double* __restrict__ acc, const int32_t* __restrict__ counts, int S, int B,
                                        float* __restrict__ loss) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int n = counts[b] * S;
    loss[b] = n > 0 ? (float)(acc[b] / (double)n) : 1.0f;     // empty surface -> constant 1 (deftet.py:162-166)
}

// backward of chamfer + sampling fused: d loss_b / d q_i = g_b / n_b * (q_i - p_i) / dist_i, scattered to the
// three face vertices with the sampling weights.
__global__ void __launch_bounds__(256) chamfer_bwd_kernel(const float* __restrict__ q, const int32_t* __restrict__ nn,
                                                          const float* __restrict__ gt, int M, int V, const int32_t* __restrict__ faces,
                                                          const int32_t* __restrict__ counts, int Fmax, int S,
                                                          const float* __restrict__ u, const float* __restrict__ v,
                                                          const float* __restrict__ g_loss, float* __restrict__ grad_pos, int gstride) {
    int b = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = counts[b] * S;
    if (i >= n) return;
    size_t o = (size_t)b * Fmax * S + i;
    const float* p = gt + ((size_t)b * M + nn[o]) * 3;
    float dx = q[o * 3] - p[0], dy = q[o * 3 + 1] - p[1], dz = q[o * 3 + 2] - p[2];
    float dist = sqrtf(dx * dx + dy * dy + dz * dz + 1e-10f);
    float s = g_loss[b] / ((float)n * dist);
    float gq[3] = {s * dx, s * dy, s * dz};
    int f = i / S;
    const int32_t* fi = faces + ((size_t)b * Fmax + f) * 3;
    float uu = u[o], vv = v[o];
    float w[3] = {1.f - uu, uu * (1.f - vv), uu * vv};
    float* gp = grad_pos + (size_t)b * V * gstride;
#pragma unroll
    for (int c = 0; c < 3; ++c) grad_add3(gp, (size_t)fi[c], gstride, w[c] * gq[0], w[c] * gq[1], w[c] * gq[2]);
}

// gather the (B, Fmax, 3, 3) vertex soup of the boundary faces (input of A4 / A5)
__global__ void __launch_bounds__(256) face_soup_kernel(const float* __restrict__ pos, int V, const int32_t* __restrict__ faces,
                                                        const int32_t* __restrict__ counts, int Fmax, float* __restrict__ soup) {
    int b = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;      // corner index
    if (i >= counts[b] * 3) return;
    int vid = faces[(size_t)b * Fmax * 3 + i];
    const float* p = pos + ((size_t)b * V + vid) * 3;
    float* o = soup + ((size_t)b * Fmax * 3 + i) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
}

}  // namespace dtb

using namespace dtb;

extern "C" size_t dtb_boundary_faces_workspace(int B, int F) {
    size_t tiles = ((size_t)(F > 0 ? F : 1) + BF_TILE - 1) / BF_TILE;
    return align_up((size_t)B * tiles * sizeof(unsigned long long) + (size_t)B * sizeof(unsigned), 256) + 512;
}

extern "C" int dtb_boundary_faces(const int32_t* face_fx3, const int32_t* face_tet_fx2, const float* occ, int B, int T, int F, int Fmax,
                                  int32_t* out_faces, int32_t* out_counts, int32_t* overflow, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    DTB_REQUIRE(face_fx3 && face_tet_fx2 && occ && out_faces && out_counts && overflow, "boundary_faces: null argument");
    DTB_REQUIRE(B > 0 && F >= 0 && Fmax >= 0, "boundary_faces: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (F == 0) { DTB_CUDA(cudaMemsetAsync(out_counts, 0, B * sizeof(int), st)); return DTB_OK; }
    const int tiles = cdiv(F, BF_TILE);
    Workspace ws(workspace, workspace_bytes);
    const size_t state_bytes = (size_t)B * tiles * sizeof(unsigned long long) + (size_t)B * sizeof(unsigned);
    unsigned long long* state = (unsigned long long*)ws.take<char>(state_bytes);
    if (!ws.ok || !workspace) { set_error("boundary_faces: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    unsigned* ticket = (unsigned*)(state + (size_t)B * tiles);
    DTB_CUDA(cudaMemsetAsync(state, 0, state_bytes, st));
    dim3 grid(tiles, B);
    bf_fused_kernel<<<grid, BF_THREADS, 0, st>>>(face_fx3, face_tet_fx2, occ, T, F, Fmax, tiles, state, ticket, out_faces, out_counts, overflow);
    DTB_LAUNCH_CHECK("bf_fused");
    return DTB_OK;
}

extern "C" int dtb_surface_sample(const float* pos, const int32_t* faces, const int32_t* counts, const float* u, const float* v, int B,
                                  int V, int Fmax, int S, float* q, void* stream) {
    DTB_REQUIRE(pos && faces && counts && u && v && q, "surface_sample: null argument");
    if (B == 0 || Fmax == 0 || S == 0) return DTB_OK;
    dim3 grid(cdiv((long long)Fmax * S, 256), B);
    sample_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pos, V, faces, counts, Fmax, S, u, v, q);
    DTB_LAUNCH_CHECK("sample_fwd");
    return DTB_OK;
}

extern "C" int dtb_chamfer_forward(const float* q, const int32_t* nn, const float* gt, const int32_t* counts, int B, int Fmax, int S,
                                   int M, double* acc, float* loss, void* stream) {
    DTB_REQUIRE(q && nn && gt && counts && acc && loss, "chamfer_forward: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    DTB_CUDA(cudaMemsetAsync(acc, 0, B * sizeof(double), st));
    if (Fmax > 0 && S > 0) {
        dim3 grid(cdiv((long long)Fmax * S, 256), B);
        chamfer_fwd_kernel<<<grid, 256, 0, st>>>(q, nn, gt, M, counts, Fmax * S, S, acc);
        DTB_LAUNCH_CHECK("chamfer_fwd");
    }
    chamfer_finalize_kernel<<<cdiv(B, 64), 64, 0, st>>>(acc, counts, S, B, loss);
    DTB_LAUNCH_CHECK("chamfer_finalize");
    return DTB_OK;
}

extern "C" int dtb_chamfer_backward(const float* q, const int32_t* nn, const float* gt, const int32_t* faces, const int32_t* counts,
                                    const float* u, const float* v, const float* g_loss, int B, int V, int Fmax, int S, int M,
                                    float* grad_pos, int grad_stride, void* stream) {
    DTB_REQUIRE(q && nn && gt && faces && counts && u && v && g_loss && grad_pos, "chamfer_backward: null argument");
    if (B == 0 || Fmax == 0 || S == 0) return DTB_OK;
    dim3 grid(cdiv((long long)Fmax * S, 256), B);
    DTB_REQUIRE(grad_stride == 3 || (grad_stride == 4 && (((size_t)grad_pos) & 15) == 0), "chamfer_backward: bad grad_stride / alignment");
    chamfer_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q, nn, gt, M, V, faces, counts, Fmax, S, u, v, g_loss, grad_pos, grad_stride);
    DTB_LAUNCH_CHECK("chamfer_bwd");
    return DTB_OK;
}

extern "C" int dtb_face_soup(const float* pos, const int32_t* faces, const int32_t* counts, int B, int V, int Fmax, float* soup,
                             void* stream) {
    DTB_REQUIRE(pos && faces && counts && soup, "face_soup: null argument");
    if (B == 0 || Fmax == 0) return DTB_OK;
    dim3 grid(cdiv((long long)Fmax * 3, 256), B);
    face_soup_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pos, V, faces, counts, Fmax, soup);
    DTB_LAUNCH_CHECK("face_soup");
    return DTB_OK;
}
