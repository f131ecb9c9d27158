// Fused per-tetrahedron energies (AMIPS + edge length + volume variance), forward and backward.
//
// Replaces the ~45 elementwise torch kernels autograd launches for
//   DefTet.amips_energy     (reference layers/DefTet/deftet.py:266-298)
//   DefTet.volume_variance  (layers/DefTet/deftet.py:239-263, pow hard-wired to 4 at :27,:82)
//   DefTet.edge_length      (layers/DefTet/deftet.py:320-338)
//   DefTet.tet_inverse_v / my_inverse (layers/DefTet/deftet.py:205-233,300-318)
// with one forward pass (+ a 4-byte/tet second pass for the centred 4th moment) and one backward pass.
//
// Data layout in HBM: pos (B,V,3) f32, tet (T,4) i32 shared by the batch, inv_v (T,3,3) f32.
// A CTA owns a tile of 256 tets: the index tile (4 KB) and the inverse-rest-matrix tile (9 KB) are
// staged once with 1-D TMA bulk copies (cp.async.bulk + mbarrier) and reused for every sample of the
// batch, so topology bytes cross HBM once per step instead of B times; vertex positions are gathered
// through L2 (12*V bytes per sample, resident).  The "soup" variants take the materialised
// (B,T,4,3) tensor the reference methods receive (deftet.py:66-68) and write a dense gradient.
//
// Algorithmic bytes per step (indexed form): fwd 16T + 36T + 12BV + 4BT(w)+4BT(r); bwd 16T + 36T + 12BV + 12BV.
#include "energies_math.cuh"
#include "deftet_b200.h"

namespace dtb {

// ---- loaders --------------------------------------------------------------------------------------
struct IndexedSrc {     // pos (B,V,3) + tet tile in smem
    const float* pos; int V;
    __device__ __forceinline__ void load(int b, long long, const int4& id, Tet12& t) const {
        const float* p = pos + (size_t)b * V * 3;
        const float* pa = p + (size_t)id.x * 3; const float* pb = p + (size_t)id.y * 3;
        const float* pc = p + (size_t)id.z * 3; const float* pd = p + (size_t)id.w * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) { t.a[k] = __ldg(pa + k); t.b[k] = __ldg(pb + k); t.c[k] = __ldg(pc + k); t.d[k] = __ldg(pd + k); }
    }
};
struct SoupSrc {        // tet_bxfx4x3 (B,T,4,3)
    const float* soup; int T;
    __device__ __forceinline__ void load(int b, long long tet, const int4&, Tet12& t) const {
        const float4* q = reinterpret_cast<const float4*>(soup + ((size_t)b * T + tet) * 12);
        float4 x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);
        t.a[0] = x.x; t.a[1] = x.y; t.a[2] = x.z; t.b[0] = x.w; t.b[1] = y.x; t.b[2] = y.y;
        t.c[0] = y.z; t.c[1] = y.w; t.c[2] = z.x; t.d[0] = z.y; t.d[1] = z.z; t.d[2] = z.w;
    }
};

// stage a tile of the topology with TMA (full tiles) or plain loads (ragged tail / unaligned)
__device__ __forceinline__ void stage_tile(const int32_t* tet, const float* inv_v, int T, int tile0, int n,
                                           int4* s_idx, float* s_inv, uint64_t* bar, bool want_idx, bool want_inv) {
    bool full = (n == E_TILE);
    if (full) {
        if (threadIdx.x == 0) {
            unsigned bytes = (want_idx ? E_TILE * 16u : 0u) + (want_inv ? E_TILE * 36u : 0u);
            mbar_expect_tx(bar, bytes);
            if (want_idx) tma_load_1d(s_idx, tet + (size_t)tile0 * 4, E_TILE * 16u, bar);
            if (want_inv) tma_load_1d(s_inv, inv_v + (size_t)tile0 * 9, E_TILE * 36u, bar);
        }
        mbar_wait(bar, 0);
    } else {
        if (want_idx)
            for (int i = threadIdx.x; i < n; i += blockDim.x) s_idx[i] = reinterpret_cast<const int4*>(tet)[tile0 + i];
        if (want_inv)
            for (int i = threadIdx.x; i < n * 9; i += blockDim.x) s_inv[i] = inv_v[(size_t)tile0 * 9 + i];
        __syncthreads();
    }
}

// ---- forward pass 1: energies + volumes ------------------------------------------------------------
// acc layout (double, per sample, 8 slots): 0 amips_sum 1 edge_sum 2 vol_sum 3 m4 4 m3 5 mu
template <typename Src>
__global__ void __launch_bounds__(E_TILE) energies_fwd_kernel(Src src, const int32_t* __restrict__ tet,
                                                              const float* __restrict__ inv_v, int B, int T, int flags,
                                                              float* __restrict__ vol_out, double* __restrict__ acc) {
    __shared__ __align__(128) int4 s_idx[E_TILE];
    __shared__ __align__(128) float s_inv[E_TILE * 9];
    __shared__ __align__(8) uint64_t bar;
    __shared__ double s_part[E_TILE / 32][E_MAXB][3];
    const int tile0 = blockIdx.x * E_TILE;
    const int n = min(E_TILE, T - tile0);
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    const bool idx_needed = (tet != nullptr);
    const bool inv_needed = (flags & DTB_ENERGY_AMIPS) != 0;
    stage_tile(tet, inv_v, T, tile0, n, s_idx, s_inv, &bar, idx_needed, inv_needed);

    const bool active = tid < n;
    int4 id = make_int4(0, 0, 0, 0);
    float M[9];
    if (active) {
        if (idx_needed) id = s_idx[tid];
        if (inv_needed) {
#pragma unroll
            for (int i = 0; i < 9; ++i) M[i] = s_inv[tid * 9 + i];
        }
    }
    for (int b = 0; b < B; ++b) {
        double e_am = 0.0, e_ed = 0.0, e_vo = 0.0;
        if (active) {
            Tet12 t;
            src.load(b, tile0 + tid, id, t);
            if (flags & DTB_ENERGY_AMIPS) {
                float J[9], det, tr, g;
                e_am = (double)amips_energy(t, M, J, det, tr, g);
            }
            if (flags & DTB_ENERGY_EDGE) e_ed = (double)edge_energy(t);
            if (flags & DTB_ENERGY_VOLUME) {
                float a[3], bb[3], c[3];
                float v = tet_volume(t, a, bb, c);
                vol_out[(size_t)b * T + tile0 + tid] = v;
                e_vo = (double)v;
            }
        }
        e_am = warp_sum(e_am); e_ed = warp_sum(e_ed); e_vo = warp_sum(e_vo);
        if ((tid & 31) == 0) {
            s_part[tid >> 5][b][0] = e_am;
            s_part[tid >> 5][b][1] = e_ed;
            s_part[tid >> 5][b][2] = e_vo;
        }
    }
    __syncthreads();
    for (int i = tid; i < B * 3; i += blockDim.x) {
        int b = i / 3, q = i % 3;
        if (flags & (1 << q)) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < E_TILE / 32; ++w) s += s_part[w][b][q];
            atomicAdd(&acc[(size_t)b * 8 + q], s);
        }
    }
}

// ---- forward pass 2: centred moments of the volume -------------------------------------------------
__global__ void __launch_bounds__(256) volume_moments_kernel(const float* __restrict__ vol, int B, int T, double* __restrict__ acc) {
    int b = blockIdx.y;
    float mu = (float)(acc[(size_t)b * 8 + 2] / (double)T);     // torch.mean in fp32 (deftet.py:258)
    double m4 = 0.0, m3 = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
        float d = vol[(size_t)b * T + i] - mu;
        float d2 = d * d;
        m4 += (double)(d2 * d2);
        m3 += (double)(d2 * d);
    }
    m4 = warp_sum(m4); m3 = warp_sum(m3);
    __shared__ double s4[8], s3[8];
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s4[w] = m4; s3[w] = m3; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, c = 0;
        for (int i = 0; i < 8; ++i) { a += s4[i]; c += s3[i]; }
        atomicAdd(&acc[(size_t)b * 8 + 3], a);
        atomicAdd(&acc[(size_t)b * 8 + 4], c);
    }
}

__global__ void energies_finalize_kernel(double* __restrict__ acc, int B, int T, int flags, float* __restrict__ amips,
                                         float* __restrict__ edge, float* __restrict__ volvar) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double* a = acc + (size_t)b * 8;
    if ((flags & DTB_ENERGY_AMIPS) && amips) amips[b] = (float)(a[0] / (double)T);
    if ((flags & DTB_ENERGY_EDGE) && edge) edge[b] = (float)(a[1] / (6.0 * (double)T));
    if ((flags & DTB_ENERGY_VOLUME) && volvar) {
        volvar[b] = (float)a[3];
        a[5] = (double)(float)(a[2] / (double)T);
    }
}

// ---- backward -------------------------------------------------------------------------------------
struct ScatterDst {     // atomics into grad_pos (B,V,3)
    float* grad; int V;
    __device__ __forceinline__ void store(int b, long long, const int4& id, const float* ga, const float* gb,
                                          const float* gc, const float* gd) const {
        float* g = grad + (size_t)b * V * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicAdd(g + (size_t)id.x * 3 + k, ga[k]);
            atomicAdd(g + (size_t)id.y * 3 + k, gb[k]);
            atomicAdd(g + (size_t)id.z * 3 + k, gc[k]);
            atomicAdd(g + (size_t)id.w * 3 + k, gd[k]);
        }
    }
};
struct ScatterDst4 {    // same, into a padded (B,V,4) buffer: ONE 16-byte vector reduction per vertex (red.global.add.v4.f32, sm_90+)
    float4* grad; int V;
    __device__ __forceinline__ void store(int b, long long, const int4& id, const float* ga, const float* gb,
                                          const float* gc, const float* gd) const {
        float4* g = grad + (size_t)b * V;
        atomicAdd(g + id.x, make_float4(ga[0], ga[1], ga[2], 0.f));
        atomicAdd(g + id.y, make_float4(gb[0], gb[1], gb[2], 0.f));
        atomicAdd(g + id.z, make_float4(gc[0], gc[1], gc[2], 0.f));
        atomicAdd(g + id.w, make_float4(gd[0], gd[1], gd[2], 0.f));
    }
};
struct SoupDst {        // dense gradient (B,T,4,3), overwritten
    float* grad; int T;
    __device__ __forceinline__ void store(int b, long long tet, const int4&, const float* ga, const float* gb,
                                          const float* gc, const float* gd) const {
        float4* q = reinterpret_cast<float4*>(grad + ((size_t)b * T + tet) * 12);
        q[0] = make_float4(ga[0], ga[1], ga[2], gb[0]);
        q[1] = make_float4(gb[1], gb[2], gc[0], gc[1]);
        q[2] = make_float4(gc[2], gd[0], gd[1], gd[2]);
    }
};

template <typename Src, typename Dst>
__global__ void __launch_bounds__(E_TILE) energies_bwd_kernel(Src src, Dst dst, const int32_t* __restrict__ tet,
                                                              const float* __restrict__ inv_v, int B, int T, int flags,
                                                              const double* __restrict__ acc, const float* __restrict__ g_amips,
                                                              const float* __restrict__ g_edge, const float* __restrict__ g_vol) {
    __shared__ __align__(128) int4 s_idx[E_TILE];
    __shared__ __align__(128) float s_inv[E_TILE * 9];
    __shared__ __align__(8) uint64_t bar;
    const int tile0 = blockIdx.x * E_TILE;
    const int n = min(E_TILE, T - tile0);
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    const bool idx_needed = (tet != nullptr);
    const bool inv_needed = (flags & DTB_ENERGY_AMIPS) != 0;
    stage_tile(tet, inv_v, T, tile0, n, s_idx, s_inv, &bar, idx_needed, inv_needed);
    if (tid >= n) return;
    int4 id = make_int4(0, 0, 0, 0);
    if (idx_needed) id = s_idx[tid];
    float M[9];
    if (inv_needed) {
#pragma unroll
        for (int i = 0; i < 9; ++i) M[i] = s_inv[tid * 9 + i];
    }
    const float invT = 1.0f / (float)T;
    for (int b = 0; b < B; ++b) {
        Tet12 t;
        src.load(b, tile0 + tid, id, t);
        float ga[3] = {0, 0, 0}, gb[3] = {0, 0, 0}, gc[3] = {0, 0, 0}, gd[3] = {0, 0, 0};
        if ((flags & DTB_ENERGY_AMIPS) && g_amips) {
            float J[9], det, tr, g;
            amips_energy(t, M, J, det, tr, g);
            amips_grad(M, J, det, tr, g, g_amips[b] * invT, ga, gb, gc, gd);
        }
        if ((flags & DTB_ENERGY_EDGE) && g_edge) edge_grad(t, g_edge[b] * invT * (1.0f / 6.0f), ga, gb, gc, gd);
        if ((flags & DTB_ENERGY_VOLUME) && g_vol) {
            float a[3], bb[3], c[3];
            float v = tet_volume(t, a, bb, c);
            float mu = (float)acc[(size_t)b * 8 + 5];
            float s3 = (float)acc[(size_t)b * 8 + 4];
            float d = v - mu;
            float dLdV = 4.f * d * d * d - 4.f * invT * s3;     // mean term is not detached (deftet.py:258-262)
            volume_grad(a, bb, c, g_vol[b] * dLdV, ga, gb, gc, gd);
        }
        dst.store(b, tile0 + tid, id, ga, gb, gc, gd);
    }
}

// ---- rest-pose inverse (tet_inverse_v + my_inverse) -------------------------------------------------
__global__ void inverse_v_kernel(const float* __restrict__ pos0, const int32_t* __restrict__ tet, int T, float* __restrict__ inv_v) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int4 id = reinterpret_cast<const int4*>(tet)[t];
    float O[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float a = pos0[(size_t)id.x * 3 + k] * AMIPS_SCALE;
        O[0 + k] = pos0[(size_t)id.y * 3 + k] * AMIPS_SCALE - a;
        O[3 + k] = pos0[(size_t)id.z * 3 + k] * AMIPS_SCALE - a;
        O[6 + k] = pos0[(size_t)id.w * 3 + k] * AMIPS_SCALE - a;
    }
    float c0[3], c1[3], c2[3];
    cross3(O + 3, O + 6, c0); cross3(O + 6, O + 0, c1); cross3(O + 0, O + 3, c2);
    float det = dot3(O, c0);
    float R[9];
    if (fabsf(det) < 1e-10f) {       // singular rest tet -> identity (deftet.py:213-218)
        R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    } else {
        float r = 1.0f / det;        // inverse = adj(O)/det, adj = cof^T
#pragma unroll
        for (int k = 0; k < 3; ++k) { R[k * 3 + 0] = c0[k] * r; R[k * 3 + 1] = c1[k] * r; R[k * 3 + 2] = c2[k] * r; }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) inv_v[(size_t)t * 9 + i] = R[i];
}

}  // namespace dtb

using namespace dtb;

extern "C" size_t dtb_tet_energies_workspace(int B, int V, int T) {
    (void)V;
    return align_up((size_t)B * T * sizeof(float), 256);
}

static int energies_forward_impl(const float* pos, const float* soup, const int32_t* tet, const float* inv_v, int B, int V, int T,
                                 int flags, float* amips, float* edge, float* volvar, double* stats, void* workspace,
                                 size_t workspace_bytes, cudaStream_t st) {
    DTB_REQUIRE(B > 0 && T > 0, "tet_energies: empty batch or grid (B=%d T=%d)", B, T);
    DTB_REQUIRE(stats != nullptr, "tet_energies: stats buffer (B*8 doubles) is required");
    DTB_REQUIRE(!(flags & DTB_ENERGY_AMIPS) || inv_v, "tet_energies: AMIPS requested without inverse_v");
    DTB_REQUIRE_ALIGNED16(tet, "tet_energies: tet");
    DTB_REQUIRE_ALIGNED16(inv_v, "tet_energies: inverse_v");
    DTB_REQUIRE_ALIGNED16(soup, "tet_energies: tet_bxfx4x3");
    float* vol = nullptr;
    if (flags & DTB_ENERGY_VOLUME) {
        if (workspace_bytes < dtb_tet_energies_workspace(B, V, T) || !workspace) {
            set_error("tet_energies: workspace too small (%zu < %zu)", workspace_bytes, dtb_tet_energies_workspace(B, V, T));
            return DTB_EWORKSPACE;
        }
        vol = (float*)workspace;
    }
    DTB_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * 8 * sizeof(double), st));
    int tiles = cdiv(T, E_TILE);
    for (int b0 = 0; b0 < B; b0 += E_MAXB) {
        int nb = min(E_MAXB, B - b0);
        if (soup) {
            SoupSrc s{soup + (size_t)b0 * T * 12, T};
            energies_fwd_kernel<SoupSrc><<<tiles, E_TILE, 0, st>>>(s, nullptr, inv_v, nb, T, flags, vol ? vol + (size_t)b0 * T : nullptr,
                                                                    stats + (size_t)b0 * 8);
        } else {
            IndexedSrc s{pos + (size_t)b0 * V * 3, V};
            prof_begin(PROF_ENERGIES_FWD, st);
            energies_fwd_kernel<IndexedSrc><<<tiles, E_TILE, 0, st>>>(s, tet, inv_v, nb, T, flags, vol ? vol + (size_t)b0 * T : nullptr,
                                                                       stats + (size_t)b0 * 8);
            prof_end(PROF_ENERGIES_FWD, st);
        }
        DTB_LAUNCH_CHECK("energies_fwd");
    }
    if (flags & DTB_ENERGY_VOLUME) {
        dim3 g(min(cdiv(T, 256 * 4), 64), B);
        volume_moments_kernel<<<g, 256, 0, st>>>(vol, B, T, stats);
        DTB_LAUNCH_CHECK("volume_moments");
    }
    energies_finalize_kernel<<<cdiv(B, 128), 128, 0, st>>>(stats, B, T, flags, amips, edge, volvar);
    DTB_LAUNCH_CHECK("energies_finalize");
    return DTB_OK;
}

extern "C" int dtb_tet_energies_forward(const float* pos, const int32_t* tet, const float* inv_v, int B, int V, int T, int flags,
                                        float* amips, float* edge, float* volvar, double* stats, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(pos && tet, "tet_energies_forward: null pos/tet");
    return energies_forward_impl(pos, nullptr, tet, inv_v, B, V, T, flags, amips, edge, volvar, stats, workspace, workspace_bytes,
                                 (cudaStream_t)stream);
}

extern "C" int dtb_tet_energies_forward_soup(const float* tet_bxfx4x3, const float* inv_v, int B, int T, int flags, float* amips,
                                             float* edge, float* volvar, double* stats, void* workspace, size_t workspace_bytes,
                                             void* stream) {
    DTB_REQUIRE(tet_bxfx4x3, "tet_energies_forward_soup: null input");
    return energies_forward_impl(nullptr, tet_bxfx4x3, nullptr, inv_v, B, 0, T, flags, amips, edge, volvar, stats, workspace,
                                 workspace_bytes, (cudaStream_t)stream);
}

extern "C" int dtb_tet_energies_backward(const float* pos, const int32_t* tet, const float* inv_v, int B, int V, int T, int flags,
                                         const double* stats, const float* g_amips, const float* g_edge, const float* g_volvar,
                                         float* grad_pos, void* stream) {
    DTB_REQUIRE(pos && tet && grad_pos && stats, "tet_energies_backward: null argument");
    DTB_REQUIRE(B > 0 && T > 0, "tet_energies_backward: empty batch or grid");
    DTB_REQUIRE_ALIGNED16(tet, "tet_energies_backward: tet");
    DTB_REQUIRE_ALIGNED16(inv_v, "tet_energies_backward: inverse_v");
    cudaStream_t st = (cudaStream_t)stream;
    int tiles = cdiv(T, E_TILE);
    IndexedSrc s{pos, V};
    ScatterDst d{grad_pos, V};
    prof_begin(PROF_ENERGIES_BWD, st);
    energies_bwd_kernel<IndexedSrc, ScatterDst><<<tiles, E_TILE, 0, st>>>(s, d, tet, inv_v, B, T, flags, stats, g_amips, g_edge, g_volvar);
    DTB_LAUNCH_CHECK("energies_bwd");
    prof_end(PROF_ENERGIES_BWD, st);
    return DTB_OK;
}

extern "C" int dtb_tet_energies_backward_soup(const float* tet_bxfx4x3, const float* inv_v, int B, int T, int flags,
                                              const double* stats, const float* g_amips, const float* g_edge, const float* g_volvar,
                                              float* grad_soup, void* stream) {
    DTB_REQUIRE(tet_bxfx4x3 && grad_soup && stats, "tet_energies_backward_soup: null argument");
    DTB_REQUIRE(B > 0 && T > 0, "tet_energies_backward_soup: empty batch or grid");
    DTB_REQUIRE_ALIGNED16(tet_bxfx4x3, "tet_energies_backward_soup: tet_bxfx4x3");
    DTB_REQUIRE_ALIGNED16(grad_soup, "tet_energies_backward_soup: grad_soup");
    DTB_REQUIRE_ALIGNED16(inv_v, "tet_energies_backward_soup: inverse_v");
    cudaStream_t st = (cudaStream_t)stream;
    int tiles = cdiv(T, E_TILE);
    SoupSrc s{tet_bxfx4x3, T};
    SoupDst d{grad_soup, T};
    energies_bwd_kernel<SoupSrc, SoupDst><<<tiles, E_TILE, 0, st>>>(s, d, nullptr, inv_v, B, T, flags, stats, g_amips, g_edge, g_volvar);
    DTB_LAUNCH_CHECK("energies_bwd_soup");
    return DTB_OK;
}

extern "C" int dtb_tet_inverse_v(const float* pos0, const int32_t* tet, int V, int T, float* inv_v, void* stream) {
    (void)V;
    DTB_REQUIRE(pos0 && tet && inv_v, "tet_inverse_v: null argument");
    DTB_REQUIRE_ALIGNED16(tet, "tet_inverse_v: tet");
    if (T == 0) return DTB_OK;
    inverse_v_kernel<<<cdiv(T, 256), 256, 0, (cudaStream_t)stream>>>(pos0, tet, T, inv_v);
    DTB_LAUNCH_CHECK("inverse_v");
    return DTB_OK;
}

// Same as dtb_tet_energies_backward, but grad_pos4 is a zero-filled PADDED (B,V,4) f32 buffer (16-byte aligned): every vertex
// update is one 16-byte vector reduction instead of three scalar ones (the kernel is bound by the RED issue rate).
extern "C" int dtb_tet_energies_backward_v4(const float* pos, const int32_t* tet, const float* inv_v, int B, int V, int T, int flags,
                                            const double* stats, const float* g_amips, const float* g_edge, const float* g_volvar,
                                            float* grad_pos4, void* stream) {
    DTB_REQUIRE(pos && tet && grad_pos4 && stats, "tet_energies_backward_v4: null argument");
    DTB_REQUIRE(B > 0 && T > 0, "tet_energies_backward_v4: empty batch or grid");
    DTB_REQUIRE_ALIGNED16(tet, "tet_energies_backward_v4: tet");
    DTB_REQUIRE_ALIGNED16(inv_v, "tet_energies_backward_v4: inverse_v");
    DTB_REQUIRE((((size_t)grad_pos4) & 15) == 0, "tet_energies_backward_v4: grad_pos4 must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int tiles = cdiv(T, E_TILE);
    IndexedSrc s{pos, V};
    ScatterDst4 d{reinterpret_cast<float4*>(grad_pos4), V};
    prof_begin(PROF_ENERGIES_BWD, st);
    energies_bwd_kernel<IndexedSrc, ScatterDst4><<<tiles, E_TILE, 0, st>>>(s, d, tet, inv_v, B, T, flags, stats, g_amips, g_edge, g_volvar);
    DTB_LAUNCH_CHECK("energies_bwd_v4");
    prof_end(PROF_ENERGIES_BWD, st);
    return DTB_OK;
}
