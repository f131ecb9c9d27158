// N4, second half: sampling the encoder's voxel features at grid vertices / tet centroids.
// Replaces layers/pv_module/functional/devoxelization.py:47-53 `trilinear_devoxelize` (the definition that is live in the reference:
// coords -> (coords*2+1)/r-1, flip, torch F.grid_sample(mode='bilinear', padding_mode='border', align_corners=False)) as
// layers/pc_model.py:182-194 `sample_f` calls it once per encoder level ((64 ch, R=32), (128, 16), (512, 8), pc_model.py:50), and its
// autograd (grid_sample backward w.r.t. the volume and the sampling grid).
//
// Data layout in HBM: feat (B,C,R,R,R) f32 contiguous; coords (B,3,N) with arbitrary element strides (the reference hands over a
// permuted view) or, fused `sample_f` mode, positions (B,N,3); out (B,C,N) f32, optionally a channel slice of a wider
// (B,sumC,N) tensor (batch stride passed in) so that the torch.cat over levels (pc_model.py:194) is never materialised.
//
// Kernels.  A CTA owns (sample b, a chunk of channels, a range of points): the chunk's R^3 planes are staged in shared memory
// ONCE (1-D TMA bulk copies when a plane cannot be interleaved, else a 4-channel interleaved layout so that one LDS.128 fetches a
// corner for 4 channels) and every point of the range gathers its 8 corners from there; outputs leave as coalesced 4-byte stores,
// n fastest.  The volume therefore crosses HBM/L2 once per CTA instead of 8 scattered sectors per (point, channel).
//   forward          : devox_gather_kernel<VEC,0>   HBM bound on the output write, 4*B*C*N bytes (+ 4*B*C*R^3 + 12*B*N in)
//   d/d coords       : devox_gather_kernel<VEC,1>   same staging; reads grad_out instead of writing out; three reductions per
//                      (point, chunk) leave as warp-coalesced REDs
//   d/d feat         : devox_scatter_kernel         accumulation planes in shared memory; consecutive points that fall in the same
//                      voxel (the grid's vertex order is spatially coherent) are merged first with a segmented warp scan, because a
//                      float atomicAdd on shared memory is a CAS loop (ATOMS.CAST.SPIN) and would serialise on such runs
// R^3 planes that do not fit in shared memory (R > 36) fall back to *_global kernels that gather through L2.
#include "common.cuh"
#include "prims.cuh"
#include "deftet_b200.h"

namespace dtb {

struct DevoxArgs {
    const float* feat;        // (B,C,R^3)
    const float* coords;      // element strides cs_b, cs_k, cs_n
    long long cs_b, cs_k, cs_n;
    int C, N, R, R3;
    int cc;                   // channels per CTA
    int pts_per_cta;
    int from_pos;             // coords are positions: c = clamp((p + 0.5) * R, 0, R - 1)   (pc_model.py:186-191)
    float inv_r;              // 1/R, used instead of the division when R is a power of two (same rounding, a tenth of the instructions)
    int pow2;
};

// one axis of torch's grid_sampler_compute_source_index_set_grad after the reference's normalisation, same rounding sequence:
//   g = (c*2+1)/r - 1   (devoxelization.py:48);  u = ((g+1)*size - 1)/2  (unnormalize, align_corners=False);  clip to [0, size-1]
// mult = d u_clipped / d (input coordinate): 0 where the border clip (or the caller's clamp in from_pos mode) is active.
__device__ __forceinline__ void dv_axis(float c, float rf, float inv_r, bool pow2, int R, bool from_pos, int& l, int& hi, float& f, float& g,
                                        float& mult) {
    float m = 1.f;
    if (from_pos) {
        float t = xmul(xadd(c, 0.5f), rf);
        m = (t >= 0.f && t <= rf - 1.f) ? rf : 0.f;         // torch.clamp backward passes on the closed interval
        c = fminf(fmaxf(t, 0.f), rf - 1.f);
    }
    float t2 = xadd(xmul(c, 2.f), 1.f);
    float gn = xsub(pow2 ? xmul(t2, inv_r) : xdiv(t2, rf), 1.f);
    float u = xmul(xsub(xmul(xadd(gn, 1.f), rf), 1.f), 0.5f);          // "/ 2" is exact as a product
    if (u <= 0.f) { u = 0.f; m = 0.f; }
    else if (u >= rf - 1.f) { u = rf - 1.f; m = 0.f; }
    float lf = floorf(u);
    l = (int)lf;
    f = u - lf;                // weight of the upper neighbour
    g = (lf + 1.f) - u;        // weight of the lower neighbour, as torch forms it (ix_bse - ix)
    hi = (l + 1 < R) ? 1 : 0;  // the upper neighbour is out of bounds only when u == R-1, where f == 0
    mult = m;
}

struct DvPoint {
    int base, o0, o1, o2;      // linear voxel index of the low corner; offsets to the upper neighbour along coords[0], [1], [2]
    float f0, g0, f1, g1, f2, g2;
    float m0, m1, m2;
};

__device__ __forceinline__ DvPoint dv_point(const DevoxArgs& a, int b, int n) {
    const float* c = a.coords + (size_t)b * a.cs_b + (size_t)n * a.cs_n;
    float c0 = c[0], c1 = c[a.cs_k], c2 = c[2 * a.cs_k];
    const float rf = (float)a.R;
    DvPoint p;
    int l0, l1, l2, h0, h1, h2;
    dv_axis(c0, rf, a.inv_r, a.pow2 != 0, a.R, a.from_pos != 0, l0, h0, p.f0, p.g0, p.m0);
    dv_axis(c1, rf, a.inv_r, a.pow2 != 0, a.R, a.from_pos != 0, l1, h1, p.f1, p.g1, p.m1);
    dv_axis(c2, rf, a.inv_r, a.pow2 != 0, a.R, a.from_pos != 0, l2, h2, p.f2, p.g2, p.m2);
    p.base = (l0 * a.R + l1) * a.R + l2;       // coords[0] indexes the slowest volume axis (the flip in devoxelization.py:50)
    p.o0 = h0 * a.R * a.R; p.o1 = h1 * a.R; p.o2 = h2;
    return p;
}

// the 8 weights in torch's order tnw, tne, tsw, tse, bnw, bne, bsw, bse (x = coords[2] fastest, "t/b" = coords[0])
__device__ __forceinline__ void dv_weights(const DvPoint& p, float w[8]) {
    float gg = p.g2 * p.g1, fg = p.f2 * p.g1, gf = p.g2 * p.f1, ff = p.f2 * p.f1;
    w[0] = gg * p.g0; w[1] = fg * p.g0; w[2] = gf * p.g0; w[3] = ff * p.g0;
    w[4] = gg * p.f0; w[5] = fg * p.f0; w[6] = gf * p.f0; w[7] = ff * p.f0;
}

__device__ __forceinline__ void dv_offsets(const DvPoint& p, int o[8]) {
    o[0] = 0; o[1] = p.o2; o[2] = p.o1; o[3] = p.o1 + p.o2;
    o[4] = p.o0; o[5] = p.o0 + p.o2; o[6] = p.o0 + p.o1; o[7] = p.o0 + p.o1 + p.o2;
}

// ---- staging -------------------------------------------------------------------------------------------------------------------
// VEC == 1: planes [j][v] as in HBM -> one contiguous run of ncc*R3 floats, fetched with 1-D TMA in <= 32 KB pieces when aligned.
// VEC == 4: [j/4][v][4] so that a corner of four channels is one 16-byte shared-memory load.
template <int VEC>
__device__ __forceinline__ void dv_stage(const float* __restrict__ src, int ncc, int R3, float* s, uint64_t* bar) {
    if (VEC == 1) {
        size_t bytes = (size_t)ncc * R3 * 4;
        bool aligned = ((reinterpret_cast<uintptr_t>(src) | bytes) & 15) == 0;
        if (aligned) {
            if (threadIdx.x == 0) {
                mbar_init(bar, 1);
                mbar_fence_init();
                mbar_expect_tx(bar, (unsigned)bytes);
                for (size_t off = 0; off < bytes; off += 32768) {
                    unsigned piece = (unsigned)min((size_t)32768, bytes - off);
                    tma_load_1d(reinterpret_cast<char*>(s) + off, reinterpret_cast<const char*>(src) + off, piece, bar);
                }
            }
            __syncthreads();          // the barrier is initialised before anybody polls it
            mbar_wait(bar, 0);
        } else {
            for (int i = threadIdx.x; i < ncc * R3; i += blockDim.x) s[i] = src[i];
            __syncthreads();
        }
    } else {
        const int groups = ncc / 4;
        for (int i = threadIdx.x; i < groups * R3; i += blockDim.x) {
            int grp = i / R3, v = i - grp * R3;
            const float* q = src + (size_t)grp * 4 * R3 + v;
            reinterpret_cast<float4*>(s)[i] = make_float4(q[0], q[R3], q[2 * (size_t)R3], q[3 * (size_t)R3]);
        }
        __syncthreads();
    }
}

// MODE 0: out[b, c, n] = sum_k w_k feat[b, c, corner_k(n)]
// MODE 1: grad_coords[b, k, n] += mult_k * sum_c gout[b, c, n] * d/du_k (trilinear interpolant)
template <int VEC, int MODE>
__global__ void __launch_bounds__(VEC == 1 ? 1024 : 512) devox_gather_kernel(DevoxArgs a, float* __restrict__ out, const float* __restrict__ gout, long long ob) {
    extern __shared__ __align__(128) float s_feat[];
    __shared__ uint64_t bar;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * a.cc;
    const int ncc = min(a.cc, a.C - c0);
    const int n0 = blockIdx.x * a.pts_per_cta, n1 = min(a.N, n0 + a.pts_per_cta);
    dv_stage<VEC>(a.feat + ((size_t)b * a.C + c0) * a.R3, ncc, a.R3, s_feat, &bar);

    for (int n = n0 + threadIdx.x; n < n1; n += blockDim.x) {
        DvPoint p = dv_point(a, b, n);
        int o[8];
        dv_offsets(p, o);
        if (MODE == 0) {
            float w[8];
            dv_weights(p, w);
            float* dst = out + (size_t)b * ob + (size_t)c0 * a.N + n;
            if (VEC == 1) {
                for (int j = 0; j < ncc; ++j) {
                    const float* q = s_feat + (size_t)j * a.R3 + p.base;
                    float acc = q[0] * w[0];
#pragma unroll
                    for (int k = 1; k < 8; ++k) acc = fmaf(q[o[k]], w[k], acc);
                    dst[(size_t)j * a.N] = acc;
                }
            } else {
                for (int grp = 0; grp < ncc / 4; ++grp) {
                    const float4* q = reinterpret_cast<const float4*>(s_feat) + (size_t)grp * a.R3 + p.base;
                    float4 v = q[0];
                    float a0 = v.x * w[0], a1 = v.y * w[0], a2 = v.z * w[0], a3 = v.w * w[0];
#pragma unroll
                    for (int k = 1; k < 8; ++k) {
                        v = q[o[k]];
                        a0 = fmaf(v.x, w[k], a0); a1 = fmaf(v.y, w[k], a1); a2 = fmaf(v.z, w[k], a2); a3 = fmaf(v.w, w[k], a3);
                    }
                    float* d4 = dst + (size_t)grp * 4 * a.N;
                    d4[0] = a0; d4[a.N] = a1; d4[2 * (size_t)a.N] = a2; d4[3 * (size_t)a.N] = a3;
                }
            }
        } else {
            if (p.m0 == 0.f && p.m1 == 0.f && p.m2 == 0.f) continue;
            // derivative weights: d/du0 pairs (k, k+4); d/du1 pairs (k, k+2); d/du2 pairs (k, k+1)
            const float gg = p.g2 * p.g1, fg = p.f2 * p.g1, gf = p.g2 * p.f1, ff = p.f2 * p.f1;
            const float g2g0 = p.g2 * p.g0, f2g0 = p.f2 * p.g0, g2f0 = p.g2 * p.f0, f2f0 = p.f2 * p.f0;
            const float g1g0 = p.g1 * p.g0, f1g0 = p.f1 * p.g0, g1f0 = p.g1 * p.f0, f1f0 = p.f1 * p.f0;
            float d0 = 0.f, d1 = 0.f, d2 = 0.f;
            const float* gsrc = gout + (size_t)b * ob + (size_t)c0 * a.N + n;
            auto accumulate = [&](const float v[8], float go) {
                float e0 = (v[4] - v[0]) * gg + (v[5] - v[1]) * fg + (v[6] - v[2]) * gf + (v[7] - v[3]) * ff;
                float e1 = (v[2] - v[0]) * g2g0 + (v[3] - v[1]) * f2g0 + (v[6] - v[4]) * g2f0 + (v[7] - v[5]) * f2f0;
                float e2 = (v[1] - v[0]) * g1g0 + (v[3] - v[2]) * f1g0 + (v[5] - v[4]) * g1f0 + (v[7] - v[6]) * f1f0;
                d0 = fmaf(go, e0, d0); d1 = fmaf(go, e1, d1); d2 = fmaf(go, e2, d2);
            };
            if (VEC == 1) {
                for (int j = 0; j < ncc; ++j) {
                    const float* q = s_feat + (size_t)j * a.R3 + p.base;
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = q[o[k]];
                    accumulate(v, gsrc[(size_t)j * a.N]);
                }
            } else {
                for (int grp = 0; grp < ncc / 4; ++grp) {
                    const float4* q = reinterpret_cast<const float4*>(s_feat) + (size_t)grp * a.R3 + p.base;
                    float vx[8], vy[8], vz[8], vw[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { float4 v = q[o[k]]; vx[k] = v.x; vy[k] = v.y; vz[k] = v.z; vw[k] = v.w; }
                    const float* g4 = gsrc + (size_t)grp * 4 * a.N;
                    accumulate(vx, g4[0]); accumulate(vy, g4[a.N]); accumulate(vz, g4[2 * (size_t)a.N]); accumulate(vw, g4[3 * (size_t)a.N]);
                }
            }
            float* gc = out + (size_t)b * a.cs_b + (size_t)n * a.cs_n;
            if (p.m0 != 0.f) atomicAdd(gc, p.m0 * d0);
            if (p.m1 != 0.f) atomicAdd(gc + a.cs_k, p.m1 * d1);
            if (p.m2 != 0.f) atomicAdd(gc + 2 * a.cs_k, p.m2 * d2);
        }
    }
}

// grad_feat[b, c, corner_k(n)] += w_k(n) * gout[b, c, n]; accumulation planes [j][v] in shared memory.
__global__ void __launch_bounds__(1024) devox_scatter_kernel(DevoxArgs a, const float* __restrict__ gout, long long ob, float* __restrict__ grad_feat,
                                     int flush_atomic) {
    extern __shared__ __align__(128) float s_acc[];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * a.cc;
    const int ncc = min(a.cc, a.C - c0);
    const int n0 = blockIdx.x * a.pts_per_cta, n1 = min(a.N, n0 + a.pts_per_cta);
    for (int i = threadIdx.x; i < ncc * a.R3; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int span = n1 - n0;
    for (int i0 = 0; i0 < span; i0 += blockDim.x) {      // warp-uniform trip count: the scan below needs every lane
        const int n = n0 + i0 + threadIdx.x;
        const bool active = n < n1;
        DvPoint p;
        float w[8];
        int o[8];
        if (active) { p = dv_point(a, b, n); dv_weights(p, w); dv_offsets(p, o); }
        else {
            p.base = -2 - lane;
#pragma unroll
            for (int k = 0; k < 8; ++k) { w[k] = 0.f; o[k] = 0; }
        }
        // runs of consecutive lanes in the same voxel: head flags, first lane of my run, am I its last lane
        int prev = __shfl_up_sync(0xffffffffu, p.base, 1);
        bool head = (lane == 0) || (prev != p.base);
        unsigned heads = __ballot_sync(0xffffffffu, head);
        const bool merge = heads != 0xffffffffu;
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
        const float* gsrc = gout + (size_t)b * ob + (size_t)c0 * a.N + n;
        for (int j = 0; j < ncc; ++j) {
            float go = active ? gsrc[(size_t)j * a.N] : 0.f;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = w[k] * go;
            if (merge) {
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float t = __shfl_up_sync(0xffffffffu, v[k], d);
                        if (lane - d >= start) v[k] += t;
                    }
                }
            }
            if (active && tail) {
                float* q = s_acc + (size_t)j * a.R3 + p.base;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (v[k] != 0.f) atomicAdd(q + o[k], v[k]);
            }
        }
    }
    __syncthreads();
    float* dst = grad_feat + ((size_t)b * a.C + c0) * a.R3;
    if (flush_atomic) {
        for (int i = threadIdx.x; i < ncc * a.R3; i += blockDim.x) { float v = s_acc[i]; if (v != 0.f) atomicAdd(dst + i, v); }
    } else {
        for (int i = threadIdx.x; i < ncc * a.R3; i += blockDim.x) dst[i] = s_acc[i];
    }
}

// ---- four consecutive points per thread ---------------------------------------------------------------------------------------------
// The gather above costs 8 shared-memory operands per output, and the shared-memory data pipe is what it saturates (ncu: 81 % of
// the l1tex wavefront peak at R = 8).  Grid vertices and tet centroids arrive in lattice order, so consecutive points mostly share
// their voxel: a thread that owns points n..n+3 keeps the 8 corner values in registers and refills them only when the voxel changes,
// leaves with one 16-byte store per channel row, and amortises the per-point setup over every channel of the chunk.
struct DvQuad {
    int base[4];
    int hib[4];             // bit 0 / 1 / 2: the upper neighbour along coords[2] / [1] / [0] exists
};

__device__ __forceinline__ int dv_hib(const DvPoint& p) { return (p.o2 ? 1 : 0) | (p.o1 ? 2 : 0) | (p.o0 ? 4 : 0); }

template <typename T>
__device__ __forceinline__ void dv_refill(const T* __restrict__ plane, int base, int hib, int R, T cv[8]) {
    const int o2 = hib & 1, o1 = (hib & 2) ? R : 0, o0 = (hib & 4) ? R * R : 0;
    const T* q = plane + base;
    cv[0] = q[0]; cv[1] = q[o2]; cv[2] = q[o1]; cv[3] = q[o1 + o2];
    cv[4] = q[o0]; cv[5] = q[o0 + o2]; cv[6] = q[o0 + o1]; cv[7] = q[o0 + o1 + o2];
}

__device__ __forceinline__ void dv_store_row(float* dst, const float r[4], int nvalid) {
    if (nvalid == 4 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
    else
        for (int i = 0; i < nvalid; ++i) dst[i] = r[i];
}

__device__ __forceinline__ void dv_load_row(const float* src, float r[4], int nvalid) {
    if (nvalid == 4 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        float4 v = *reinterpret_cast<const float4*>(src);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = i < nvalid ? src[i] : 0.f;
    }
}

template <int VEC>
__global__ void __launch_bounds__(512, 1) devox_gather4_kernel(DevoxArgs a, float* __restrict__ out, long long ob) {
    extern __shared__ __align__(128) float s_feat[];
    __shared__ uint64_t bar;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * a.cc;
    const int ncc = min(a.cc, a.C - c0);
    const int n0 = blockIdx.x * a.pts_per_cta, n1 = min(a.N, n0 + a.pts_per_cta);
    dv_stage<VEC>(a.feat + ((size_t)b * a.C + c0) * a.R3, ncc, a.R3, s_feat, &bar);
    const int span = n1 - n0;
    for (int q = threadIdx.x; 4 * q < span; q += blockDim.x) {
        const int n = n0 + 4 * q;
        const int nvalid = min(4, n1 - n);
        DvQuad Q;
        float w[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            DvPoint p = dv_point(a, b, min(n + i, n1 - 1));      // a short last quad repeats its last point; the stores skip it
            Q.base[i] = p.base; Q.hib[i] = dv_hib(p);
            dv_weights(p, w[i]);
        }
        float* dst = out + (size_t)b * ob + (size_t)c0 * a.N + n;
        if (VEC == 1) {
            for (int j = 0; j < ncc; ++j) {
                const float* plane = s_feat + (size_t)j * a.R3;
                float cv[8], r[4];
                int cur = -1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (Q.base[i] != cur) { cur = Q.base[i]; dv_refill(plane, cur, Q.hib[i], a.R, cv); }
                    float acc = cv[0] * w[i][0];
#pragma unroll
                    for (int k = 1; k < 8; ++k) acc = fmaf(cv[k], w[i][k], acc);
                    r[i] = acc;
                }
                dv_store_row(dst + (size_t)j * a.N, r, nvalid);
            }
        } else {
            for (int grp = 0; grp < ncc / 4; ++grp) {
                const float4* plane = reinterpret_cast<const float4*>(s_feat) + (size_t)grp * a.R3;
                float4 cv[8];
                float r0[4], r1[4], r2[4], r3[4];
                int cur = -1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (Q.base[i] != cur) { cur = Q.base[i]; dv_refill(plane, cur, Q.hib[i], a.R, cv); }
                    float a0 = cv[0].x * w[i][0], a1 = cv[0].y * w[i][0], a2 = cv[0].z * w[i][0], a3 = cv[0].w * w[i][0];
#pragma unroll
                    for (int k = 1; k < 8; ++k) {
                        a0 = fmaf(cv[k].x, w[i][k], a0); a1 = fmaf(cv[k].y, w[i][k], a1);
                        a2 = fmaf(cv[k].z, w[i][k], a2); a3 = fmaf(cv[k].w, w[i][k], a3);
                    }
                    r0[i] = a0; r1[i] = a1; r2[i] = a2; r3[i] = a3;
                }
                float* d4 = dst + (size_t)grp * 4 * a.N;
                dv_store_row(d4, r0, nvalid); dv_store_row(d4 + a.N, r1, nvalid);
                dv_store_row(d4 + 2 * (size_t)a.N, r2, nvalid); dv_store_row(d4 + 3 * (size_t)a.N, r3, nvalid);
            }
        }
    }
}

// d/du of the trilinear interpolant at one point for one channel, times the upstream gradient
__device__ __forceinline__ void dv_dcoord(const float v[8], const float fg[6], float go, float d[3]) {
    const float f0 = fg[0], g0 = fg[1], f1 = fg[2], g1 = fg[3], f2 = fg[4], g2 = fg[5];
    float e0 = (v[4] - v[0]) * (g2 * g1) + (v[5] - v[1]) * (f2 * g1) + (v[6] - v[2]) * (g2 * f1) + (v[7] - v[3]) * (f2 * f1);
    float e1 = (v[2] - v[0]) * (g2 * g0) + (v[3] - v[1]) * (f2 * g0) + (v[6] - v[4]) * (g2 * f0) + (v[7] - v[5]) * (f2 * f0);
    float e2 = (v[1] - v[0]) * (g1 * g0) + (v[3] - v[2]) * (f1 * g0) + (v[5] - v[4]) * (g1 * f0) + (v[7] - v[6]) * (f1 * f0);
    d[0] = fmaf(go, e0, d[0]); d[1] = fmaf(go, e1, d[1]); d[2] = fmaf(go, e2, d[2]);
}

template <int VEC>
__global__ void __launch_bounds__(512, 1) devox_gradcoords4_kernel(DevoxArgs a, float* __restrict__ gcoords, const float* __restrict__ gout,
                                                                  long long ob) {
    extern __shared__ __align__(128) float s_feat[];
    __shared__ uint64_t bar;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * a.cc;
    const int ncc = min(a.cc, a.C - c0);
    const int n0 = blockIdx.x * a.pts_per_cta, n1 = min(a.N, n0 + a.pts_per_cta);
    dv_stage<VEC>(a.feat + ((size_t)b * a.C + c0) * a.R3, ncc, a.R3, s_feat, &bar);
    const int span = n1 - n0;
    for (int q = threadIdx.x; 4 * q < span; q += blockDim.x) {
        const int n = n0 + 4 * q;
        const int nvalid = min(4, n1 - n);
        DvQuad Q;
        float fg[4][6], m[4][3], d[4][3];
        bool any = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            DvPoint p = dv_point(a, b, min(n + i, n1 - 1));
            Q.base[i] = p.base; Q.hib[i] = dv_hib(p);
            fg[i][0] = p.f0; fg[i][1] = p.g0; fg[i][2] = p.f1; fg[i][3] = p.g1; fg[i][4] = p.f2; fg[i][5] = p.g2;
            const bool live = i < nvalid;
            m[i][0] = live ? p.m0 : 0.f; m[i][1] = live ? p.m1 : 0.f; m[i][2] = live ? p.m2 : 0.f;
            d[i][0] = d[i][1] = d[i][2] = 0.f;
            any = any || m[i][0] != 0.f || m[i][1] != 0.f || m[i][2] != 0.f;
        }
        if (!any) continue;                        // every coordinate of the quad is clipped: no gradient
        const float* gsrc = gout + (size_t)b * ob + (size_t)c0 * a.N + n;
        if (VEC == 1) {
            for (int j = 0; j < ncc; ++j) {
                const float* plane = s_feat + (size_t)j * a.R3;
                float cv[8], go[4];
                dv_load_row(gsrc + (size_t)j * a.N, go, nvalid);
                int cur = -1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (Q.base[i] != cur) { cur = Q.base[i]; dv_refill(plane, cur, Q.hib[i], a.R, cv); }
                    dv_dcoord(cv, fg[i], go[i], d[i]);
                }
            }
        } else {
            for (int grp = 0; grp < ncc / 4; ++grp) {
                const float4* plane = reinterpret_cast<const float4*>(s_feat) + (size_t)grp * a.R3;
                float4 cv[8];
                float g0[4], g1[4], g2[4], g3[4];
                const float* g4 = gsrc + (size_t)grp * 4 * a.N;
                dv_load_row(g4, g0, nvalid); dv_load_row(g4 + a.N, g1, nvalid);
                dv_load_row(g4 + 2 * (size_t)a.N, g2, nvalid); dv_load_row(g4 + 3 * (size_t)a.N, g3, nvalid);
                int cur = -1;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (Q.base[i] != cur) { cur = Q.base[i]; dv_refill(plane, cur, Q.hib[i], a.R, cv); }
                    float vx[8], vy[8], vz[8], vw[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) { vx[k] = cv[k].x; vy[k] = cv[k].y; vz[k] = cv[k].z; vw[k] = cv[k].w; }
                    dv_dcoord(vx, fg[i], g0[i], d[i]); dv_dcoord(vy, fg[i], g1[i], d[i]);
                    dv_dcoord(vz, fg[i], g2[i], d[i]); dv_dcoord(vw, fg[i], g3[i], d[i]);
                }
            }
        }
        // reductions into grad_coords: 16-byte vector reductions when the four points are contiguous there ((B,3,N) rows), else scalars
        float* gc = gcoords + (size_t)b * a.cs_b + (size_t)n * a.cs_n;
        if (a.cs_k == 1 && a.cs_n == 3 && nvalid == 4 && (reinterpret_cast<uintptr_t>(gc) & 15) == 0) {
            // positions (B,N,3): the quad's twelve gradients are contiguous
            float4* g4 = reinterpret_cast<float4*>(gc);
            atomicAdd(g4, make_float4(m[0][0] * d[0][0], m[0][1] * d[0][1], m[0][2] * d[0][2], m[1][0] * d[1][0]));
            atomicAdd(g4 + 1, make_float4(m[1][1] * d[1][1], m[1][2] * d[1][2], m[2][0] * d[2][0], m[2][1] * d[2][1]));
            atomicAdd(g4 + 2, make_float4(m[2][2] * d[2][2], m[3][0] * d[3][0], m[3][1] * d[3][1], m[3][2] * d[3][2]));
            continue;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float* row = gc + (size_t)k * a.cs_k;
            if (a.cs_n == 1 && nvalid == 4 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
                atomicAdd(reinterpret_cast<float4*>(row), make_float4(m[0][k] * d[0][k], m[1][k] * d[1][k], m[2][k] * d[2][k], m[3][k] * d[3][k]));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (m[i][k] != 0.f) atomicAdd(row + (size_t)i * a.cs_n, m[i][k] * d[i][k]);
            }
        }
    }
}

__device__ __forceinline__ void dv_flush_acc(float* __restrict__ plane, int base, int hib, int R, const float acc[8]) {
    const int o2 = hib & 1, o1 = (hib & 2) ? R : 0, o0 = (hib & 4) ? R * R : 0;
    float* q = plane + base;
    if (acc[0] != 0.f) atomicAdd(q, acc[0]);
    if (acc[1] != 0.f) atomicAdd(q + o2, acc[1]);
    if (acc[2] != 0.f) atomicAdd(q + o1, acc[2]);
    if (acc[3] != 0.f) atomicAdd(q + o1 + o2, acc[3]);
    if (acc[4] != 0.f) atomicAdd(q + o0, acc[4]);
    if (acc[5] != 0.f) atomicAdd(q + o0 + o2, acc[5]);
    if (acc[6] != 0.f) atomicAdd(q + o0 + o1, acc[6]);
    if (acc[7] != 0.f) atomicAdd(q + o0 + o1 + o2, acc[7]);
}

// grad_feat with the same ownership: a thread sums the contributions of its four points per corner in registers while the voxel stays
// the same; what is left at the end of the quad (its last run) is first merged with the neighbouring lanes that ended in the same
// voxel (segmented warp scan), then one lane per run issues the shared-memory atomics.
__global__ void __launch_bounds__(256) devox_scatter4_kernel(DevoxArgs a, const float* __restrict__ gout, long long ob,
                                                            float* __restrict__ grad_feat, int flush_atomic) {
    extern __shared__ __align__(128) float s_acc[];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * a.cc;
    const int ncc = min(a.cc, a.C - c0);
    const int n0 = blockIdx.x * a.pts_per_cta, n1 = min(a.N, n0 + a.pts_per_cta);
    for (int i = threadIdx.x; i < ncc * a.R3; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int span = n1 - n0;
    for (int q0 = 0; 4 * q0 < span; q0 += blockDim.x) {      // warp-uniform trip count: the scan below needs every lane
        const int q = q0 + threadIdx.x;
        const int n = n0 + 4 * q;
        const bool active = 4 * q < span;
        const int nvalid = active ? min(4, n1 - n) : 0;
        DvQuad Q;
        float w[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            DvPoint p = dv_point(a, b, min(n + i, n1 - 1));
            Q.base[i] = p.base; Q.hib[i] = dv_hib(p);
            dv_weights(p, w[i]);
            if (i >= nvalid) {
#pragma unroll
                for (int k = 0; k < 8; ++k) w[i][k] = 0.f;
            }
        }
        const int lastb = active ? Q.base[3] : -2 - lane;
        const int prev = __shfl_up_sync(0xffffffffu, lastb, 1);
        const bool head = (lane == 0) || (prev != lastb);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const bool merge = heads != 0xffffffffu;
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
        const float* gsrc = gout + (size_t)b * ob + (size_t)c0 * a.N + n;
        for (int j = 0; j < ncc; ++j) {
            float* plane = s_acc + (size_t)j * a.R3;
            float go[4], acc[8];
            if (active) dv_load_row(gsrc + (size_t)j * a.N, go, nvalid);
            else go[0] = go[1] = go[2] = go[3] = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.f;
            int cur = Q.base[0], hcur = Q.hib[0];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (Q.base[i] != cur) {            // the voxel changed inside the quad: this run is complete
                    dv_flush_acc(plane, cur, hcur, a.R, acc);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
                    cur = Q.base[i]; hcur = Q.hib[i];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(w[i][k], go[i], acc[k]);
            }
            __syncwarp();
            if (merge) {
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float t = __shfl_up_sync(0xffffffffu, acc[k], dd);
                        if (lane - dd >= start) acc[k] += t;
                    }
                }
            }
            if (active && tail) dv_flush_acc(plane, cur, hcur, a.R, acc);
        }
    }
    __syncthreads();
    float* dst = grad_feat + ((size_t)b * a.C + c0) * a.R3;
    if (flush_atomic) {
        for (int i = threadIdx.x; i < ncc * a.R3; i += blockDim.x) { float v = s_acc[i]; if (v != 0.f) atomicAdd(dst + i, v); }
    } else {
        for (int i = threadIdx.x; i < ncc * a.R3; i += blockDim.x) dst[i] = s_acc[i];
    }
}

// ---- grad_feat of small volumes: a warp owns 32 channels ----------------------------------------------------------------------------
// The kernel above is bound by the shared-memory atomics and the scan that thins them out (ncu, 512 channels at R = 8: 1.3 G warp
// instructions for 187 M (sample, channel, point) elements, 3.8 ms where the gradient rows cross HBM in 0.12 ms).  Here the ownership is
// turned round: lane = channel, and the CTA keeps a private accumulation tile acc[voxel][33] in shared memory, so adding a corner is a
// plain conflict-free read-modify-write -- no atomics, no scan.  The tile (66 KB at R = 8) allows three CTAs per SM, so the consumer
// warp that owns it must not wait for anything else:
//   * a PRODUCER warp works out the records of the next 16 points (8 weights, 8 corner slots in the tile; a corner that does not
//     exist -- upper neighbour beyond the border, weight 0 -- points at a dummy row so that the flush has no branches) and hands
//     them over through a two-slot ring guarded by named barriers;
//   * the CONSUMER streams its own gradient row: 16 consecutive points = 64 contiguous bytes per lane and tile, cp.async'ed three
//     tiles deep into a lane-private, XOR-swizzled strip; it reads the records of four points into registers at once (the compiler
//     cannot hoist those loads over the tile's read-modify-writes by itself), accumulates runs of points in the same voxel in 8
//     registers and touches the tile once per run.
// (First version, one warp, loop over a register tile fully unrolled: 4.6 k instructions, `stalled_no_instruction` 1.3 per issue;
//  second, one warp with a real loop: every instruction waited ~6 cycles on its predecessor -- 295 cycles per point.)
static const int DV_OWN_LD = 33;        // accumulation tile row: 32 channels + 1 (the transposed write-out is conflict-free too)
static const int DV_OWN_TP = 16;        // points per tile
static const int DV_OWN_NBUF = 3;       // gradient tiles in flight
static const int DV_OWN_RECV = 4;       // float4 per point record: 8 weights, 8 corner slots (slot 0 < 0: no point)
static const int DV_OWN_MAXVOX = 512;
enum { DV_BAR_FULL = 1, DV_BAR_EMPTY = 3 };   // named barriers 1,2 / 3,4 (0 is __syncthreads)

__device__ __forceinline__ void dv_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void dv_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// lane-private strip of 16 floats; 16-byte chunk i of lane l sits at chunk i ^ ((l >> 1) & 3): LDS.128 of a quarter-warp conflict-free
__device__ __forceinline__ void dv_own_issue(const float* __restrict__ src, int cnt, float* strip, int swz) {
    if (cnt == DV_OWN_TP && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const unsigned d = (unsigned)__cvta_generic_to_shared(strip);
#pragma unroll
        for (int i = 0; i < DV_OWN_TP / 4; ++i)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16 * (i ^ swz)), "l"(src + 4 * i) : "memory");
    } else if (cnt > 0) {
#pragma unroll
        for (int i = 0; i < DV_OWN_TP; ++i) strip[4 * ((i >> 2) ^ swz) + (i & 3)] = i < cnt ? src[i] : 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");      // one group per tile, empty or not: the wait below counts groups
}

__global__ void __launch_bounds__(64) devox_scatter_owner_kernel(DevoxArgs a, const float* __restrict__ gout, long long ob,
                                                                 float* __restrict__ grad_feat, int flush_atomic) {
    extern __shared__ __align__(16) float s_own[];
    float4* rec = reinterpret_cast<float4*>(s_own);                               // [2][16][4]
    float* gbuf = s_own + 2 * DV_OWN_TP * DV_OWN_RECV * 4;                        // [3][32][16]
    float* acc = gbuf + DV_OWN_NBUF * 32 * DV_OWN_TP;                             // [R3 + 1][33], the last row is the dummy
    const int lane = threadIdx.x & 31;
    const bool producer = threadIdx.x >= 32;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * 32;
    const int ncc = min(32, a.C - c0);
    const int dummy = a.R3 * DV_OWN_LD;
    for (int i = threadIdx.x; i < (a.R3 + 1) * DV_OWN_LD; i += 64) acc[i] = 0.f;
    const int n0 = blockIdx.x * a.pts_per_cta, n1 = min(a.N, n0 + a.pts_per_cta);
    const int ntiles = (n1 - n0 + DV_OWN_TP - 1) / DV_OWN_TP;
    __syncthreads();

    if (producer) {
        for (int t = 0; t < ntiles; ++t) {
            const int buf = t & 1;
            if (t >= 2) dv_bar_sync(DV_BAR_EMPTY + buf);                              // the consumer is done with tile t - 2
            if (lane < DV_OWN_TP) {
                const int n = n0 + DV_OWN_TP * t + lane;
                const bool valid = n < n1;
                DvPoint p = dv_point(a, b, valid ? n : n1 - 1);
                float w[8];
                int o[8], sl[8];
                dv_weights(p, w);
                dv_offsets(p, o);
                const int h = dv_hib(p);
#pragma unroll
                for (int k = 0; k < 8; ++k) sl[k] = ((h & k) == k) ? (p.base + o[k]) * DV_OWN_LD : dummy;   // corner k needs the upper neighbours in bits k
                if (!valid) sl[0] = -1;
                float4* r = rec + (buf * DV_OWN_TP + lane) * DV_OWN_RECV;
                r[0] = make_float4(w[0], w[1], w[2], w[3]);
                r[1] = make_float4(w[4], w[5], w[6], w[7]);
                r[2] = make_float4(__int_as_float(sl[0]), __int_as_float(sl[1]), __int_as_float(sl[2]), __int_as_float(sl[3]));
                r[3] = make_float4(__int_as_float(sl[4]), __int_as_float(sl[5]), __int_as_float(sl[6]), __int_as_float(sl[7]));
            }
            __threadfence_block();
            dv_bar_arrive(DV_BAR_FULL + buf);
        }
    } else {
        const float* grow = gout + (size_t)b * ob + (size_t)(c0 + min(lane, ncc - 1)) * a.N + n0;   // idle lanes shadow the last row
        float* gmine = gbuf + lane * DV_OWN_TP;
        float* accl = acc + lane;
        const int swz = (lane >> 1) & 3;
#pragma unroll
        for (int t = 0; t < DV_OWN_NBUF - 1; ++t) dv_own_issue(grow + DV_OWN_TP * t, min(DV_OWN_TP, n1 - n0 - DV_OWN_TP * t), gmine + t * 32 * DV_OWN_TP, swz);
        float s[8];
        int slot[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { s[k] = 0.f; slot[k] = dummy; }
        int cur = -1;
        for (int t = 0; t < ntiles; ++t) {
            const int tn = t + DV_OWN_NBUF - 1;
            dv_own_issue(grow + DV_OWN_TP * tn, min(DV_OWN_TP, n1 - n0 - DV_OWN_TP * tn), gmine + (tn % DV_OWN_NBUF) * 32 * DV_OWN_TP, swz);
            asm volatile("cp.async.wait_group %0;" ::"n"(DV_OWN_NBUF - 1) : "memory");   // this lane's strip of tile t has landed
            const int buf = t & 1;
            dv_bar_sync(DV_BAR_FULL + buf);                                           // ... and so have the records
            const float* gt = gmine + (t % DV_OWN_NBUF) * 32 * DV_OWN_TP;
            const float4* rt = rec + buf * DV_OWN_TP * DV_OWN_RECV;
#pragma unroll 1
            for (int j = 0; j < DV_OWN_TP / 4; ++j) {
                const float4 gv = *reinterpret_cast<const float4*>(gt + 4 * (j ^ swz));
                float4 rr[4][DV_OWN_RECV];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int q = 0; q < DV_OWN_RECV; ++q) rr[i][q] = rt[(4 * j + i) * DV_OWN_RECV + q];
                const float g4[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int lb = __float_as_int(rr[i][2].x);                        // slot of the low corner identifies the voxel
                    if (lb >= 0) {                                                    // else past the end of the range (warp-uniform)
                        if (lb != cur) {
                            // flush the finished run: eight independent read-modify-writes, missing corners land on the dummy row
                            float v[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) v[k] = accl[slot[k]];
#pragma unroll
                            for (int k = 0; k < 8; ++k) accl[slot[k]] = v[k] + s[k];
                            slot[0] = lb; slot[1] = __float_as_int(rr[i][2].y); slot[2] = __float_as_int(rr[i][2].z); slot[3] = __float_as_int(rr[i][2].w);
                            slot[4] = __float_as_int(rr[i][3].x); slot[5] = __float_as_int(rr[i][3].y);
                            slot[6] = __float_as_int(rr[i][3].z); slot[7] = __float_as_int(rr[i][3].w);
#pragma unroll
                            for (int k = 0; k < 8; ++k) s[k] = 0.f;
                            cur = lb;
                        }
                        const float g = g4[i];
                        s[0] = fmaf(rr[i][0].x, g, s[0]); s[1] = fmaf(rr[i][0].y, g, s[1]); s[2] = fmaf(rr[i][0].z, g, s[2]); s[3] = fmaf(rr[i][0].w, g, s[3]);
                        s[4] = fmaf(rr[i][1].x, g, s[4]); s[5] = fmaf(rr[i][1].y, g, s[5]); s[6] = fmaf(rr[i][1].z, g, s[6]); s[7] = fmaf(rr[i][1].w, g, s[7]);
                    }
                }
            }
            if (t + 2 < ntiles) dv_bar_arrive(DV_BAR_EMPTY + buf);                    // the records of tile t may be overwritten
        }
        {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = accl[slot[k]];
#pragma unroll
            for (int k = 0; k < 8; ++k) accl[slot[k]] = v[k] + s[k];
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    // transposed write-out: lanes run over the voxels of one channel row (128-byte reductions / stores)
    for (int j = threadIdx.x >> 5; j < ncc; j += 2) {
        float* dst = grad_feat + ((size_t)b * a.C + c0 + j) * a.R3;
        for (int v = lane; v < a.R3; v += 32) {
            const float val = acc[v * DV_OWN_LD + j];
            if (flush_atomic) { if (val != 0.f) atomicAdd(dst + v, val); }
            else dst[v] = val;
        }
    }
}

// ---- grad_feat from points sorted by voxel -------------------------------------------------------------------------------------------
// With a workspace the reduction needs neither atomics on shared memory nor an accumulation tile: the points of all samples are
// radix-sorted once by (sample, low-corner voxel) -- stable, so a voxel lists its points in ascending index -- and a warp then walks a
// block of 128 sorted points for 32 channels (lane = channel), keeps the 8 corner sums of the current voxel in registers and sends them
// to HBM as reductions when the voxel changes: 8 REDs per (voxel, 32 channels) instead of 8 shared-memory read-modify-writes per run of
// the unsorted order.  No shared tile means full occupancy, which is what the channel-owner kernel lacks (three warps per SM).
// The gradient values of a tile of 32 points are fetched with lane = point (consecutive points of a voxel are mostly consecutive in
// memory: a few sectors per request) and turned round through a 32 x 33 shared tile; the point records are worked out by lane = point too.
static const int DV_SR_WARPS = 8;
static const int DV_SR_PTS = 128;      // sorted points per warp

__global__ void __launch_bounds__(256) devox_sort_keys_kernel(DevoxArgs a, long long total, unsigned long long* __restrict__ keys,
                                                             unsigned* __restrict__ vals) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = (int)(i / a.N), n = (int)(i - (long long)b * a.N);
    const DvPoint p = dv_point(a, b, n);
    keys[i] = (unsigned long long)b * (unsigned)a.R3 + (unsigned)p.base;
    vals[i] = (unsigned)i;
}

__global__ void __launch_bounds__(32 * DV_SR_WARPS) devox_sorted_reduce_kernel(DevoxArgs a, const unsigned long long* __restrict__ keys,
                                                                               const unsigned* __restrict__ vals, long long total,
                                                                               const float* __restrict__ gout, long long ob,
                                                                               float* __restrict__ grad_feat) {
    __shared__ float s_g[DV_SR_WARPS][32][33];
    __shared__ __align__(16) float s_w[DV_SR_WARPS][32][8];
    __shared__ int s_key[DV_SR_WARPS][32];
    __shared__ int s_hib[DV_SR_WARPS][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long i0 = ((long long)blockIdx.x * DV_SR_WARPS + wid) * DV_SR_PTS;
    if (i0 >= total) return;                                   // warps are independent: no CTA-wide barrier below
    const long long i1 = min(total, i0 + (long long)DV_SR_PTS);
    const int c0 = blockIdx.y * 32;
    const int ncc = min(32, a.C - c0);
    float (*tg)[33] = s_g[wid];
    const float4* tw = reinterpret_cast<const float4*>(s_w[wid]);
    const int* tk = s_key[wid];
    const int* th = s_hib[wid];
    const int RR = a.R * a.R;
    float s[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] = 0.f;
    int cur = -1, hcur = 0;

    auto flush = [&]() {
        const int b = cur / a.R3;
        const int cell = cur - b * a.R3;
        if (lane < ncc) {
            float* dst = grad_feat + ((size_t)b * a.C + c0 + lane) * a.R3 + cell;
            atomicAdd(dst, s[0]);
            if (hcur & 1) atomicAdd(dst + 1, s[1]);
            if (hcur & 2) atomicAdd(dst + a.R, s[2]);
            if ((hcur & 3) == 3) atomicAdd(dst + a.R + 1, s[3]);
            if (hcur & 4) {
                atomicAdd(dst + RR, s[4]);
                if (hcur & 1) atomicAdd(dst + RR + 1, s[5]);
                if (hcur & 2) atomicAdd(dst + RR + a.R, s[6]);
                if ((hcur & 3) == 3) atomicAdd(dst + RR + a.R + 1, s[7]);
            }
        }
    };

    for (long long it = i0; it < i1; it += 32) {
        const long long i = it + lane;
        int key = -1;
        if (i < i1) {
            key = (int)keys[i];
            const unsigned v = vals[i];
            const int b = (int)(v / (unsigned)a.N), n = (int)(v - (unsigned)b * (unsigned)a.N);
            const DvPoint p = dv_point(a, b, n);
            float w[8];
            dv_weights(p, w);
            float4* wd = reinterpret_cast<float4*>(s_w[wid][lane]);
            wd[0] = make_float4(w[0], w[1], w[2], w[3]);
            wd[1] = make_float4(w[4], w[5], w[6], w[7]);
            s_hib[wid][lane] = dv_hib(p);
            const float* src = gout + (size_t)b * ob + (size_t)c0 * a.N + n;
            if (ncc == 32) {
#pragma unroll
                for (int c = 0; c < 32; ++c) tg[lane][c] = src[(size_t)c * a.N];
            } else {
                for (int c = 0; c < ncc; ++c) tg[lane][c] = src[(size_t)c * a.N];
            }
        }
        s_key[wid][lane] = key;
        __syncwarp();
        const int cnt = (int)min(32LL, i1 - it);
#pragma unroll 4
        for (int p = 0; p < cnt; ++p) {
            const int k = tk[p];
            if (k != cur) {                                    // warp-uniform
                if (cur >= 0) flush();
#pragma unroll
                for (int q = 0; q < 8; ++q) s[q] = 0.f;
                cur = k; hcur = th[p];
            }
            const float4 wa = tw[2 * p], wb = tw[2 * p + 1];
            const float g = tg[p][lane];
            s[0] = fmaf(wa.x, g, s[0]); s[1] = fmaf(wa.y, g, s[1]); s[2] = fmaf(wa.z, g, s[2]); s[3] = fmaf(wa.w, g, s[3]);
            s[4] = fmaf(wb.x, g, s[4]); s[5] = fmaf(wb.y, g, s[5]); s[6] = fmaf(wb.z, g, s[6]); s[7] = fmaf(wb.w, g, s[7]);
        }
        __syncwarp();
    }
    if (cur >= 0) flush();
}

// ---- planes too large for shared memory: gather / scatter through L2 ------------------------------------------------------------
// one thread per (b, n), channels in the loop; mode 0 forward, 1 grad coords, 2 grad feat (global reductions)
template <int MODE>
__global__ void __launch_bounds__(256) devox_global_kernel(DevoxArgs a, float* __restrict__ out, const float* __restrict__ gout, long long ob) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.N) return;
    DvPoint p = dv_point(a, b, n);
    float w[8];
    int o[8];
    dv_weights(p, w);
    dv_offsets(p, o);
    const float* fb = a.feat + (size_t)b * a.C * a.R3 + p.base;
    if (MODE == 0) {
        float* dst = out + (size_t)b * ob + n;
        for (int c = 0; c < a.C; ++c) {
            const float* q = fb + (size_t)c * a.R3;
            float acc = __ldg(q) * w[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) acc = fmaf(__ldg(q + o[k]), w[k], acc);
            dst[(size_t)c * a.N] = acc;
        }
    } else if (MODE == 1) {
        if (p.m0 == 0.f && p.m1 == 0.f && p.m2 == 0.f) return;
        const float gg = p.g2 * p.g1, fg = p.f2 * p.g1, gf = p.g2 * p.f1, ff = p.f2 * p.f1;
        const float g2g0 = p.g2 * p.g0, f2g0 = p.f2 * p.g0, g2f0 = p.g2 * p.f0, f2f0 = p.f2 * p.f0;
        const float g1g0 = p.g1 * p.g0, f1g0 = p.f1 * p.g0, g1f0 = p.g1 * p.f0, f1f0 = p.f1 * p.f0;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        const float* gsrc = gout + (size_t)b * ob + n;
        for (int c = 0; c < a.C; ++c) {
            const float* q = fb + (size_t)c * a.R3;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldg(q + o[k]);
            float go = gsrc[(size_t)c * a.N];
            float e0 = (v[4] - v[0]) * gg + (v[5] - v[1]) * fg + (v[6] - v[2]) * gf + (v[7] - v[3]) * ff;
            float e1 = (v[2] - v[0]) * g2g0 + (v[3] - v[1]) * f2g0 + (v[6] - v[4]) * g2f0 + (v[7] - v[5]) * f2f0;
            float e2 = (v[1] - v[0]) * g1g0 + (v[3] - v[2]) * f1g0 + (v[5] - v[4]) * g1f0 + (v[7] - v[6]) * f1f0;
            d0 = fmaf(go, e0, d0); d1 = fmaf(go, e1, d1); d2 = fmaf(go, e2, d2);
        }
        float* gc = out + (size_t)b * a.cs_b + (size_t)n * a.cs_n;
        if (p.m0 != 0.f) atomicAdd(gc, p.m0 * d0);
        if (p.m1 != 0.f) atomicAdd(gc + a.cs_k, p.m1 * d1);
        if (p.m2 != 0.f) atomicAdd(gc + 2 * a.cs_k, p.m2 * d2);
    } else {
        float* dst = out + (size_t)b * a.C * a.R3 + p.base;      // grad_feat, zeroed by the entry point
        const float* gsrc = gout + (size_t)b * ob + n;
        for (int c = 0; c < a.C; ++c) {
            float go = gsrc[(size_t)c * a.N];
            float* q = dst + (size_t)c * a.R3;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (w[k] != 0.f) atomicAdd(q + o[k], w[k] * go);
        }
    }
}

// ---- launch geometry --------------------------------------------------------------------------------------------------------------
struct DevoxPlan { int vec, cc, chunks, threads, nsplit, pts_per_cta; size_t smem; bool global; };

static const size_t DV_SMEM_SMALL = 64 * 1024;        // several CTAs per SM
static const size_t DV_SMEM_MAX = 220 * 1024;         // one CTA per SM (R = 32: one 128 KB plane)
enum DevoxKind { DV_GATHER1 = 0, DV_SCATTER1, DV_GATHER4, DV_SCATTER4 };

// Channels per CTA, CTA size and how many point ranges the N points are cut into.
//   one point per thread (the DTB_DEVOX_SIMPLE kernels): 64 KB of planes, 512 threads, several CTAs per SM;
//   gather, four points per thread: ~110 registers -> one 512-thread CTA per SM, so it takes all the shared memory it can use
//     (R = 8: 88 channels, R = 16: 12, R = 32: 1) and amortises the per-point setup over them;
//   scatter, four points per thread: 256 threads, 64 KB of accumulation planes, three CTAs per SM.
// The number of ranges minimises  waves * (points per range + staging)  over the SM slots: with 48..512 (sample, chunk) units a fixed
// "4 CTAs per SM" rule left the last wave half empty (ncu: 640 CTAs on 444 slots at R = 8).
static DevoxPlan devox_plan(int B, int C, int N, int R, int flags, DevoxKind kind) {
    DevoxPlan pl{};
    const size_t plane = (size_t)R * R * R * 4;
    pl.global = (flags & DTB_DEVOX_GLOBAL_GATHER) || plane > DV_SMEM_MAX;
    if (pl.global) return pl;
    const bool gather = (kind == DV_GATHER1 || kind == DV_GATHER4);
    const size_t want_budget = (kind == DV_GATHER4) ? DV_SMEM_MAX : DV_SMEM_SMALL;
    pl.vec = (gather && C % 4 == 0 && 4 * plane <= want_budget) ? 4 : 1;
    const size_t budget = ((size_t)pl.vec * plane <= want_budget) ? want_budget : DV_SMEM_MAX;
    int cc_max = (int)(budget / plane);
    cc_max = cc_max / pl.vec * pl.vec;
    if (cc_max > C) cc_max = C;
    const int chunks = cdiv(C, cc_max);
    int cc = cdiv(C, chunks);
    cc = cdiv(cc, pl.vec) * pl.vec;                      // balanced chunks, still a multiple of VEC and <= cc_max
    pl.cc = cc;
    pl.chunks = cdiv(C, cc);
    pl.smem = (size_t)cc * plane;
    int per_thread, ctas_per_sm;
    if (kind == DV_GATHER4) { pl.threads = 512; per_thread = 4; ctas_per_sm = 1; }
    else if (kind == DV_SCATTER4) { pl.threads = 256; per_thread = 4; ctas_per_sm = 3; }
    else { pl.threads = pl.smem > 100 * 1024 ? 1024 : 512; per_thread = 1; ctas_per_sm = pl.threads == 1024 ? 2 : 4; }
    const int by_smem = (int)((227 * 1024) / (pl.smem + 1024));
    if (ctas_per_sm > by_smem) ctas_per_sm = by_smem < 1 ? 1 : by_smem;
    const long long slots = (long long)DTB_SM_COUNT * ctas_per_sm;
    const long long units = (long long)pl.chunks * B;
    const int quantum = pl.threads * per_thread;          // points one pass of the CTA covers
    int most = N / quantum;
    if (most < 1) most = 1;
    if (most > 4096) most = 4096;
    const double stage = (double)R * R * R / 64.0;        // staging / zeroing + flushing, in units of one point's work
    double best = 0.0;
    int best_ns = 1;
    for (int ns = 1; ns <= most; ++ns) {
        const long long waves = (units * ns + slots - 1) / slots;
        const double cost = (double)waves * ((double)cdiv(N, ns) + stage);
        if (ns == 1 || cost < best * 0.999) { best = cost; best_ns = ns; }
    }
    pl.pts_per_cta = cdiv(cdiv(N, best_ns), 4) * 4;       // ranges start on a multiple of four points
    pl.nsplit = cdiv(N, pl.pts_per_cta);
    return pl;
}

// Geometry of devox_scatter_owner_kernel: point ranges so that the two-warp CTAs (three per SM by shared memory at R = 8) come in
// full waves.
struct DevoxOwnerPlan { bool ok; int chunks, nsplit, pts_per_cta; size_t smem; };

static DevoxOwnerPlan devox_owner_plan(int B, int C, int N, int R, int flags) {
    DevoxOwnerPlan pl{};
    const long long R3 = (long long)R * R * R;
    if ((flags & (DTB_DEVOX_SIMPLE | DTB_DEVOX_GLOBAL_GATHER | DTB_DEVOX_NO_OWNER)) || R3 > DV_OWN_MAXVOX) return pl;
    pl.smem = (size_t)((R3 + 1) * DV_OWN_LD + 2 * DV_OWN_TP * DV_OWN_RECV * 4 + DV_OWN_NBUF * 32 * DV_OWN_TP) * 4;
    pl.chunks = cdiv(C, 32);
    if (pl.chunks > 65535) return pl;
    long long per_sm = (long long)(228 * 1024) / (long long)(pl.smem + 1024);
    if (per_sm > 16) per_sm = 16;
    const long long slots = (long long)DTB_SM_COUNT * per_sm;
    const long long units = (long long)pl.chunks * B;
    int most = N / 256;
    if (most < 1) most = 1;
    if (most > 4096) most = 4096;
    const double stage = 64.0 + (double)R3 / 8.0;           // zeroing + write-out of the tile, in units of one point's work
    double best = 0.0;
    int best_ns = 1;
    for (int ns = 1; ns <= most; ++ns) {
        const long long waves = (units * ns + slots - 1) / slots;
        const double cost = (double)waves * ((double)cdiv(N, ns) + stage);
        if (ns == 1 || cost < best * 0.999) { best = cost; best_ns = ns; }
    }
    pl.pts_per_cta = cdiv(cdiv(N, best_ns), DV_OWN_TP) * DV_OWN_TP;   // ranges start on a tile boundary: 16-byte aligned rows when N % 4 == 0
    pl.nsplit = cdiv(N, pl.pts_per_cta);
    pl.ok = true;
    return pl;
}

// The sorted reduction pays off when a voxel collects several points (its 8 reductions per 32 channels are amortised over them).
static bool devox_sorted_ok(int B, int C, int N, int R, int flags) {
    if (flags & (DTB_DEVOX_SIMPLE | DTB_DEVOX_GLOBAL_GATHER | DTB_DEVOX_NO_SORT)) return false;
    const long long R3 = (long long)R * R * R, total = (long long)B * N;
    if (total <= 0 || C <= 0 || total >= (1LL << 31) || (long long)B * R3 >= (1LL << 31)) return false;
    return (flags & DTB_DEVOX_FORCE_SORT) || (long long)N >= 8 * R3;
}

struct DevoxSortBufs { unsigned long long *keys_in, *keys_out; unsigned *vals_in, *vals_out; void* sort_ws; size_t sort_ws_bytes; };

static size_t devox_sorted_carve(Workspace& ws, long long total, DevoxSortBufs& sb) {
    sb.keys_in = ws.take<unsigned long long>((size_t)total);
    sb.keys_out = ws.take<unsigned long long>((size_t)total);
    sb.vals_in = ws.take<unsigned>((size_t)total);
    sb.vals_out = ws.take<unsigned>((size_t)total);
    sb.sort_ws_bytes = sort_workspace_bytes((size_t)total);
    sb.sort_ws = ws.take<char>(sb.sort_ws_bytes);
    return ws.off;
}

template <typename K>
static int devox_allow_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024) DTB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DV_SMEM_MAX));
    return DTB_OK;
}

static int devox_check(const float* feat, const float* coords, int B, int C, int N, int R, const char* what) {
    DTB_REQUIRE(B >= 0 && C >= 0 && N >= 0 && R >= 1, "%s: bad sizes B=%d C=%d N=%d R=%d", what, B, C, N, R);
    DTB_REQUIRE(R <= 1024, "%s: resolution %d too large", what, R);
    DTB_REQUIRE(B <= 65535 && C <= 65535 * 4, "%s: batch/channels exceed the grid limits", what);
    if ((long long)B * C * N == 0) return 1;
    DTB_REQUIRE(feat && coords, "%s: null argument", what);
    return DTB_OK;
}

}  // namespace dtb

using namespace dtb;

static DevoxArgs devox_args(const float* feat, const float* coords, long long cs_b, long long cs_k, long long cs_n, int C, int N, int R,
                            int flags) {
    DevoxArgs a{};
    a.feat = feat; a.coords = coords; a.cs_b = cs_b; a.cs_k = cs_k; a.cs_n = cs_n;
    a.C = C; a.N = N; a.R = R; a.R3 = R * R * R;
    a.from_pos = (flags & DTB_DEVOX_FROM_POSITIONS) ? 1 : 0;
    a.inv_r = 1.0f / (float)R;
    a.pow2 = (R & (R - 1)) == 0;
    return a;
}

extern "C" int dtb_trilinear_devoxelize_forward(const float* feat, const float* coords, long long cs_b, long long cs_k, long long cs_n,
                                                int B, int C, int N, int R, int flags, float* out, long long out_batch_stride,
                                                void* stream) {
    int rc = devox_check(feat, coords, B, C, N, R, "trilinear_devoxelize_forward");
    if (rc < 0) return rc;
    if (rc == 1) return DTB_OK;
    DTB_REQUIRE(out && out_batch_stride >= (long long)C * N, "trilinear_devoxelize_forward: bad output");
    cudaStream_t st = (cudaStream_t)stream;
    DevoxArgs a = devox_args(feat, coords, cs_b, cs_k, cs_n, C, N, R, flags);
    const bool simple = (flags & DTB_DEVOX_SIMPLE) != 0;
    DevoxPlan pl = devox_plan(B, C, N, R, flags, simple ? DV_GATHER1 : DV_GATHER4);
    if (pl.global) {
        devox_global_kernel<0><<<dim3(cdiv(N, 256), B), 256, 0, st>>>(a, out, nullptr, out_batch_stride);
        DTB_LAUNCH_CHECK("devox_global_fwd");
        return DTB_OK;
    }
    a.cc = pl.cc; a.pts_per_cta = pl.pts_per_cta;
    dim3 grid(pl.nsplit, pl.chunks, B);
    if (simple) {
        if (pl.vec == 4) {
            if (int e = devox_allow_smem(devox_gather_kernel<4, 0>, pl.smem)) return e;
            devox_gather_kernel<4, 0><<<grid, pl.threads, pl.smem, st>>>(a, out, nullptr, out_batch_stride);
        } else {
            if (int e = devox_allow_smem(devox_gather_kernel<1, 0>, pl.smem)) return e;
            devox_gather_kernel<1, 0><<<grid, pl.threads, pl.smem, st>>>(a, out, nullptr, out_batch_stride);
        }
    } else if (pl.vec == 4) {
        if (int e = devox_allow_smem(devox_gather4_kernel<4>, pl.smem)) return e;
        devox_gather4_kernel<4><<<grid, pl.threads, pl.smem, st>>>(a, out, out_batch_stride);
    } else {
        if (int e = devox_allow_smem(devox_gather4_kernel<1>, pl.smem)) return e;
        devox_gather4_kernel<1><<<grid, pl.threads, pl.smem, st>>>(a, out, out_batch_stride);
    }
    DTB_LAUNCH_CHECK("devox_gather_fwd");
    return DTB_OK;
}

extern "C" size_t dtb_trilinear_devoxelize_backward_workspace(int B, int C, int N, int R, int flags) {
    if (B < 0 || C < 0 || N < 0 || R < 1 || R > 1024 || !devox_sorted_ok(B, C, N, R, flags)) return 0;
    Workspace ws(nullptr, 0);
    DevoxSortBufs sb;
    return devox_sorted_carve(ws, (long long)B * N, sb);
}

extern "C" int dtb_trilinear_devoxelize_backward(const float* feat, const float* coords, long long cs_b, long long cs_k, long long cs_n,
                                                 const float* grad_out, long long grad_out_batch_stride, int B, int C, int N, int R,
                                                 int flags, float* grad_feat, float* grad_coords, void* stream) {
    return dtb_trilinear_devoxelize_backward_ws(feat, coords, cs_b, cs_k, cs_n, grad_out, grad_out_batch_stride, B, C, N, R, flags, grad_feat,
                                                grad_coords, nullptr, 0, stream);
}

extern "C" int dtb_trilinear_devoxelize_backward_ws(const float* feat, const float* coords, long long cs_b, long long cs_k, long long cs_n,
                                                    const float* grad_out, long long grad_out_batch_stride, int B, int C, int N, int R,
                                                    int flags, float* grad_feat, float* grad_coords, void* workspace, size_t workspace_bytes,
                                                    void* stream) {
    DTB_REQUIRE(B >= 0 && C >= 0 && N >= 0 && R >= 1 && R <= 1024, "trilinear_devoxelize_backward: bad sizes B=%d C=%d N=%d R=%d", B, C, N, R);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t R3 = (size_t)R * R * R;
    if ((long long)B * C * N == 0) {                   // no samples: the volume gradient is zero, the coordinate gradient untouched
        if (grad_feat && (size_t)B * C > 0) DTB_CUDA(cudaMemsetAsync(grad_feat, 0, (size_t)B * C * R3 * 4, st));
        return DTB_OK;
    }
    int rc = devox_check(feat ? feat : grad_out, coords, B, C, N, R, "trilinear_devoxelize_backward");
    if (rc < 0) return rc;
    DTB_REQUIRE(grad_out && grad_out_batch_stride >= (long long)C * N, "trilinear_devoxelize_backward: bad grad_out");
    DTB_REQUIRE(!grad_coords || feat, "trilinear_devoxelize_backward: the coordinate gradient needs the volume");
    DevoxArgs a = devox_args(feat, coords, cs_b, cs_k, cs_n, C, N, R, flags);
    const bool simple = (flags & DTB_DEVOX_SIMPLE) != 0;
    if (grad_coords) {
        DevoxPlan pl = devox_plan(B, C, N, R, flags, simple ? DV_GATHER1 : DV_GATHER4);
        if (pl.global) {
            devox_global_kernel<1><<<dim3(cdiv(N, 256), B), 256, 0, st>>>(a, grad_coords, grad_out, grad_out_batch_stride);
        } else {
            a.cc = pl.cc; a.pts_per_cta = pl.pts_per_cta;
            dim3 grid(pl.nsplit, pl.chunks, B);
            if (simple) {
                if (pl.vec == 4) {
                    if (int e = devox_allow_smem(devox_gather_kernel<4, 1>, pl.smem)) return e;
                    devox_gather_kernel<4, 1><<<grid, pl.threads, pl.smem, st>>>(a, grad_coords, grad_out, grad_out_batch_stride);
                } else {
                    if (int e = devox_allow_smem(devox_gather_kernel<1, 1>, pl.smem)) return e;
                    devox_gather_kernel<1, 1><<<grid, pl.threads, pl.smem, st>>>(a, grad_coords, grad_out, grad_out_batch_stride);
                }
            } else if (pl.vec == 4) {
                if (int e = devox_allow_smem(devox_gradcoords4_kernel<4>, pl.smem)) return e;
                devox_gradcoords4_kernel<4><<<grid, pl.threads, pl.smem, st>>>(a, grad_coords, grad_out, grad_out_batch_stride);
            } else {
                if (int e = devox_allow_smem(devox_gradcoords4_kernel<1>, pl.smem)) return e;
                devox_gradcoords4_kernel<1><<<grid, pl.threads, pl.smem, st>>>(a, grad_coords, grad_out, grad_out_batch_stride);
            }
        }
        DTB_LAUNCH_CHECK("devox_grad_coords");
    }
    if (grad_feat) {
        DevoxPlan pl = devox_plan(B, C, N, R, flags, simple ? DV_SCATTER1 : DV_SCATTER4);
        DevoxOwnerPlan op = devox_owner_plan(B, C, N, R, flags);
        Workspace ws(workspace, workspace_bytes);
        DevoxSortBufs sb{};
        const long long total = (long long)B * N;
        const bool sorted = workspace && devox_sorted_ok(B, C, N, R, flags) && devox_sorted_carve(ws, total, sb) <= workspace_bytes && ws.ok;
        if (sorted) {
            DTB_CUDA(cudaMemsetAsync(grad_feat, 0, (size_t)B * C * R3 * 4, st));
            devox_sort_keys_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(a, total, sb.keys_in, sb.vals_in);
            DTB_LAUNCH_CHECK("devox_sort_keys");
            int bits = 1;
            while (((long long)B * (long long)R3 - 1) >> bits) ++bits;
            if (int e = radix_sort_pairs_u64(sb.keys_in, sb.vals_in, sb.keys_out, sb.vals_out, (size_t)total, bits, sb.sort_ws, sb.sort_ws_bytes, st)) return e;
            const long long nblk = cdiv(total, (long long)DV_SR_PTS);
            devox_sorted_reduce_kernel<<<dim3((unsigned)cdiv(nblk, (long long)DV_SR_WARPS), cdiv(C, 32)), 32 * DV_SR_WARPS, 0, st>>>(
                a, sb.keys_out, sb.vals_out, total, grad_out, grad_out_batch_stride, grad_feat);
        } else if (op.ok) {
            a.pts_per_cta = op.pts_per_cta;
            const int flush_atomic = op.nsplit > 1;
            if (flush_atomic) DTB_CUDA(cudaMemsetAsync(grad_feat, 0, (size_t)B * C * R3 * 4, st));
            if (int e = devox_allow_smem(devox_scatter_owner_kernel, op.smem)) return e;
            devox_scatter_owner_kernel<<<dim3(op.nsplit, op.chunks, B), 64, op.smem, st>>>(a, grad_out, grad_out_batch_stride, grad_feat, flush_atomic);
        } else if (pl.global) {
            DTB_CUDA(cudaMemsetAsync(grad_feat, 0, (size_t)B * C * R3 * 4, st));
            devox_global_kernel<2><<<dim3(cdiv(N, 256), B), 256, 0, st>>>(a, grad_feat, grad_out, grad_out_batch_stride);
        } else {
            a.cc = pl.cc; a.pts_per_cta = pl.pts_per_cta;
            const int flush_atomic = pl.nsplit > 1;
            if (flush_atomic) DTB_CUDA(cudaMemsetAsync(grad_feat, 0, (size_t)B * C * R3 * 4, st));
            dim3 grid(pl.nsplit, pl.chunks, B);
            if (simple) {
                if (int e = devox_allow_smem(devox_scatter_kernel, pl.smem)) return e;
                devox_scatter_kernel<<<grid, pl.threads, pl.smem, st>>>(a, grad_out, grad_out_batch_stride, grad_feat, flush_atomic);
            } else {
                if (int e = devox_allow_smem(devox_scatter4_kernel, pl.smem)) return e;
                devox_scatter4_kernel<<<grid, pl.threads, pl.smem, st>>>(a, grad_out, grad_out_batch_stride, grad_feat, flush_atomic);
            }
        }
        DTB_LAUNCH_CHECK("devox_grad_feat");
    }
    return DTB_OK;
}
