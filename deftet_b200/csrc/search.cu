// Binned exact-semantics search kernels:
//   A1  point-in-tet query   (reference layers/DefTet/check_condition_tetrahedron_base/check_condition_tet_for.cu:105-189)
//       + barycentric weights and their backward (reference has none: utils.py:56-58 returns None; the
//         weights follow utils/tet_utils.py:28-45 `bary_centric_tet`, the gradient is its autograd)
//   A2  1-nearest-neighbour  (reference layers/nearest_neighbor/nearest_neighbor_cuda.cu:17-55)
//
// The reference scans all T tets / all M points per query thread (O(P*T), O(Q*M)).  Here the query
// points (A1) or the target points (A2) are counting-sorted into a uniform grid (pointgrid.cu) and
//   A1: every tet visits only the cells its (slightly inflated) bounding box overlaps and publishes
//       itself with atomicMin on the point's slot  -> "first containing tet in ascending id" exactly;
//   A2: every query walks cubic shells of cells outward, pruning with a conservative lower bound, and
//       keeps the lexicographic minimum of (distance, index) -> "strict <, lowest index wins" exactly.
// Predicates and distances use the non-contracted fp32 sequence of the reference source (common.cuh x*),
// so indices are bit-identical to the CPU oracle.
#include <stdlib.h>
#include "brickwalk.cuh"
#include "deftet_b200.h"

namespace dtb {

// ---------------------------------------------------------------------------------------------------
// A1
// ---------------------------------------------------------------------------------------------------
struct FacePlane { float ax, ay, az, nx, ny, nz; bool sv; float dv; };

// normal of (b-a)x(c-a), sign of its dot with (d-a): check_condition_tet_for.cu:105-121
__device__ __forceinline__ FacePlane make_plane(const float* a, const float* b, const float* c, const float* d) {
    float r1x = xsub(b[0], a[0]), r1y = xsub(b[1], a[1]), r1z = xsub(b[2], a[2]);
    float r2x = xsub(c[0], a[0]), r2y = xsub(c[1], a[1]), r2z = xsub(c[2], a[2]);
    FacePlane f;
    f.ax = a[0]; f.ay = a[1]; f.az = a[2];
    f.nx = xsub(xmul(r1y, r2z), xmul(r1z, r2y));
    f.ny = xsub(xmul(r1z, r2x), xmul(r1x, r2z));
    f.nz = xsub(xmul(r1x, r2y), xmul(r1y, r2x));
    float dx = xsub(d[0], a[0]), dy = xsub(d[1], a[1]), dz = xsub(d[2], a[2]);
    float dotv4 = xadd(xadd(xmul(f.nx, dx), xmul(f.ny, dy)), xmul(f.nz, dz));
    f.sv = dotv4 > 0.f;
    f.dv = dotv4;
    return f;
}
__device__ __forceinline__ bool same_side(const FacePlane& f, float px, float py, float pz) {
    float dx = xsub(px, f.ax), dy = xsub(py, f.ay), dz = xsub(pz, f.az);
    float dotp = xadd(xadd(xmul(f.nx, dx), xmul(f.ny, dy)), xmul(f.nz, dz));
    return (dotp > 0.f) == f.sv;
}

struct PitIndexed {
    const float* pos; const int32_t* tet; int V; int T;
    __device__ __forceinline__ void load(int b, int t, float v[4][3]) const {
        int4 id = reinterpret_cast<const int4*>(tet)[t];
        const float* p = pos + (size_t)b * V * 3;
        int ids[4] = {id.x, id.y, id.z, id.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k][0] = __ldg(p + (size_t)ids[k] * 3); v[k][1] = __ldg(p + (size_t)ids[k] * 3 + 1); v[k][2] = __ldg(p + (size_t)ids[k] * 3 + 2);
        }
    }
};
struct PitSoup {
    const float* soup; int T;
    __device__ __forceinline__ void load(int b, int t, float v[4][3]) const {
        const float4* q = reinterpret_cast<const float4*>(soup + ((size_t)b * T + t) * 12);
        float4 x = __ldg(q), y = __ldg(q + 1), z = __ldg(q + 2);
        v[0][0] = x.x; v[0][1] = x.y; v[0][2] = x.z; v[1][0] = x.w; v[1][1] = y.x; v[1][2] = y.y;
        v[2][0] = y.z; v[2][1] = y.w; v[2][2] = z.x; v[3][0] = z.y; v[3][1] = z.z; v[3][2] = z.w;
    }
};

#ifndef PIT_UNROLL
#define PIT_UNROLL 1
#endif
#ifndef PIT_XMULT
#define PIT_XMULT 4            // x refinement of the query-point grid rows
#endif
#ifndef PIT_MIN_CTAS
#define PIT_MIN_CTAS 1
#endif
template <typename Src>
__global__ void __launch_bounds__(128, PIT_MIN_CTAS) pit_tet_kernel(Src src, int T, int P, int G, const unsigned* __restrict__ bbox_ord,
                                                      const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_end,
                                                      const float4* __restrict__ sorted, int* __restrict__ hit, int* __restrict__ weak_list,
                                                      int* __restrict__ n_weak) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    float v[4][3];
    src.load(b, t, v);
    // the four rotations (a,b,c,d),(b,a,d,c),(c,d,a,b),(d,c,b,a) of check_condition_tet_for.cu:172-175
    FacePlane f1 = make_plane(v[0], v[1], v[2], v[3]);
    FacePlane f2 = make_plane(v[1], v[0], v[3], v[2]);
    FacePlane f3 = make_plane(v[2], v[3], v[0], v[1]);
    FacePlane f4 = make_plane(v[3], v[2], v[1], v[0]);
    float mn[3], mx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        mn[k] = fminf(fminf(v[0][k], v[1][k]), fminf(v[2][k], v[3][k]));
        mx[k] = fmaxf(fmaxf(v[0][k], v[1][k]), fmaxf(v[2][k], v[3][k]));
    }
    float ext = fmaxf(fmaxf(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
    // `weak`: the sign of the four dotv4 (each is +-6 x the signed volume, formed from edges no longer than 3 ext in L1) is not
    // certain in fp32 -- zero-volume / sliver tets, NaN or inf vertices.  For such a tet the four predicates can agree for points
    // ANYWHERE (identical vertices: every dot product is 0, both sign tests false, every point accepted,
    // check_condition_tet_for.cu:105-121), so it is not pruned by its bounding box but handed to pit_weak_kernel (brute force).
    if (!(fabsf(f1.dv) > 2.7e-4f * ext * ext * ext)) {
        weak_list[(size_t)b * T + atomicAdd(n_weak + b, 1)] = t;
        return;
    }
    // conservative inflation: the fp32 predicates can accept points a few ulps outside the exact tet
    float mag = fmaxf(fmaxf(fmaxf(fabsf(mn[0]), fabsf(mx[0])), fmaxf(fabsf(mn[1]), fabsf(mx[1]))), fmaxf(fabsf(mn[2]), fabsf(mx[2])));
    float margin = 1e-4f * ext + 1e-5f * mag;
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] -= margin; mx[k] += margin; }
    GridParams g = grid_params(bbox_ord, b, G);
    float gmax = g.h * (float)G;
    if (mx[0] < g.ox || mx[1] < g.oy || mx[2] < g.oz || mn[0] > g.ox + gmax * 1.0001f || mn[1] > g.oy + gmax * 1.0001f ||
        mn[2] > g.oz + gmax * 1.0001f)
        return;
    // rows are (z, y) lines of G * PIT_XMULT cells: the finer x resolution only tightens the [x0, x1] candidate range of a row (still
    // two look-ups per row), it adds no rows
    const int Gx = G * PIT_XMULT;
    int x0 = cell_coord(mn[0], g.ox, g.inv_h * (float)PIT_XMULT, Gx), x1 = cell_coord(mx[0], g.ox, g.inv_h * (float)PIT_XMULT, Gx);
    int y0 = cell_coord(mn[1], g.oy, g.inv_h, G), y1 = cell_coord(mx[1], g.oy, g.inv_h, G);
    int z0 = cell_coord(mn[2], g.oz, g.inv_h, G), z1 = cell_coord(mx[2], g.oz, g.inv_h, G);
    int* hb = hit + (size_t)b * P;
    const size_t cbase = (size_t)b * G * G * Gx;
    for (int z = z0; z <= z1; ++z)
        for (int y = y0; y <= y1; ++y) {
            size_t row = cbase + ((size_t)z * G + y) * Gx;
            unsigned j0 = cell_start[row + x0], j1 = cell_end[row + x1];
#if PIT_UNROLL > 1
            for (unsigned j = j0; j < j1; j += PIT_UNROLL) {              // loads first: PIT_UNROLL candidates in flight per lane
                float4 qs[PIT_UNROLL];
#pragma unroll
                for (int k = 0; k < PIT_UNROLL; ++k) qs[k] = __ldg(sorted + min(j + k, j1 - 1));
#pragma unroll
                for (int k = 0; k < PIT_UNROLL; ++k) {
                    const float4 q = qs[k];
                    if (j + k >= j1 || q.x < mn[0] || q.x > mx[0] || q.y < mn[1] || q.y > mx[1] || q.z < mn[2] || q.z > mx[2]) continue;
                    bool s1 = same_side(f1, q.x, q.y, q.z), s2 = same_side(f2, q.x, q.y, q.z);
                    bool s3 = same_side(f3, q.x, q.y, q.z), s4 = same_side(f4, q.x, q.y, q.z);
                    if (s1 == s2 && s2 == s3 && s3 == s4) atomicMin(hb + __float_as_int(q.w), t);
                }
            }
#else
            for (unsigned j = j0; j < j1; ++j) {
                float4 q = __ldg(sorted + j);
                if (q.x < mn[0] || q.x > mx[0] || q.y < mn[1] || q.y > mx[1] || q.z < mn[2] || q.z > mx[2]) continue;
                bool s1 = same_side(f1, q.x, q.y, q.z), s2 = same_side(f2, q.x, q.y, q.z);
                bool s3 = same_side(f3, q.x, q.y, q.z), s4 = same_side(f4, q.x, q.y, q.z);
                if (s1 == s2 && s2 == s3 && s3 == s4) atomicMin(hb + __float_as_int(q.w), t);
            }
#endif
        }
}

// Tets whose orientation is numerically uncertain (see make_plane) against ALL points of their sample.  Launched with a fixed grid
// behind pit_tet_kernel; returns at once when the list is empty (every well-formed grid).  Points that the binning could not place
// (non-finite coordinates) are covered as well: the loop runs over the original point array.
template <typename Src>
__global__ void __launch_bounds__(256) pit_weak_kernel(Src src, int T, const float* __restrict__ points, int P, const int* __restrict__ weak_list,
                                                       const int* __restrict__ n_weak, int* __restrict__ hit) {
    const int b = blockIdx.y;
    const int n = n_weak[b];
    if (n == 0) return;
    const float* pb = points + (size_t)b * P * 3;
    int* hb = hit + (size_t)b * P;
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const int t = weak_list[(size_t)b * T + e];
        float v[4][3];
        src.load(b, t, v);
        FacePlane f1 = make_plane(v[0], v[1], v[2], v[3]);
        FacePlane f2 = make_plane(v[1], v[0], v[3], v[2]);
        FacePlane f3 = make_plane(v[2], v[3], v[0], v[1]);
        FacePlane f4 = make_plane(v[3], v[2], v[1], v[0]);
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            const float qx = pb[(size_t)i * 3], qy = pb[(size_t)i * 3 + 1], qz = pb[(size_t)i * 3 + 2];
            bool s1 = same_side(f1, qx, qy, qz), s2 = same_side(f2, qx, qy, qz);
            bool s3 = same_side(f3, qx, qy, qz), s4 = same_side(f4, qx, qy, qz);
            if (s1 == s2 && s2 == s3 && s3 == s4) atomicMin(hb + i, t);
        }
    }
}

// bary_centric_tet (utils/tet_utils.py:28-45): ratios of scalar triple products
__device__ __forceinline__ float triple(const float* a, const float* b, const float* c) {
    return a[0] * (b[1] * c[2] - b[2] * c[1]) + a[1] * (b[2] * c[0] - b[0] * c[2]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
}
__device__ __forceinline__ void bary_weights(const float v[4][3], const float* p, float* w) {
    float vap[3], vbp[3], vab[3], vac[3], vad[3], vbc[3], vbd[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        vap[k] = p[k] - v[0][k]; vbp[k] = p[k] - v[1][k];
        vab[k] = v[1][k] - v[0][k]; vac[k] = v[2][k] - v[0][k]; vad[k] = v[3][k] - v[0][k];
        vbc[k] = v[2][k] - v[1][k]; vbd[k] = v[3][k] - v[1][k];
    }
    float v6 = 1.0f / triple(vab, vac, vad);
    w[0] = triple(vbp, vbd, vbc) * v6;
    w[1] = triple(vap, vac, vad) * v6;
    w[2] = triple(vap, vad, vab) * v6;
    w[3] = triple(vap, vab, vac) * v6;
}

template <typename Src>
__global__ void __launch_bounds__(256) pit_finalize_kernel(Src src, int T, const float* __restrict__ points, int P, const int* __restrict__ hit,
                                                           float* __restrict__ cond, float* __restrict__ bary) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    int t = hit[(size_t)b * P + i];
    {
        // a point with a non-finite coordinate cannot be binned, and the reference's predicates treat it in their own way (NaN: both
        // sign tests false -> the first consistently oriented tet "contains" it): answer it with the reference's own serial scan
        const float* p = points + ((size_t)b * P + i) * 3;
        const float s = p[0] + p[1] + p[2];
        if (!(fabsf(s) <= 3.0e38f)) {
            t = 0x7f7f7f7f;
            for (int k = 0; k < T; ++k) {
                float v[4][3];
                src.load(b, k, v);
                FacePlane f1 = make_plane(v[0], v[1], v[2], v[3]), f2 = make_plane(v[1], v[0], v[3], v[2]);
                FacePlane f3 = make_plane(v[2], v[3], v[0], v[1]), f4 = make_plane(v[3], v[2], v[1], v[0]);
                bool s1 = same_side(f1, p[0], p[1], p[2]), s2 = same_side(f2, p[0], p[1], p[2]);
                bool s3 = same_side(f3, p[0], p[1], p[2]), s4 = same_side(f4, p[0], p[1], p[2]);
                if (s1 == s2 && s2 == s3 && s3 == s4) { t = k; break; }
            }
        }
    }
    bool found = t != 0x7f7f7f7f;
    if (cond) cond[(size_t)b * P + i] = found ? (float)t : -1.0f;
    if (bary) {
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        if (found) {
            float v[4][3];
            src.load(b, t, v);
            const float* p = points + ((size_t)b * P + i) * 3;
            float pp[3] = {p[0], p[1], p[2]};
            bary_weights(v, pp, w);
        }
        reinterpret_cast<float4*>(bary)[(size_t)b * P + i] = make_float4(w[0], w[1], w[2], w[3]);
    }
}

// backward of the barycentric weights: with E = [b-a, c-a, d-a] (columns), lambda = E^-1 (p - a),
// h = E^-T (g_b-g_a, g_c-g_a, g_d-g_a):  dL/dp = h,  dL/dv_i = -w_i h  (i = a,b,c,d).
__global__ void __launch_bounds__(256) bary_backward_kernel(const float* __restrict__ pos, const int32_t* __restrict__ tet, int V,
                                                            const float* __restrict__ points, int P, const float* __restrict__ cond,
                                                            const float* __restrict__ g_w, float* __restrict__ grad_pos, int gstride,
                                                            float* __restrict__ grad_points) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    size_t o = (size_t)b * P + i;
    float c = cond[o];
    float h[3] = {0.f, 0.f, 0.f};
    if (c >= 0.f) {
        int t = (int)c;
        int4 id = reinterpret_cast<const int4*>(tet)[t];
        const float* pb = pos + (size_t)b * V * 3;
        int ids[4] = {id.x, id.y, id.z, id.w};
        float v[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k][0] = pb[(size_t)ids[k] * 3]; v[k][1] = pb[(size_t)ids[k] * 3 + 1]; v[k][2] = pb[(size_t)ids[k] * 3 + 2]; }
        float p[3] = {points[o * 3], points[o * 3 + 1], points[o * 3 + 2]};
        float w[4];
        bary_weights(v, p, w);
        float4 g = reinterpret_cast<const float4*>(g_w)[o];
        float gb = g.y - g.x, gc = g.z - g.x, gd = g.w - g.x;
        // rows of E^-1 are (c-a)x(d-a), (d-a)x(b-a), (b-a)x(c-a) over det; h = sum_i gbar_i * row_i
        float e1[3], e2[3], e3[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { e1[k] = v[1][k] - v[0][k]; e2[k] = v[2][k] - v[0][k]; e3[k] = v[3][k] - v[0][k]; }
        float r1[3] = {e2[1] * e3[2] - e2[2] * e3[1], e2[2] * e3[0] - e2[0] * e3[2], e2[0] * e3[1] - e2[1] * e3[0]};
        float r2[3] = {e3[1] * e1[2] - e3[2] * e1[1], e3[2] * e1[0] - e3[0] * e1[2], e3[0] * e1[1] - e3[1] * e1[0]};
        float r3[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        float inv = 1.0f / (e1[0] * r1[0] + e1[1] * r1[1] + e1[2] * r1[2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) h[k] = (gb * r1[k] + gc * r2[k] + gd * r3[k]) * inv;
        if (grad_pos) {
            float* gp = grad_pos + (size_t)b * V * gstride;
#pragma unroll
            for (int q = 0; q < 4; ++q) grad_add3(gp, (size_t)ids[q], gstride, -w[q] * h[0], -w[q] * h[1], -w[q] * h[2]);
        }
    }
    if (grad_points) { grad_points[o * 3] = h[0]; grad_points[o * 3 + 1] = h[1]; grad_points[o * 3 + 2] = h[2]; }
}

// ---------------------------------------------------------------------------------------------------
// A2
// ---------------------------------------------------------------------------------------------------
struct NnVisitor {
    float qx, qy, qz;
    float best;      // nearest_neighbor_cuda.cu:28-29: min_distance = 1e20, min_point = 0
    int bi;
    __device__ __forceinline__ float bound() const { return best; }
    __device__ static __forceinline__ float no_hit() { return 1e20f; }
    __device__ __forceinline__ void item(const float4& p) {
        float dx = xsub(p.x, qx), dy = xsub(p.y, qy), dz = xsub(p.z, qz);
        float d = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
        int idx = __float_as_int(p.w);
        if (d < best || (d == best && idx < bi)) { best = d; bi = idx; }
    }
};

// ---- 1-NN: the queries are counting-sorted by the cell of the TARGET grid they fall in, so that the queries of a
// warp start from the same cell (coherent mask / cell / point loads, similar control flow).

__global__ void __launch_bounds__(256) nn_qbin_count_kernel(const float* __restrict__ queries, int Q, int G, const unsigned* __restrict__ bbox_ord,
                                                            const int32_t* __restrict__ q_counts, int q_mult, unsigned* __restrict__ qcount,
                                                            unsigned* __restrict__ qcell) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q) return;
    if (q_counts && i >= q_counts[b] * q_mult) { qcell[(size_t)b * Q + i] = 0xffffffffu; return; }
    GridParams g = grid_params(bbox_ord, b, G);
    const float* p = queries + ((size_t)b * Q + i) * 3;
    unsigned id = (unsigned)b * G * G * G +
                  cell_index(cell_coord(p[0], g.ox, g.inv_h, G), cell_coord(p[1], g.oy, g.inv_h, G), cell_coord(p[2], g.oz, g.inv_h, G), G, true);
    qcell[(size_t)b * Q + i] = id;
    atomicAdd(qcount + id, 1u);
}
__global__ void __launch_bounds__(256) nn_qbin_fill_kernel(const float* __restrict__ queries, int Q, const unsigned* __restrict__ qcell,
                                                           unsigned* __restrict__ qend, float4* __restrict__ qsorted) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q) return;
    unsigned id = qcell[(size_t)b * Q + i];
    if (id == 0xffffffffu) return;
    const float* p = queries + ((size_t)b * Q + i) * 3;
    unsigned dst = atomicAdd(qend + id, 1u);
    qsorted[dst] = make_float4(p[0], p[1], p[2], __int_as_float(i));
}

// per-thread walk; queries come either in their original order (qsorted == nullptr) or cell-sorted
__global__ void __launch_bounds__(128) nn_query_thread_kernel(const float* __restrict__ queries, int Q, int G, const unsigned* __restrict__ bbox_ord,
                                                              const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_end,
                                                              const float4* __restrict__ sorted, const unsigned long long* __restrict__ mask,
                                                              const float4* __restrict__ qsorted, const unsigned* __restrict__ qstart,
                                                              const unsigned* __restrict__ qend, const int32_t* __restrict__ q_counts, int q_mult,
                                                              int* __restrict__ result) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q) return;
    NnVisitor v{0.f, 0.f, 0.f, 1e20f, 0};
    int orig = i;
    if (qsorted) {
        const size_t cells = (size_t)G * G * G;
        unsigned lo = qstart[(size_t)b * cells], hi = qend[(size_t)b * cells + cells - 1];     // this sample's slice of the sorted queries
        if ((unsigned)i >= hi - lo) return;
        float4 q = qsorted[lo + i];
        v.qx = q.x; v.qy = q.y; v.qz = q.z; orig = __float_as_int(q.w);
    } else {
        if (q_counts && i >= q_counts[b] * q_mult) return;
        const float* qp = queries + ((size_t)b * Q + i) * 3;
        v.qx = qp[0]; v.qy = qp[1]; v.qz = qp[2];
    }
    GridParams g = grid_params(bbox_ord, b, G);
    brick_walk(v.qx, v.qy, v.qz, g, G, 0.0f, cell_start, cell_end, sorted, mask, (size_t)b * G * G * G, v);
    result[(size_t)b * Q + orig] = v.bi;
}

// ---- warp-cooperative walk: one CTA per (sample, brick of the target grid), one warp per chunk of 32 of the brick's queries.
// The queries are sorted by target cell in brick-major order, so the queries of a brick are contiguous and a chunk of 32
// of them covers a few neighbouring cells of that brick.  Every candidate point the chunk needs is loaded ONCE with a
// warp-uniform address and evaluated by all 32 lanes (an extra candidate can only tighten a lane's minimum, never change the
// answer), so the warp runs without divergence where the per-thread walk above kept ~9 of 32 lanes busy:
//   phase 1  the points of the chunk's own home cells (a contiguous range of `sorted`);
//   phase 2  every other occupied cell inside the bounding box of the lanes' search balls (capped at NN_COOP_RCAP cells),
//            visited if ANY lane's ball reaches it (same conservative box bound as brick_walk);
//   fallback lanes whose final ball does not fit into NN_COOP_RCAP cells finish with the per-thread brick_walk, which starts
//            from the minimum found so far.
// Result = lexicographic minimum of (distance, index) over a superset of every point that can attain the minimum: identical
// to the brute-force scan of nearest_neighbor_cuda.cu:17-55.
#ifndef NN_DENSE
#define NN_DENSE 0
#endif
#ifndef NN_COOP_RCAP
#define NN_COOP_RCAP 2.0f
#endif
#ifndef NN_UNROLL
#define NN_UNROLL 8            // A/B on the chamfer op (res 70 b8): 2 -> 0.284, 4 -> 0.288, 8 -> 0.273 ms
#endif
__device__ __forceinline__ void nn_scan_range(unsigned j0, unsigned j1, const float4* __restrict__ sorted, NnVisitor& v) {
    unsigned j = j0;
    for (; j + NN_UNROLL <= j1; j += NN_UNROLL) {         // all loads first: NN_UNROLL candidates in flight per warp
        float4 c[NN_UNROLL];
#pragma unroll
        for (int k = 0; k < NN_UNROLL; ++k) c[k] = __ldg(sorted + j + k);
#pragma unroll
        for (int k = 0; k < NN_UNROLL; ++k) v.item(c[k]);
    }
    for (; j < j1; ++j) v.item(__ldg(sorted + j));
}

// phase 2 + fallback of the cooperative kernels (see nn_query_brick_kernel): all 32 lanes call it together.
// Every lane searches the ball of radius rs = min(its current ball, RCAP cells); a lane that has found nothing yet searches the
// full RCAP ball.  A cell is skipped for a lane only if it lies outside that ball or cannot beat the lane's running minimum; a
// lane is COMPLETE when its final ball fits inside RCAP cells (then every cell that could hold the minimum was inside the region
// and was not skipped); the others finish with the per-thread brick_walk.
// Cells already scanned in phase 1 are skipped: either the local range `home_done` of brick (hbx,hby,hbz), or (by_ballot) every
// cell that is the home cell `hc` of some lane.
__device__ __forceinline__ void nn_coop_finish(float qx, float qy, float qz, bool active, NnVisitor& v, const GridParams& g, int G, float shrink,
                                               size_t cell_base, const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_end,
                                               const float4* __restrict__ sorted, const unsigned long long* __restrict__ mask, bool by_ballot,
                                               unsigned hc, int hbx, int hby, int hbz, unsigned long long home_done) {
    const int NB = G >> 2;
    const float rcap = NN_COOP_RCAP * g.h;
    const float rcap2 = rcap * rcap;
    float rs = rcap;
    if (v.best < 1e20f) rs = fminf(sqrtf(v.best) * 1.0001f + shrink + 0.01f * g.h, rcap);
    const int x0 = max((int)floorf((qx - rs - g.ox) * g.inv_h), 0), x1 = min((int)floorf((qx + rs - g.ox) * g.inv_h), G - 1);
    const int y0 = max((int)floorf((qy - rs - g.oy) * g.inv_h), 0), y1 = min((int)floorf((qy + rs - g.oy) * g.inv_h), G - 1);
    const int z0 = max((int)floorf((qz - rs - g.oz) * g.inv_h), 0), z1 = min((int)floorf((qz + rs - g.oz) * g.inv_h), G - 1);
    const int X0 = __reduce_min_sync(0xffffffffu, x0), X1 = __reduce_max_sync(0xffffffffu, x1);
    const int Y0 = __reduce_min_sync(0xffffffffu, y0), Y1 = __reduce_max_sync(0xffffffffu, y1);
    const int Z0 = __reduce_min_sync(0xffffffffu, z0), Z1 = __reduce_max_sync(0xffffffffu, z1);
    for (int bz = Z0 >> 2; bz <= (Z1 >> 2); ++bz)
        for (int by = Y0 >> 2; by <= (Y1 >> 2); ++by)
            for (int bx = X0 >> 2; bx <= (X1 >> 2); ++bx) {
                const size_t bk = ((size_t)bz * NB + by) * NB + bx;
                unsigned long long m = __ldg(mask + (cell_base >> 6) + bk);
                if (!by_ballot && bx == hbx && by == hby && bz == hbz) m &= ~home_done;
                if (!m) continue;
                // cells of this brick inside the region
                m &= brick_xmask(max(X0 - 4 * bx, 0), min(X1 - 4 * bx, 3)) & brick_ymask(max(Y0 - 4 * by, 0), min(Y1 - 4 * by, 3)) &
                     brick_zmask(max(Z0 - 4 * bz, 0), min(Z1 - 4 * bz, 3));
                const float lx = g.ox + (float)(4 * bx) * g.h, ly = g.oy + (float)(4 * by) * g.h, lz = g.oz + (float)(4 * bz) * g.h;
                const size_t cb = cell_base + bk * 64;
                while (m) {
                    const int k = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    if (by_ballot && __any_sync(0xffffffffu, hc == (unsigned)(bk * 64 + k))) continue;      // scanned in phase 1
                    const float cxl = lx + (float)(k & 3) * g.h, cyl = ly + (float)((k >> 2) & 3) * g.h, czl = lz + (float)(k >> 4) * g.h;
                    const bool need = !(box_dist2(qx, qy, qz, cxl, cyl, czl, g.h, shrink) > fminf(v.best, rcap2));
                    if (!__any_sync(0xffffffffu, need)) continue;
                    nn_scan_range(__ldg(cell_start + cb + k), __ldg(cell_end + cb + k), sorted, v);
                }
            }
    const bool complete = (v.best < 1e20f) && (sqrtf(v.best) * 1.0001f + shrink + 0.01f * g.h <= rcap);
    if (active && !complete)
        brick_walk(qx, qy, qz, g, G, 0.0f, cell_start, cell_end, sorted, mask, cell_base, v);
}

__global__ void __launch_bounds__(128) nn_query_brick_kernel(int Q, int G, const unsigned* __restrict__ bbox_ord,
                                                             const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_end,
                                                             const float4* __restrict__ sorted, const unsigned long long* __restrict__ mask,
                                                             const float4* __restrict__ qsorted, const unsigned* __restrict__ qstart,
                                                             const unsigned* __restrict__ qend, int* __restrict__ result) {
    const int b = blockIdx.y;
    const int NB = G >> 2;
    const int brick = blockIdx.x;
    const size_t cell_base = (size_t)b * G * G * G;
    const size_t c0 = cell_base + (size_t)brick * 64;
    const unsigned q0 = __ldg(qstart + c0), q1 = __ldg(qend + c0 + 63);
    if (q0 >= q1) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bx0 = brick % NB, by0 = (brick / NB) % NB, bz0 = brick / (NB * NB);
    const GridParams g = grid_params(bbox_ord, b, G);
    const float slack = 1e-3f * g.h + 1e-6f * (fabsf(g.ox) + fabsf(g.oy) + fabsf(g.oz) + (float)G * g.h);   // as brick_walk (inflate = 0)
    const float shrink = slack;
    for (unsigned base = q0 + warp * 32u; base < q1; base += 4u * 32u) {
        const unsigned i = base + lane;
        const bool active = i < q1;
        float4 q = __ldg(qsorted + (active ? i : q1 - 1));
        NnVisitor v{q.x, q.y, q.z, 1e20f, 0};
        const int cx = cell_coord(q.x, g.ox, g.inv_h, G), cy = cell_coord(q.y, g.oy, g.inv_h, G), cz = cell_coord(q.z, g.oz, g.inv_h, G);
        const int hloc = ((cz & 3) << 4) | ((cy & 3) << 2) | (cx & 3);            // home cell inside the brick (an inactive lane copies the last query)
        // ---- phase 1: the chunk's home cells (sorted order => contiguous local ids) -----------------------------------------
        const int hmin = __reduce_min_sync(0xffffffffu, hloc), hmax = __reduce_max_sync(0xffffffffu, hloc);
        nn_scan_range(__ldg(cell_start + c0 + hmin), __ldg(cell_end + c0 + hmax), sorted, v);
        const unsigned long long home_done = (hmax >= 63 ? ~0ull : ((1ull << (hmax + 1)) - 1ull)) & ~((1ull << hmin) - 1ull);
        nn_coop_finish(q.x, q.y, q.z, active, v, g, G, shrink, cell_base, cell_start, cell_end, sorted, mask, false, 0u, bx0, by0, bz0, home_done);
        if (active) result[(size_t)b * Q + __float_as_int(q.w)] = v.bi;
    }
}

// ---- grouped queries (the chamfer workload): the queries arrive in groups of `q_mult` consecutive points sampled on ONE
// boundary face (dtb_surface_sample), i.e. each group is already a spatially compact patch a fraction of a grid cell wide.
// One warp takes (a chunk of up to 32 queries of) one group in its ORIGINAL order -- no query binning pass at all -- and
// proceeds like the brick kernel: phase 1 scans the distinct home cells of its lanes, phase 2 the remaining cells any lane's
// ball reaches.  Because the patch is compact the union of candidates over the warp is close to what a single lane needs.
__global__ void __launch_bounds__(128) nn_query_group_kernel(const float* __restrict__ queries, int Q, const int32_t* __restrict__ q_counts, int q_mult,
                                                             int G, const unsigned* __restrict__ bbox_ord, const unsigned* __restrict__ cell_start,
                                                             const unsigned* __restrict__ cell_end, const float4* __restrict__ sorted,
                                                             const unsigned long long* __restrict__ mask, int* __restrict__ result) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
#if NN_DENSE
    // 32 CONSECUTIVE queries per warp whatever the group size (a face's 20 samples leave 12 lanes idle otherwise): consecutive
    // boundary faces come from neighbouring tets, so the warp still covers one compact patch
    const long long nq = (long long)q_counts[b] * q_mult;
    const long long i0 = (long long)w * 32;
    if (i0 >= nq) return;                                     // warp-uniform
    const bool active = i0 + lane < nq;
    const size_t qi = (size_t)b * Q + (size_t)(active ? i0 + lane : i0);
#else
    const int cpg = (q_mult + 31) >> 5;                       // chunks per group
    const int group = w / cpg, chunk = w - group * cpg;
    if (group >= q_counts[b]) return;                         // warp-uniform
    const int l0 = chunk * 32;
    const bool active = l0 + lane < q_mult;
    const size_t qi = (size_t)b * Q + (size_t)group * q_mult + (active ? l0 + lane : l0);      // idle lanes copy the chunk's first query
#endif
    const float qx = __ldg(queries + qi * 3), qy = __ldg(queries + qi * 3 + 1), qz = __ldg(queries + qi * 3 + 2);
    const size_t cell_base = (size_t)b * G * G * G;
    const GridParams g = grid_params(bbox_ord, b, G);
    const float shrink = 1e-3f * g.h + 1e-6f * (fabsf(g.ox) + fabsf(g.oy) + fabsf(g.oz) + (float)G * g.h);   // as brick_walk (inflate = 0)
    NnVisitor v{qx, qy, qz, 1e20f, 0};
    const unsigned hc = cell_index(cell_coord(qx, g.ox, g.inv_h, G), cell_coord(qy, g.oy, g.inv_h, G), cell_coord(qz, g.oz, g.inv_h, G), G, true);
    // ---- phase 1: the distinct home cells of the lanes, every lane evaluates every candidate ------------------------------------
    unsigned todo = 0xffffffffu;
    while (todo) {
        const unsigned c = __shfl_sync(0xffffffffu, hc, __ffs(todo) - 1);
        nn_scan_range(__ldg(cell_start + cell_base + c), __ldg(cell_end + cell_base + c), sorted, v);
        todo &= ~__ballot_sync(0xffffffffu, hc == c);
    }
    nn_coop_finish(qx, qy, qz, active, v, g, G, shrink, cell_base, cell_start, cell_end, sorted, mask, true, hc, 0, 0, 0, 0ull);
    if (active) result[qi] = v.bi;
}

// ---- interpolation of a per-vertex field at the query points through the barycentric weights ----------
// out[b,p,:] = sum_k w[b,p,k] * field[b, tet[cond[b,p]][k], :]   (zeros where cond < 0)
__global__ void __launch_bounds__(256) interp_fwd_kernel(const float* __restrict__ field, const int32_t* __restrict__ tet, int V, int C,
                                                         const float* __restrict__ cond, const float* __restrict__ bary, int P,
                                                         float* __restrict__ out) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    size_t o = (size_t)b * P + i;
    float c = cond[o];
    float* y = out + o * C;
    if (!(c >= 0.f)) { for (int k = 0; k < C; ++k) y[k] = 0.f; return; }
    int4 id = reinterpret_cast<const int4*>(tet)[(int)c];
    float4 w = reinterpret_cast<const float4*>(bary)[o];
    const float* fb = field + (size_t)b * V * C;
    for (int k = 0; k < C; ++k)
        y[k] = w.x * fb[(size_t)id.x * C + k] + w.y * fb[(size_t)id.y * C + k] + w.z * fb[(size_t)id.z * C + k] + w.w * fb[(size_t)id.w * C + k];
}
__global__ void __launch_bounds__(256) interp_bwd_kernel(const float* __restrict__ field, const int32_t* __restrict__ tet, int V, int C,
                                                         const float* __restrict__ cond, const float* __restrict__ bary, int P,
                                                         const float* __restrict__ g_out, float* __restrict__ g_field,
                                                         float* __restrict__ g_bary) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    size_t o = (size_t)b * P + i;
    float c = cond[o];
    float4 gw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c >= 0.f) {
        int4 id = reinterpret_cast<const int4*>(tet)[(int)c];
        float4 w = reinterpret_cast<const float4*>(bary)[o];
        const float* fb = field + (size_t)b * V * C;
        const float* g = g_out + o * C;
        for (int k = 0; k < C; ++k) {
            float gk = g[k];
            gw.x += gk * fb[(size_t)id.x * C + k]; gw.y += gk * fb[(size_t)id.y * C + k];
            gw.z += gk * fb[(size_t)id.z * C + k]; gw.w += gk * fb[(size_t)id.w * C + k];
            if (g_field) {
                float* gf = g_field + (size_t)b * V * C;
                atomicAdd(gf + (size_t)id.x * C + k, gk * w.x); atomicAdd(gf + (size_t)id.y * C + k, gk * w.y);
                atomicAdd(gf + (size_t)id.z * C + k, gk * w.z); atomicAdd(gf + (size_t)id.w * C + k, gk * w.w);
            }
        }
    }
    if (g_bary) reinterpret_cast<float4*>(g_bary)[o] = gw;
}


// ---- masked mean-squared error over the located points (the occupancy-regression loss on interpolated values) ----
// loss[b] = sum_i m_i (x_i - t_i)^2 / max(sum_i m_i, 1), m_i = [cond_i >= 0]
__global__ void __launch_bounds__(256) mmse_fwd_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ cond, int P,
                                                       double* __restrict__ acc) {
    int b = blockIdx.y;
    double s = 0.0, c = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        size_t o = (size_t)b * P + i;
        if (cond[o] >= 0.f) { float d = x[o] - t[o]; s += (double)(d * d); c += 1.0; }
    }
    s = warp_sum(s); c = warp_sum(c);
    __shared__ double ss[8], sc[8];
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sc[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, n = 0;
        for (int w = 0; w < 8; ++w) { a += ss[w]; n += sc[w]; }
        atomicAdd(acc + b * 2, a); atomicAdd(acc + b * 2 + 1, n);
    }
}
__global__ void mmse_finalize_kernel(const double* __restrict__ acc, int B, float* __restrict__ loss) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) loss[b] = (float)(acc[b * 2] / fmax(acc[b * 2 + 1], 1.0));
}
__global__ void __launch_bounds__(256) mmse_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t, const float* __restrict__ cond, int P,
                                                       const double* __restrict__ acc, const float* __restrict__ g_loss, float* __restrict__ g_x) {
    int b = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    size_t o = (size_t)b * P + i;
    float s = 2.f * g_loss[b] / (float)fmax(acc[b * 2 + 1], 1.0);
    g_x[o] = cond[o] >= 0.f ? s * (x[o] - t[o]) : 0.f;
}
}  // namespace dtb

using namespace dtb;

static int default_grid_res(long long n_items, int lo, int hi) {
    int g = (int)ceil(cbrt((double)(n_items > 1 ? n_items : 1)));
    if (g < lo) g = lo;
    if (g > hi) g = hi;
    return g;
}

// ---- A1 -------------------------------------------------------------------------------------------
extern "C" int dtb_point_in_tet_grid_res(int T, int P) {
    (void)P;
    // cell edge ~ two tet bounding boxes: fewest (row visits + candidate tests) in the res-70 sweep (tools/sweep.py)
    int g = default_grid_res(T, 4, 320);
    g = (g + 1) / 2;
    return g < 4 ? 4 : g;
}
extern "C" size_t dtb_point_in_tet_workspace(int B, int P, int T, int G) {
    if (G <= 0) G = dtb_point_in_tet_grid_res(T, P);
    return pointgrid_workspace_bytes(B, P, G, false, false, PIT_XMULT) + align_up((size_t)B * P * sizeof(int), 256) +
           align_up((size_t)B * T * sizeof(int), 256) + align_up((size_t)B * sizeof(int), 256);
}

template <typename Src>
static int point_in_tet_impl(Src src, const float* points, int B, int T, int P, int G, float* cond, float* bary, void* workspace,
                             size_t workspace_bytes, cudaStream_t st) {
    DTB_REQUIRE(B > 0 && P >= 0 && T >= 0 && T < 0x7f7f7f7f, "point_in_tet: bad sizes B=%d P=%d T=%d", B, P, T);
    if (P == 0) return DTB_OK;
    if (G <= 0) G = dtb_point_in_tet_grid_res(T, P);
    Workspace ws(workspace, workspace_bytes);
    PointGrid pg;
    pointgrid_carve(pg, B, P, G, false, false, ws, PIT_XMULT);
    int* hit = ws.take<int>((size_t)B * P);
    int* weak_list = ws.take<int>((size_t)B * T);
    int* n_weak = ws.take<int>(B);
    if (!ws.ok || !workspace) { set_error("point_in_tet: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    int rc = pointgrid_build(pg, points, false, st);
    if (rc) return rc;
    DTB_CUDA(cudaMemsetAsync(hit, 0x7f, (size_t)B * P * sizeof(int), st));    // 0x7f7f7f7f: larger than any tet id
    DTB_CUDA(cudaMemsetAsync(n_weak, 0, (size_t)B * sizeof(int), st));
    if (T > 0) {
        dim3 grid(cdiv(T, 128), B);
        prof_begin(PROF_PIT_TET, st);
        pit_tet_kernel<Src><<<grid, 128, 0, st>>>(src, T, P, G, pg.bbox_ord, pg.cell_start, pg.cell_end, pg.sorted, hit, weak_list, n_weak);
        DTB_LAUNCH_CHECK("pit_tet");
        prof_end(PROF_PIT_TET, st);
        dim3 gw(296, B);
        pit_weak_kernel<Src><<<gw, 256, 0, st>>>(src, T, points, P, weak_list, n_weak, hit);
        DTB_LAUNCH_CHECK("pit_weak");
    }
    dim3 gf(cdiv(P, 256), B);
    pit_finalize_kernel<Src><<<gf, 256, 0, st>>>(src, T, points, P, hit, cond, bary);
    DTB_LAUNCH_CHECK("pit_finalize");
    return DTB_OK;
}

extern "C" int dtb_point_in_tet(const float* pos, const int32_t* tet, const float* points, int B, int V, int T, int P, int G,
                                float* cond, float* bary, void* workspace, size_t workspace_bytes, void* stream) {
    if (P == 0) return DTB_OK;
    DTB_REQUIRE(pos && tet && points, "point_in_tet: null argument");
    DTB_REQUIRE_ALIGNED16(tet, "point_in_tet: tet");
    DTB_REQUIRE_ALIGNED16(bary, "point_in_tet: bary");
    PitIndexed src{pos, tet, V, T};
    return point_in_tet_impl(src, points, B, T, P, G, cond, bary, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int dtb_point_in_tet_soup(const float* tet_bxfx4x3, const float* points, int B, int T, int P, int G, float* cond,
                                     float* bary, void* workspace, size_t workspace_bytes, void* stream) {
    if (P == 0) return DTB_OK;
    DTB_REQUIRE(tet_bxfx4x3 && points, "point_in_tet_soup: null argument");
    DTB_REQUIRE_ALIGNED16(tet_bxfx4x3, "point_in_tet_soup: tet_bxfx4x3");
    DTB_REQUIRE_ALIGNED16(bary, "point_in_tet_soup: bary");
    PitSoup src{tet_bxfx4x3, T};
    return point_in_tet_impl(src, points, B, T, P, G, cond, bary, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int dtb_tet_barycentric_backward(const float* pos, const int32_t* tet, const float* points, const float* cond,
                                            const float* g_w, int B, int V, int T, int P, float* grad_pos, int grad_stride,
                                            float* grad_points, void* stream) {
    (void)T;
    DTB_REQUIRE(pos && tet && points && cond && g_w, "tet_barycentric_backward: null argument");
    DTB_REQUIRE_ALIGNED16(tet, "tet_barycentric_backward: tet");
    DTB_REQUIRE_ALIGNED16(g_w, "tet_barycentric_backward: g_w");
    DTB_REQUIRE(grad_stride == 3 || (grad_stride == 4 && (((size_t)grad_pos) & 15) == 0), "tet_barycentric_backward: grad_stride must be 3, or 4 with a 16-byte aligned buffer");
    if (P == 0 || B == 0) return DTB_OK;
    dim3 grid(cdiv(P, 256), B);
    prof_begin(PROF_BARY_BWD, (cudaStream_t)stream);
    bary_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pos, tet, V, points, P, cond, g_w, grad_pos, grad_stride, grad_points);
    DTB_LAUNCH_CHECK("bary_backward");
    prof_end(PROF_BARY_BWD, (cudaStream_t)stream);
    return DTB_OK;
}

// ---- A2 -------------------------------------------------------------------------------------------
extern "C" int dtb_nearest_neighbor_grid_res(int M) {
    // target points usually sample a surface: ~M^(1/2) cells per axis keeps a few points per occupied cell
    // (res-70 sweep, 100 k surface targets: 32 -> 0.197, 40 -> 0.187, 48 -> 0.186, 56 -> 0.193, 64 -> 0.205 ms for the grouped kernel)
    int g = (int)ceil(sqrt((double)(M > 1 ? M : 1)) * 0.15);
    g = (g + 3) / 4 * 4;                 // brick layout: multiple of 4
    if (g < 4) g = 4;
    if (g > 128) g = 128;
    return g;
}
extern "C" size_t dtb_nearest_neighbor_workspace(int B, int Q, int M, int G) {
    if (G <= 0) G = dtb_nearest_neighbor_grid_res(M);
    G = (G + 3) / 4 * 4;
    size_t cells = (size_t)B * G * G * G;
    return pointgrid_workspace_bytes(B, M, G, true, true) + 2 * align_up(cells * 4, 256) + align_up((size_t)B * Q * 4, 256) +
           align_up((size_t)B * Q * 16, 256) + scan_workspace_bytes(cells) + 512;
}
// DTB_NN_KERNEL=thread selects the round-1 per-thread walk (A/B measurements, tests of both kernels); default = brick kernel
static bool nn_use_thread_walk() {
    const char* e = getenv("DTB_NN_KERNEL");          // read per call: tests and A/B runs toggle it inside one process
    return e && e[0] == 't';
}
static bool nn_use_brick_for_groups() {
    const char* e = getenv("DTB_NN_KERNEL");          // "brick": sort the grouped queries by cell like ungrouped ones (A/B)
    return e && e[0] == 'b';
}
static int nearest_neighbor_impl(const float* queries, const float* points, int32_t* result, int B, int Q, int M, int G,
                                 const int32_t* q_counts, int q_mult, void* workspace, size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(B > 0 && Q >= 0 && M >= 0, "nearest_neighbor: bad sizes");
    if (Q == 0) return DTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DTB_REQUIRE(queries && points && result, "nearest_neighbor: null argument");
    if (M == 0) {                       // reference leaves the zero-initialised result untouched
        DTB_CUDA(cudaMemsetAsync(result, 0, (size_t)B * Q * sizeof(int), st));
        return DTB_OK;
    }
    if (G <= 0) G = dtb_nearest_neighbor_grid_res(M);
    G = (G + 3) / 4 * 4;
    Workspace ws(workspace, workspace_bytes);
    PointGrid pg;
    pointgrid_carve(pg, B, M, G, true, true, ws);
    const size_t cells = (size_t)B * G * G * G;
    unsigned* qstart = ws.take<unsigned>(cells);
    unsigned* qend = ws.take<unsigned>(cells);
    unsigned* qcell = ws.take<unsigned>((size_t)B * Q);
    float4* qsorted = ws.take<float4>((size_t)B * Q);
    size_t qsb = scan_workspace_bytes(cells);
    void* qsws = ws.take<char>(qsb);
    if (!ws.ok || !workspace) { set_error("nearest_neighbor: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    int rc = pointgrid_build(pg, points, false, st);
    if (rc) return rc;
    if (q_counts && q_mult > 0 && !nn_use_thread_walk() && !nn_use_brick_for_groups()) {
        // grouped queries: no query binning, one warp per (chunk of a) group in the original order
#if NN_DENSE
        const long long warps = ((long long)Q + 31) / 32;
#else
        const int cpg = (q_mult + 31) / 32;
        const long long warps = (long long)(Q / q_mult) * cpg;
#endif
        dim3 grid(cdiv(warps, 4), B);
        prof_begin(PROF_NN_QUERY, st);
        nn_query_group_kernel<<<grid, 128, 0, st>>>(queries, Q, q_counts, q_mult, G, pg.bbox_ord, pg.cell_start, pg.cell_end, pg.sorted, pg.mask, result);
        DTB_LAUNCH_CHECK("nn_query_group");
        prof_end(PROF_NN_QUERY, st);
        return DTB_OK;
    }
    dim3 gq(cdiv(Q, 256), B);
    DTB_CUDA(cudaMemsetAsync(qstart, 0, cells * sizeof(unsigned), st));
    nn_qbin_count_kernel<<<gq, 256, 0, st>>>(queries, Q, G, pg.bbox_ord, q_counts, q_mult, qstart, qcell);
    DTB_LAUNCH_CHECK("nn_qbin_count");
    rc = exclusive_scan_u32_dup(qstart, qstart, qend, cells, nullptr, qsws, qsb, st);
    if (rc) return rc;
    nn_qbin_fill_kernel<<<gq, 256, 0, st>>>(queries, Q, qcell, qend, qsorted);
    DTB_LAUNCH_CHECK("nn_qbin_fill");
    prof_begin(PROF_NN_QUERY, st);
    if (nn_use_thread_walk()) {
        dim3 grid(cdiv(Q, 128), B);
        nn_query_thread_kernel<<<grid, 128, 0, st>>>(queries, Q, G, pg.bbox_ord, pg.cell_start, pg.cell_end, pg.sorted, pg.mask, qsorted, qstart, qend,
                                                     q_counts, q_mult, result);
        DTB_LAUNCH_CHECK("nn_query_thread_sorted");
    } else {
        dim3 grid((G / 4) * (G / 4) * (G / 4), B);
        nn_query_brick_kernel<<<grid, 128, 0, st>>>(Q, G, pg.bbox_ord, pg.cell_start, pg.cell_end, pg.sorted, pg.mask, qsorted, qstart, qend, result);
        DTB_LAUNCH_CHECK("nn_query_brick");
    }
    prof_end(PROF_NN_QUERY, st);
    return DTB_OK;
}

extern "C" int dtb_nearest_neighbor(const float* queries, const float* points, int32_t* result, int B, int Q, int M, int G,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    return nearest_neighbor_impl(queries, points, result, B, Q, M, G, nullptr, 1, workspace, workspace_bytes, stream);
}

// padded-ragged queries: only the first q_counts[b] * q_mult queries of sample b are answered
extern "C" int dtb_nearest_neighbor_ragged(const float* queries, const int32_t* q_counts, int q_mult, const float* points,
                                           int32_t* result, int B, int Qmax, int M, int G, void* workspace, size_t workspace_bytes,
                                           void* stream) {
    DTB_REQUIRE(q_counts != nullptr, "nearest_neighbor_ragged: null q_counts");
    return nearest_neighbor_impl(queries, points, result, B, Qmax, M, G, q_counts, q_mult, workspace, workspace_bytes, stream);
}

// field (B,V,C), out (B,P,C): interpolate a per-vertex field at the query points (see kernel comment)
extern "C" int dtb_tet_interpolate_forward(const float* field, const int32_t* tet, const float* cond, const float* bary, int B, int V, int C,
                                           int P, float* out, void* stream) {
    if (P == 0 || B == 0) return DTB_OK;
    DTB_REQUIRE(field && tet && cond && bary && out && C > 0, "tet_interpolate_forward: bad argument");
    DTB_REQUIRE_ALIGNED16(tet, "tet_interpolate_forward: tet");
    DTB_REQUIRE_ALIGNED16(bary, "tet_interpolate_forward: bary");
    dim3 grid(cdiv(P, 256), B);
    interp_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(field, tet, V, C, cond, bary, P, out);
    DTB_LAUNCH_CHECK("interp_fwd");
    return DTB_OK;
}
// g_field (B,V,C) is ACCUMULATED (may be NULL); g_bary (B,P,4) is overwritten (may be NULL)
extern "C" int dtb_tet_interpolate_backward(const float* field, const int32_t* tet, const float* cond, const float* bary, const float* g_out,
                                            int B, int V, int C, int P, float* g_field, float* g_bary, void* stream) {
    if (P == 0 || B == 0) return DTB_OK;
    DTB_REQUIRE(field && tet && cond && bary && g_out && C > 0, "tet_interpolate_backward: bad argument");
    DTB_REQUIRE_ALIGNED16(tet, "tet_interpolate_backward: tet");
    DTB_REQUIRE_ALIGNED16(bary, "tet_interpolate_backward: bary");
    DTB_REQUIRE_ALIGNED16(g_bary, "tet_interpolate_backward: g_bary");
    dim3 grid(cdiv(P, 256), B);
    interp_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(field, tet, V, C, cond, bary, P, g_out, g_field, g_bary);
    DTB_LAUNCH_CHECK("interp_bwd");
    return DTB_OK;
}

// x, t, cond (B,P); acc (B,2) f64 scratch kept for backward; loss (B,)
extern "C" int dtb_masked_mse_forward(const float* x, const float* t, const float* cond, int B, int P, double* acc, float* loss, void* stream) {
    DTB_REQUIRE(acc && loss, "masked_mse_forward: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    DTB_CUDA(cudaMemsetAsync(acc, 0, (size_t)B * 2 * sizeof(double), st));
    if (P > 0) {
        DTB_REQUIRE(x && t && cond, "masked_mse_forward: null argument");
        dim3 grid(min(cdiv(P, 256 * 4), 64), B);
        mmse_fwd_kernel<<<grid, 256, 0, st>>>(x, t, cond, P, acc);
        DTB_LAUNCH_CHECK("mmse_fwd");
    }
    mmse_finalize_kernel<<<cdiv(B, 64), 64, 0, st>>>(acc, B, loss);
    DTB_LAUNCH_CHECK("mmse_finalize");
    return DTB_OK;
}
extern "C" int dtb_masked_mse_backward(const float* x, const float* t, const float* cond, const double* acc, const float* g_loss, int B, int P,
                                       float* g_x, void* stream) {
    if (P == 0 || B == 0) return DTB_OK;
    DTB_REQUIRE(x && t && cond && acc && g_loss && g_x, "masked_mse_backward: null argument");
    dim3 grid(cdiv(P, 256), B);
    mmse_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, t, cond, P, acc, g_loss, g_x);
    DTB_LAUNCH_CHECK("mmse_bwd");
    return DTB_OK;
}
