// A4: point -> triangle-set squared distance, forward and backward.
//   forward   reference layers/DefTet/tet_analytic_distance_batch/tet_analytic_distance_for.cu:139-307
//   backward  reference layers/DefTet/tet_analytic_distance_batch/tet_analytic_distance_back.cu:222-317,348-483,592-686
// The reference tests every point against every face (O(S*F)).  Here the faces are binned by centroid into
// a brick grid and each point walks outwards with the bound  dist(point, cell box) - R_max, R_max = largest
// centroid-to-vertex distance of the sample, evaluating the reference's exact (non-contracted fp32)
// triangle-distance expression only on the surviving faces and keeping the lexicographic minimum
// (distance, face id) = "first strict minimum" of the reference.  The reference distance is the true
// point-triangle distance except that its inside test is done on the xy-projection: faces whose normal is
// (almost) horizontal can be mis-classified, so they are not pruned geometrically but kept on a short
// per-sample "always test" list; faces with k3 == 0 exactly are invisible to the reference (:176-183) and
// stay invisible here.  All reference quirks (edge-case gradient overwrite at _back.cu:309-315, MAX_DIS
// sentinels) are reproduced.
#include "brickwalk.cuh"
#include "deftet_b200.h"

namespace dtb {

constexpr float FWD_MAX_DIS = 10000.0f;     // tet_analytic_distance_for.cu:17
constexpr float BWD_MAX_DIS = 9999999.0f;   // tet_analytic_distance_back.cu:19

// cuda_divide_non_zero: `a + eps` with a double eps is evaluated in double and narrowed on return
__device__ __forceinline__ float div_nz(float a) {
    if (a == 0.f) return (float)1e-10;
    if (a < 0.f) return (float)((double)a - 1e-10);
    if (a > 0.f) return (float)((double)a + 1e-10);
    return (float)1e-10;
}
__device__ __forceinline__ float xdot(const float* a, const float* b) { return xadd(xadd(xmul(a[0], b[0]), xmul(a[1], b[1])), xmul(a[2], b[2])); }
__device__ __forceinline__ float xmin3(float a, float b, float c) { float m = a; if (b < m) m = b; if (c < m) m = c; return m; }
__device__ __forceinline__ int xmin3_idx(float a, float b, float c) { float m = a; int i = 0; if (b < m) { m = b; i = 1; } if (c < m) { m = c; i = 2; } return i; }
__device__ __forceinline__ float xabs(float a) { return a > 0.0f ? a : -a; }
__device__ __forceinline__ float pt_dist2(const float* a, const float* p) {
    float e0 = xsub(a[0], p[0]), e1 = xsub(a[1], p[1]), e2 = xsub(a[2], p[2]);
    return xadd(xadd(xmul(e0, e0), xmul(e1, e1)), xmul(e2, e2));
}
// distance_line_square: negative when the foot point is outside the segment
__device__ __forceinline__ float line_dist2(const float* A, const float* B, const float* P, float* t_out) {
    float PA[3] = {xsub(P[0], A[0]), xsub(P[1], A[1]), xsub(P[2], A[2])};
    float BA[3] = {xsub(B[0], A[0]), xsub(B[1], A[1]), xsub(B[2], A[2])};
    float t = xdiv(xdot(PA, BA), div_nz(xdot(BA, BA)));
    float d[3] = {xsub(PA[0], xmul(BA[0], t)), xsub(PA[1], xmul(BA[1], t)), xsub(PA[2], xmul(BA[2], t))};
    float dist = xdot(d, d);
    if (t_out) *t_out = t;
    return (t >= 0.f && t <= 1.f) ? dist : -dist;
}

struct TriHit { int type; float plane2; float inplane; int idx; float ip[3]; float l1, l2, l3; };

// cuda_min_triangle_distance + cuda_line_distance.  Returns the reference distance; fills hit.
__device__ __forceinline__ float tri_distance(const float* a, const float* b, const float* c, const float* p, float max_dis, TriHit& h) {
    float r1[3] = {xsub(b[0], a[0]), xsub(b[1], a[1]), xsub(b[2], a[2])};
    float r2[3] = {xsub(c[0], a[0]), xsub(c[1], a[1]), xsub(c[2], a[2])};
    float n[3] = {xsub(xmul(r1[1], r2[2]), xmul(r1[2], r2[1])), xsub(xmul(r1[2], r2[0]), xmul(r1[0], r2[2])),
                  xsub(xmul(r1[0], r2[1]), xmul(r1[1], r2[0]))};
    float len = div_nz(xsqrt(xadd(xadd(xmul(n[0], n[0]), xmul(n[1], n[1])), xmul(n[2], n[2]))));
    n[0] = xdiv(n[0], len); n[1] = xdiv(n[1], len); n[2] = xdiv(n[2], len);
    float t = xsub(xdot(n, a), xdot(n, p));
    float ip[3] = {xadd(p[0], xmul(n[0], t)), xadd(p[1], xmul(n[1], t)), xadd(p[2], xmul(n[2], t))};
    h.ip[0] = ip[0]; h.ip[1] = ip[1]; h.ip[2] = ip[2];
    h.plane2 = xmul(t, t);
    h.idx = 0; h.inplane = 0.f;
    // xy-projected barycentric test
    float k1 = xadd(xmul(xsub(b[1], c[1]), xsub(ip[0], c[0])), xmul(xsub(c[0], b[0]), xsub(ip[1], c[1])));
    float k2 = xadd(xmul(xsub(a[0], c[0]), xsub(ip[1], c[1])), xmul(xsub(c[1], a[1]), xsub(ip[0], c[0])));
    float k3 = xadd(xmul(xsub(b[1], c[1]), xsub(a[0], c[0])), xmul(xsub(c[0], b[0]), xsub(a[1], c[1])));
    if (k3 == 0.f) { h.type = -1; return max_dis; }
    float l1 = xdiv(k1, k3), l2 = xdiv(k2, k3), l3 = xsub(xsub(1.0f, l1), l2);
    h.l1 = l1; h.l2 = l2; h.l3 = l3;
    float d12 = line_dist2(a, b, ip, nullptr), d23 = line_dist2(b, c, ip, nullptr), d13 = line_dist2(a, c, ip, nullptr);
    if (l1 >= 0.f && l2 >= 0.f && l3 >= 0.f) {
        h.type = 0;
        h.inplane = xmin3(xabs(d12), xabs(d23), xabs(d13));
        h.idx = xmin3_idx(xabs(d12), xabs(d23), xabs(d13));
        return h.plane2;
    }
    if (d12 <= 0.f) d12 = max_dis;
    if (d23 <= 0.f) d23 = max_dis;
    if (d13 <= 0.f) d13 = max_dis;
    float ml = xmin3(d12, d23, d13);
    int mli = xmin3_idx(d12, d23, d13);
    float e1 = pt_dist2(a, ip), e2 = pt_dist2(b, ip), e3 = pt_dist2(c, ip);
    float mp = xmin3(e1, e2, e3);
    int mpi = xmin3_idx(e1, e2, e3);
    if (ml < mp) { h.type = 1; h.inplane = ml; h.idx = mli; }
    else { h.type = 2; h.inplane = mp; h.idx = mpi; }
    return xadd(h.plane2, h.inplane);
}

// Query-independent part of cuda_min_triangle_distance / cuda_line_distance, hoisted out of the per-query loop.
// Every value is produced by the same operation sequence as in tri_distance(), so results stay bit-identical.
struct FacePre {
    float a[3], b[3], c[3];
    float n[3];          // unit normal
    float na;            // dot(n, a)
    float k3;
    float bc1, cb0, ac0, ca1;        // (b1-c1), (c0-b0), (a0-c0), (c1-a1)
    float ab[3], bcv[3], ac[3];      // edge vectors B-A for the three edges (a,b), (b,c), (a,c)
    float den_ab, den_bc, den_ac;    // div_nz(dot(BA,BA))
    float pad[2];                    // 32 floats = 8 x float4
};
constexpr int FACEPRE_FLOATS = sizeof(FacePre) / 4;
// shared-memory stride of a staged FacePre: odd, so that lanes reading the same field of different candidates hit
// different banks (with a stride of 32 floats every such access would be a 32-way bank conflict)
constexpr int FACEPRE_STRIDE = FACEPRE_FLOATS + 1;

__device__ __forceinline__ void face_precompute(const float* t, FacePre& f) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { f.a[k] = t[k]; f.b[k] = t[3 + k]; f.c[k] = t[6 + k]; }
    float r1[3] = {xsub(f.b[0], f.a[0]), xsub(f.b[1], f.a[1]), xsub(f.b[2], f.a[2])};
    float r2[3] = {xsub(f.c[0], f.a[0]), xsub(f.c[1], f.a[1]), xsub(f.c[2], f.a[2])};
    float n[3] = {xsub(xmul(r1[1], r2[2]), xmul(r1[2], r2[1])), xsub(xmul(r1[2], r2[0]), xmul(r1[0], r2[2])),
                  xsub(xmul(r1[0], r2[1]), xmul(r1[1], r2[0]))};
    float raw_len = xsqrt(xadd(xadd(xmul(n[0], n[0]), xmul(n[1], n[1])), xmul(n[2], n[2])));
    float len = div_nz(raw_len);
    f.pad[0] = raw_len;                  // |cross| = twice the area: used to flag faces too small for geometric pruning
    f.pad[1] = 0.f;
    f.n[0] = xdiv(n[0], len); f.n[1] = xdiv(n[1], len); f.n[2] = xdiv(n[2], len);
    f.na = xdot(f.n, f.a);
    f.bc1 = xsub(f.b[1], f.c[1]); f.cb0 = xsub(f.c[0], f.b[0]); f.ac0 = xsub(f.a[0], f.c[0]); f.ca1 = xsub(f.c[1], f.a[1]);
    f.k3 = xadd(xmul(f.bc1, f.ac0), xmul(f.cb0, xsub(f.a[1], f.c[1])));
#pragma unroll
    for (int k = 0; k < 3; ++k) { f.ab[k] = xsub(f.b[k], f.a[k]); f.bcv[k] = xsub(f.c[k], f.b[k]); f.ac[k] = xsub(f.c[k], f.a[k]); }
    f.den_ab = div_nz(xdot(f.ab, f.ab)); f.den_bc = div_nz(xdot(f.bcv, f.bcv)); f.den_ac = div_nz(xdot(f.ac, f.ac));
}
__device__ __forceinline__ float line_dist2_pre(const float* A, const float* BA, float den, const float* P) {
    float PA[3] = {xsub(P[0], A[0]), xsub(P[1], A[1]), xsub(P[2], A[2])};
    float t = xdiv(xdot(PA, BA), den);
    float d[3] = {xsub(PA[0], xmul(BA[0], t)), xsub(PA[1], xmul(BA[1], t)), xsub(PA[2], xmul(BA[2], t))};
    float dist = xdot(d, d);
    return (t >= 0.f && t <= 1.f) ? dist : -dist;
}
// forward-only distance (same value as tri_distance(..., FWD_MAX_DIS, ...))
__device__ __forceinline__ float tri_distance_pre(const FacePre& f, const float* p) {
    float t = xsub(f.na, xdot(f.n, p));
    float ip[3] = {xadd(p[0], xmul(f.n[0], t)), xadd(p[1], xmul(f.n[1], t)), xadd(p[2], xmul(f.n[2], t))};
    float plane2 = xmul(t, t);
    if (f.k3 == 0.f) return FWD_MAX_DIS;
    float ipc0 = xsub(ip[0], f.c[0]), ipc1 = xsub(ip[1], f.c[1]);
    float k1 = xadd(xmul(f.bc1, ipc0), xmul(f.cb0, ipc1));
    float k2 = xadd(xmul(f.ac0, ipc1), xmul(f.ca1, ipc0));
    float l1 = xdiv(k1, f.k3), l2 = xdiv(k2, f.k3), l3 = xsub(xsub(1.0f, l1), l2);
    if (l1 >= 0.f && l2 >= 0.f && l3 >= 0.f) return plane2;
    float d12 = line_dist2_pre(f.a, f.ab, f.den_ab, ip), d23 = line_dist2_pre(f.b, f.bcv, f.den_bc, ip), d13 = line_dist2_pre(f.a, f.ac, f.den_ac, ip);
    if (d12 <= 0.f) d12 = FWD_MAX_DIS;
    if (d23 <= 0.f) d23 = FWD_MAX_DIS;
    if (d13 <= 0.f) d13 = FWD_MAX_DIS;
    float ml = xmin3(d12, d23, d13);
    float mp = xmin3(pt_dist2(f.a, ip), pt_dist2(f.b, ip), pt_dist2(f.c, ip));
    return xadd(plane2, (ml < mp) ? ml : mp);
}

// ---- per-sample face statistics: R_max and the "always test" list ---------------------------------------
__global__ void __launch_bounds__(256) face_stats_kernel(const float* __restrict__ soup, const int32_t* __restrict__ counts, int Fmax,
                                                         unsigned* __restrict__ rmax_bits, int32_t* __restrict__ always,
                                                         int32_t* __restrict__ n_always, int always_cap, float4* __restrict__ pre) {
    int b = blockIdx.y;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    float r = 0.f;
    if (f < counts[b]) {
        const float* t = soup + ((size_t)b * Fmax + f) * 9;
        float cx = (t[0] + t[3] + t[6]) * (1.f / 3.f), cy = (t[1] + t[4] + t[7]) * (1.f / 3.f), cz = (t[2] + t[5] + t[8]) * (1.f / 3.f);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float dx = t[k * 3] - cx, dy = t[k * 3 + 1] - cy, dz = t[k * 3 + 2] - cz;
            r = fmaxf(r, sqrtf(dx * dx + dy * dy + dz * dz));
        }
        if (pre) {                                   // query-independent half of the distance, once per face
            __align__(16) float fp_buf[FACEPRE_FLOATS];          // 16-byte aligned: copied out below as eight float4
            FacePre& fp = *reinterpret_cast<FacePre*>(fp_buf);
            face_precompute(t, fp);
            // this face's own bounding radius around its centroid (inflated like rmax; +inf when undefined: never pruned by it)
            fp.pad[1] = (r == r) ? r * 1.001f + 1e-7f : __int_as_float(0x7f800000);
            const float4* src = reinterpret_cast<const float4*>(fp_buf);
            float4* dst = pre + ((size_t)b * Fmax + f) * 8;
#pragma unroll
            for (int m = 0; m < 8; ++m) dst[m] = src[m];
        }
        float e1[3] = {t[3] - t[0], t[4] - t[1], t[5] - t[2]}, e2[3] = {t[6] - t[0], t[7] - t[1], t[8] - t[2]};
        float nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
        float nn = sqrtf(nx * nx + ny * ny + nz * nz);
        // Never prune a face geometrically when the reference distance can deviate from the Euclidean one by more than the
        // pruning slack: xy-projection unreliable (normal almost horizontal), not a finite triangle, or so small that the
        // reference's absolute epsilons (cuda_divide_non_zero: len + 1e-10, |BA|^2 + 1e-10) are no longer negligible.
        float l1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2], l2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
        float e3x = t[6] - t[3], e3y = t[7] - t[4], e3z = t[8] - t[5];
        float l3 = e3x * e3x + e3y * e3y + e3z * e3z;
        bool tiny = !(nn > 1e-5f) || !(fminf(l1, fminf(l2, l3)) > 1e-5f);
        bool unreliable = !(fabsf(nz) > 1e-3f * nn) || !(r == r) || !(nn == nn) || tiny;
        if (unreliable) {
            int k = atomicAdd(n_always + b, 1);
            if (k < always_cap) always[(size_t)b * always_cap + k] = f;
        }
        if (!(r == r)) r = 0.f;
    }
    r = warp_max(r);
    if ((threadIdx.x & 31) == 0 && r > 0.f) atomicMax(rmax_bits + b, __float_as_uint(r));
}

struct TriVisitor {
    const float* soup;      // this sample's (Fmax,3,3)
    float p[3];
    float rmax;             // largest centroid-to-vertex distance of the sample (inflated)
    float best; int bi;
    __device__ __forceinline__ float bound() const { return best; }
    __device__ static __forceinline__ float no_hit() { return 10000.0f; }
    __device__ __forceinline__ void face(int f) {
        const float* t = soup + (size_t)f * 9;
        float a[3] = {__ldg(t), __ldg(t + 1), __ldg(t + 2)}, b[3] = {__ldg(t + 3), __ldg(t + 4), __ldg(t + 5)},
              c[3] = {__ldg(t + 6), __ldg(t + 7), __ldg(t + 8)};
        TriHit h;
        float d = tri_distance(a, b, c, p, FWD_MAX_DIS, h);
        if (best > d || (d == best && bi >= 0 && f < bi)) { best = d; bi = f; }
    }
    __device__ __forceinline__ void item(const float4& it) {
        // cheap conservative reject: the face lies inside the ball (centroid, rmax)
        float dx = it.x - p[0], dy = it.y - p[1], dz = it.z - p[2];
        float lb = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz) - rmax, 0.f) * 0.9999f;
        if (lb * lb > best) return;
        face(__float_as_int(it.w));
    }
};

// ---- query binning: the S points of each sample counting-sorted by the CELL of the FACE grid they fall in (brick-major cell order:
// the queries of a brick stay contiguous, and 32 consecutive queries are spatially compact) ----
__global__ void __launch_bounds__(256) qbin_count_kernel(const float* __restrict__ points, int S, int G, const unsigned* __restrict__ bbox_ord,
                                                         unsigned* __restrict__ qcount, unsigned* __restrict__ qbrick) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    GridParams g = grid_params(bbox_ord, b, G);
    const float* p = points + ((size_t)b * S + i) * 3;
    unsigned id = (unsigned)b * G * G * G +
                  cell_index(cell_coord(p[0], g.ox, g.inv_h, G), cell_coord(p[1], g.oy, g.inv_h, G), cell_coord(p[2], g.oz, g.inv_h, G), G, true);
    qbrick[(size_t)b * S + i] = id;
    atomicAdd(qcount + id, 1u);
}
__global__ void __launch_bounds__(256) qbin_fill_kernel(const float* __restrict__ points, int S, const unsigned* __restrict__ qbrick,
                                                        unsigned* __restrict__ qend, float4* __restrict__ qsorted) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const float* p = points + ((size_t)b * S + i) * 3;
    unsigned dst = atomicAdd(qend + qbrick[(size_t)b * S + i], 1u);
    qsorted[dst] = make_float4(p[0], p[1], p[2], __int_as_float(i));
}

// One CTA per brick of queries: the faces binned in the 3x3x3 surrounding bricks are staged once in shared
// memory (centroid + the query-independent half of the distance computation) and every query of the brick
//   1. finds the candidate with the nearest centroid and evaluates it (all lanes together: no divergence, and the
//      running minimum is tight from the start),
//   2. re-scans the candidates with a bounding-sphere reject and evaluates the few survivors.
// A query whose best distance cannot be certified against faces outside the neighbourhood falls back to the general
// brick walk.  Results are identical to the brute-force scan (lexicographic minimum of (distance, face id)).
#ifndef PFD_THREADS_N
#define PFD_THREADS_N 128
#endif
constexpr int PFD_THREADS = PFD_THREADS_N;
#ifndef PFD_MIN_CTAS
#define PFD_MIN_CTAS 6
#endif
#ifndef PFD_REACH_H
#define PFD_REACH_H 1.25f
#endif
#ifndef PFD_CHUNK_N
#define PFD_CHUNK_N 128
#endif
#ifndef PFD_SCAN_UNROLL
#define PFD_SCAN_UNROLL 4          // A/B on the surface-distance op (res 70 b8): 1 -> 0.339, 2 -> 0.341, 4 -> 0.328 ms
#endif
#define DTB_PRAGMA_(x) _Pragma(#x)
#define DTB_UNROLL(n) DTB_PRAGMA_(unroll n)
#ifndef PFD_PLANE_PRUNE
#define PFD_PLANE_PRUNE 1            // scan 2 also rejects faces whose PLANE is farther than the bound (A/B: -8 % on the op, bit-identical)
#endif
constexpr int PFD_CHUNK = PFD_CHUNK_N;      // staged candidate capacity (< 256: survivor lists hold uint8 indices)
constexpr int PFD_SCAN = PFD_CHUNK_N;       // candidates of the 27 bricks inspected per staging round (<= PFD_CHUNK)
constexpr int PFD_LIST = 12;        // survivors remembered per query and chunk

__device__ __forceinline__ unsigned long long pack_df(float d, int f) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)f; }

__global__ void __launch_bounds__(PFD_THREADS, PFD_MIN_CTAS) pfd_forward_tiled_kernel(
    int S, const float* __restrict__ soup, const float4* __restrict__ pre, const int32_t* __restrict__ counts, int Fmax, int G,
    const unsigned* __restrict__ bbox_ord, const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_end,
    const float4* __restrict__ sorted, const unsigned long long* __restrict__ mask, const unsigned* __restrict__ rmax_bits,
    const int32_t* __restrict__ always, const int32_t* __restrict__ n_always, int always_cap, const unsigned* __restrict__ qstart,
    const unsigned* __restrict__ qend, const float4* __restrict__ qsorted, float* __restrict__ closest_d, float* __restrict__ closest_f) {
    constexpr int WARPS = PFD_THREADS / 32;
    __shared__ float4 s_cen[PFD_CHUNK];
    __shared__ float s_pre[PFD_CHUNK * FACEPRE_STRIDE];
    __shared__ unsigned s_rs[27], s_re[27];
    __shared__ unsigned s_total;
    __shared__ unsigned s_n;
    __shared__ unsigned char s_rel[PFD_CHUNK];        // 1 = the centroid distance is a valid upper bound for this face
    __shared__ unsigned char s_sub[WARPS][PFD_CHUNK]; // per warp: the staged candidates near this warp's 32 (cell-sorted, compact) queries
    __shared__ unsigned char s_list[PFD_THREADS][PFD_LIST];
    __shared__ float s_q[WARPS][32][3];               // the warp's queries (pooled evaluation reads other lanes' queries)
    __shared__ unsigned long long s_best[WARPS][32];  // (distance bits, face id): atomicMin == lexicographic minimum
    __shared__ unsigned short s_pairs[WARPS][32 * PFD_LIST];
    const int b = blockIdx.y;
    const int NB = G >> 2;
    const int brick = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t qb = ((size_t)b * NB * NB * NB + brick) * 64;        // first cell of this brick (queries are sorted by cell)
    const unsigned q0 = qstart[qb], q1 = qend[qb + 63];
    if (q0 == q1) return;
    const int nf = counts[b];
    const int na = n_always[b];
    const int bz0 = brick / (NB * NB), by0 = (brick / NB) % NB, bx0 = brick % NB;
    const size_t cell_base = (size_t)b * G * G * G;
    const float* sb = soup + (size_t)b * Fmax * 9;
    const float4* preb = pre + (size_t)b * Fmax * 8;
    const GridParams g = grid_params(bbox_ord, b, G);
    const float rmax = __uint_as_float(rmax_bits[b]) * 1.001f + 1e-7f;
    const bool brute = na > always_cap;
    if (threadIdx.x < 27) {
        // slot 0 = home brick, then the 26 neighbours
        int o = threadIdx.x == 0 ? 13 : (threadIdx.x <= 13 ? threadIdx.x - 1 : threadIdx.x);
        int dz = o / 9 - 1, dy = (o / 3) % 3 - 1, dx = o % 3 - 1;
        int z = bz0 + dz, y = by0 + dy, x = bx0 + dx;
        unsigned rs = 0, re = 0;
        if (!brute && nf > 0 && z >= 0 && z < NB && y >= 0 && y < NB && x >= 0 && x < NB) {
            size_t br = ((size_t)z * NB + y) * NB + x;
            if (mask[(cell_base >> 6) + br]) { rs = cell_start[cell_base + br * 64]; re = cell_end[cell_base + br * 64 + 63]; }
        }
        s_rs[threadIdx.x] = rs; s_re[threadIdx.x] = re;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int k = 0; k < 27; ++k) t += s_re[k] - s_rs[k];
        s_total = t;
    }
    __syncthreads();
    const unsigned total = s_total;
    const float bw = 4.0f * g.h;
    const float slack = 1e-3f * g.h + 1e-6f * (fabsf(g.ox) + fabsf(g.oy) + fabsf(g.oz) + (float)G * g.h);
    // only faces whose centroid lies within `reach` of this brick are staged; a query is final when its search ball
    // (sqrt(best) + rmax) stays inside that region, otherwise it falls back to the general walk
    // The region can only be certified where it was staged, i.e. inside the 3x3x3 bricks: when the faces are large against the
    // grid (rmax + 1.25 h > one brick; found by the res-40 scale parity test, where a small shape gives a fine grid) it is clamped
    // to them, and the queries whose ball leaves the clamped region take the general walk.
    const float reach = fminf(rmax + PFD_REACH_H * g.h, bw);
    const float margin = fmaxf(reach - rmax, 0.f);          // = PFD_REACH_H cells unless the reach was clamped
    const float lx0 = g.ox + (float)bx0 * bw, ly0 = g.oy + (float)by0 * bw, lz0 = g.oz + (float)bz0 * bw;
    const float rlo[3] = {lx0 - reach, ly0 - reach, lz0 - reach}, rhi[3] = {lx0 + bw + reach, ly0 + bw + reach, lz0 + bw + reach};
    const float gmax = (float)G * g.h;
    for (unsigned qbase = q0; qbase < q1; qbase += PFD_THREADS) {
        const unsigned qi = qbase + threadIdx.x;
        const bool active = qi < q1;
        TriVisitor v;
        v.soup = sb; v.rmax = rmax; v.best = 10000.0f; v.bi = -1;
        int orig = 0;
        if (active) {
            float4 q = qsorted[qi];
            v.p[0] = q.x; v.p[1] = q.y; v.p[2] = q.z;
            orig = __float_as_int(q.w);
            if (brute) { for (int f = 0; f < nf; ++f) v.face(f); }
        } else { v.p[0] = v.p[1] = v.p[2] = 0.f; }
        s_q[warp][lane][0] = v.p[0]; s_q[warp][lane][1] = v.p[1]; s_q[warp][lane][2] = v.p[2];
        // The warp's 32 queries are neighbours (sorted by cell): per query only those staged faces are scanned whose bounding sphere
        // (centroid, the face's OWN radius) comes within `margin` of the queries' bounding box.  A staged face that is left out is
        // farther than `margin` from every query of the warp, so a query whose best distance is below `margin` needs none of them
        // (certified below, together with the faces that were not staged at all).
        float wlo[3], whi[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float lo = active ? v.p[k] : 3.0e38f, hi = active ? v.p[k] : -3.0e38f;
            wlo[k] = warp_min(lo); whi[k] = warp_max(hi);
        }
        float ub = 3.0e38f;                                 // upper bound of the answer: a centroid is a point of its face
        for (unsigned c0 = 0; c0 < total; c0 += PFD_SCAN) {
            // ---- stage (compacting): candidates c0 .. c0+PFD_SCAN of the 27 bricks that lie in the reach region ----
            __syncthreads();
            if (threadIdx.x == 0) s_n = 0;
            __syncthreads();
            for (unsigned k0 = 0; k0 < PFD_SCAN && c0 + k0 < total; k0 += PFD_THREADS) {
                unsigned off = c0 + k0 + threadIdx.x;
                bool keep = false;
                float4 it = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k0 + threadIdx.x < PFD_SCAN && off < total) {
                    int r = 0;
                    while (off >= s_re[r] - s_rs[r]) { off -= s_re[r] - s_rs[r]; ++r; }
                    it = sorted[s_rs[r] + off];
                    keep = it.x >= rlo[0] && it.x <= rhi[0] && it.y >= rlo[1] && it.y <= rhi[1] && it.z >= rlo[2] && it.z <= rhi[2];
                }
                unsigned bal = __ballot_sync(0xffffffffu, keep);
                unsigned base = 0;
                if (lane == 0 && bal) base = atomicAdd(&s_n, (unsigned)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (keep) {
                    unsigned k = base + __popc(bal & ((1u << lane) - 1u));
                    s_cen[k] = it;
                    const float4* src = preb + (size_t)__float_as_int(it.w) * 8;
                    float* dst = s_pre + (size_t)k * FACEPRE_STRIDE;
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        float4 w4 = __ldg(src + m);
                        dst[4 * m] = w4.x; dst[4 * m + 1] = w4.y; dst[4 * m + 2] = w4.z; dst[4 * m + 3] = w4.w;
                    }
                    const FacePre& fp = *reinterpret_cast<const FacePre*>(dst);
                    // same reliability rule as face_stats_kernel (unit normal here): unreliable or invisible faces may
                    // have a reference distance larger than the distance to their centroid
                    s_rel[k] = (fp.k3 != 0.f && fabsf(fp.n[2]) > 2e-3f && fminf(fp.den_ab, fminf(fp.den_bc, fp.den_ac)) > 1.0001e-5f &&
                                fp.pad[0] > 1e-5f) ? 1 : 0;
                }
            }
            __syncthreads();
            const int n_staged = (int)s_n;
            // ---- the warp's sub-list of the staged candidates ----
            int n = 0;
            for (int k0 = 0; k0 < n_staged; k0 += 32) {
                const int k = k0 + lane;
                bool in = false;
                if (k < n_staged) {
                    const float4 it = s_cen[k];
                    const float dx = fmaxf(fmaxf(wlo[0] - it.x, it.x - whi[0]), 0.f), dy = fmaxf(fmaxf(wlo[1] - it.y, it.y - whi[1]), 0.f);
                    const float dz = fmaxf(fmaxf(wlo[2] - it.z, it.z - whi[2]), 0.f);
                    const float rr = fminf(s_pre[(size_t)k * FACEPRE_STRIDE + (FACEPRE_FLOATS - 1)], rmax) + margin;      // own radius (inflated)
                    in = !(dx * dx + dy * dy + dz * dz > rr * rr * 1.0002f);          // conservative inclusion (NaN: included)
                }
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (in) s_sub[warp][n + __popc(bal & ((1u << lane) - 1u))] = (unsigned char)k;
                n += __popc(bal);
            }
            __syncwarp();
            const unsigned char* sub = s_sub[warp];
            // ---- scan 1: upper bound from the centroid distances, and the nearest centroid ----
            int kn = -1;
            float dn = 3.0e38f;
            if (active) {
DTB_UNROLL(PFD_SCAN_UNROLL)
                for (int i = 0; i < n; ++i) {
                    const int k = sub[i];
                    float4 it = s_cen[k];
                    float dx = it.x - v.p[0], dy = it.y - v.p[1], dz = it.z - v.p[2];
                    float d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 < dn) { dn = d2; kn = k; }
                    if (s_rel[k]) ub = fminf(ub, d2);
                }
            }
            // the nearest-centroid face is evaluated by every lane at the same time (convergent): tight running minimum
            if (kn >= 0) {
                const FacePre& fp = *reinterpret_cast<const FacePre*>(s_pre + (size_t)kn * FACEPRE_STRIDE);
                float d = tri_distance_pre(fp, v.p);
                int f = __float_as_int(s_cen[kn].w);
                if (v.best > d || (d == v.best && v.bi >= 0 && f < v.bi)) { v.best = d; v.bi = f; }
            }
            // ---- scan 2: remember the candidates the bounding sphere cannot reject against that bound ----
            int ns = 0;
            if (active) {
                // a face whose centroid is farther than sqrt(bound) + its bounding radius cannot beat the bound (it lies inside
                // the ball (centroid, radius)); the margins cover the rounding of this test, which only prunes and never decides
                const float bound = fminf(ub * 1.0001f, v.best);
                const float sb = sqrtf(bound) * 1.0002f;
DTB_UNROLL(PFD_SCAN_UNROLL)
                for (int i = 0; i < n; ++i) {
                    const int k = sub[i];
                    if (k == kn) continue;
                    float4 it = s_cen[k];
                    float dx = it.x - v.p[0], dy = it.y - v.p[1], dz = it.z - v.p[2];
                    const float rr = sb + fminf(s_pre[(size_t)k * FACEPRE_STRIDE + (FACEPRE_FLOATS - 1)], rmax);   // the face's own radius
                    if (dx * dx + dy * dy + dz * dz > rr * rr * 1.0002f) continue;
#if PFD_PLANE_PRUNE
                    {
                        // the reference distance is plane^2 + (in-plane term >= 0), both formed exactly as below: a face whose plane is
                        // farther than the bound cannot win (nor tie), whatever its in-plane term
                        const float* fq = s_pre + (size_t)k * FACEPRE_STRIDE;
                        const float nrm[3] = {fq[9], fq[10], fq[11]};
                        const float t = xsub(fq[12], xdot(nrm, v.p));
                        if (xmul(t, t) > bound) continue;
                    }
#endif
                    if (ns < PFD_LIST) { s_list[threadIdx.x][ns++] = (unsigned char)k; }
                    else {                                   // list full (rare): evaluate on the spot
                        const FacePre& fp = *reinterpret_cast<const FacePre*>(s_pre + (size_t)k * FACEPRE_STRIDE);
                        float d = tri_distance_pre(fp, v.p);
                        int f = __float_as_int(it.w);
                        if (v.best > d || (d == v.best && v.bi >= 0 && f < v.bi)) { v.best = d; v.bi = f; }
                    }
                }
            }
            // ---- pooled evaluation: the (query, candidate) pairs of the whole warp are evaluated 32 at a time, whoever owns
            //      them, and min-reduced per query with a 64-bit atomicMin on (distance bits, face id) ----
            s_best[warp][lane] = pack_df(v.best, v.bi);
            unsigned incl = (unsigned)ns;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            const unsigned tot_pairs = __shfl_sync(0xffffffffu, incl, 31);
            unsigned wofs = incl - (unsigned)ns;
            for (int j = 0; j < ns; ++j) s_pairs[warp][wofs + j] = (unsigned short)((lane << 8) | s_list[threadIdx.x][j]);
            __syncwarp();
            for (unsigned p0 = 0; p0 < tot_pairs; p0 += 32) {
                unsigned pi = p0 + lane;
                if (pi < tot_pairs) {
                    unsigned pr = s_pairs[warp][pi];
                    int ql = (int)(pr >> 8), k = (int)(pr & 255u);
                    float qp[3] = {s_q[warp][ql][0], s_q[warp][ql][1], s_q[warp][ql][2]};
                    const FacePre& fp = *reinterpret_cast<const FacePre*>(s_pre + (size_t)k * FACEPRE_STRIDE);
                    float d = tri_distance_pre(fp, qp);
                    // a face at MAX_DIS (invisible, k3 == 0) can never win the strict '<' against the initial 10000
                    if (d < FWD_MAX_DIS) atomicMin(&s_best[warp][ql], pack_df(d, __float_as_int(s_cen[k].w)));
                }
            }
            __syncwarp();
            {
                unsigned long long pb = s_best[warp][lane];
                float d = __uint_as_float((unsigned)(pb >> 32));
                int f = (int)(unsigned)(pb & 0xffffffffull);
                if (active) { v.best = d; v.bi = f; }        // s_best only ever decreased from (v.best, v.bi)
            }
        }
        if (active && !brute) {
            // the "always test" faces (unreliable xy-projection / tiny: geometric pruning by the centroid is not valid for them) are
            // looked at last, when the running minimum is tight: the reference distance is plane^2 + (in-plane term >= 0), so a face
            // whose plane is farther than the current best cannot win; the others are evaluated from the precomputed half
            for (int k = 0; k < na; ++k) {
                const int f = always[(size_t)b * always_cap + k];
                const float4* src = preb + (size_t)f * 8;
                const float4 w2 = __ldg(src + 2), w3 = __ldg(src + 3);          // floats 9..11 = unit normal, 12 = dot(n, a)
                const float nrm[3] = {w2.y, w2.z, w2.w};
                const float t = xsub(w3.x, xdot(nrm, v.p));
                if (xmul(t, t) > v.best) continue;
                __align__(16) float buf[FACEPRE_FLOATS];
#pragma unroll
                for (int m = 0; m < 8; ++m) reinterpret_cast<float4*>(buf)[m] = __ldg(src + m);
                const float d = tri_distance_pre(*reinterpret_cast<const FacePre*>(buf), v.p);
                if (v.best > d || (d == v.best && v.bi >= 0 && f < v.bi)) { v.best = d; v.bi = f; }
            }
        }
        if (active && !brute && nf > 0) {
            // certify against the faces that were not scanned.  (i) not staged: their centroids lie outside the staged reach region
            // (inside the 3x3x3 bricks); sides of it beyond the face grid's bounding box hold no face.  (ii) staged but not on the
            // warp's sub-list: farther than `margin` from every query of the warp.
            float db = 3.0e38f;
            const float o3[3] = {g.ox, g.oy, g.oz};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (rlo[k] > o3[k]) db = fminf(db, v.p[k] - rlo[k]);
                if (rhi[k] < o3[k] + gmax) db = fminf(db, rhi[k] - v.p[k]);
            }
            float lb = fmaxf(db - rmax - slack, 0.f) * 0.9999f;
            const float lm = margin * 0.999f;
            if (!(lb * lb > v.best) || !(lm * lm > v.best))
                brick_walk(v.p[0], v.p[1], v.p[2], g, G, rmax, cell_start, cell_end, sorted, mask, cell_base, v);
        }
        if (active) {
            closest_d[(size_t)b * S + orig] = v.best;
            closest_f[(size_t)b * S + orig] = (float)v.bi;
        }
    }
}

// ---- backward ---------------------------------------------------------------------------------------
// grads (3x3) of the reference distance w.r.t. the vertices of the closest face, times `scale`
__device__ __forceinline__ void tri_grad(const float* face, const float* p, float scale, float* g9) {
#pragma unroll
    for (int k = 0; k < 9; ++k) g9[k] = 0.f;
    TriHit h;
    tri_distance(face, face + 3, face + 6, p, BWD_MAX_DIS, h);
    if (h.type == 0) {                  // cuda_gradient_triangle_distance (_back.cu:440-483)
        float l[3] = {h.l1, h.l2, h.l3};
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
            for (int k = 0; k < 3; ++k) g9[v * 3 + k] = scale * (2.f * (h.ip[k] - p[k]) * l[v]);
    } else if (h.type == 1) {           // cuda_gradient_line_distance (_back.cu:291-317): grad[0..2] overwritten with the *t term
        int i1 = h.idx;
        int i2 = (i1 + 1) % 3;
        const float* A = face + i1 * 3;
        const float* B = face + i2 * 3;
        float PA[3] = {xsub(p[0], A[0]), xsub(p[1], A[1]), xsub(p[2], A[2])};
        float BA[3] = {xsub(B[0], A[0]), xsub(B[1], A[1]), xsub(B[2], A[2])};
        float t = xdiv(xdot(PA, BA), div_nz(xdot(BA, BA)));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // non-contracted like the source (_back.cu:309-315): q - p cancels to ~sqrt(d), so an FMA here shows up as a relative
            // error of up to 1e-3 in the gradient of points that lie almost on the surface (found by the res-70 scale parity run)
            float q = xadd(xmul(A[k], xsub(1.f, t)), xmul(B[k], t));
            g9[i1 * 3 + k] = scale * (2.f * xsub(q, p[k]) * t);
        }
    } else if (h.type == 2) {
        int iv = h.idx;
#pragma unroll
        for (int k = 0; k < 3; ++k) g9[iv * 3 + k] = 2.f * scale * (face[iv * 3 + k] - p[k]);
    }
}

// drop-in: atomics into dldface (B,F,3,3), upstream dl_dd (B,S)
__global__ void __launch_bounds__(256) pfd_backward_kernel(const float* __restrict__ points, int S, const float* __restrict__ soup, int Fmax,
                                                           const float* __restrict__ closest_f, const float* __restrict__ dl_dd,
                                                           float* __restrict__ dldface) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    size_t o = (size_t)b * S + i;
    int f = (int)closest_f[o];
    if (f < 0 || f >= Fmax) return;                 // the reference would read out of bounds for -1
    const float* face = soup + ((size_t)b * Fmax + f) * 9;
    float fc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) fc[k] = face[k];
    float p[3] = {points[o * 3], points[o * 3 + 1], points[o * 3 + 2]};
    float g9[9];
    tri_grad(fc, p, dl_dd[o], g9);
    float* g = dldface + ((size_t)b * Fmax + f) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k)
        if (g9[k] != 0.f) atomicAdd(g + k, g9[k]);
}

// engine: loss_b = mean_i sqrt(d_i + 1e-10) (mesh_utils.py:373, deftet.py:181); scatter straight to grad_pos
__global__ void __launch_bounds__(256) pfd_backward_indexed_kernel(const float* __restrict__ points, int S, const float* __restrict__ soup,
                                                                   const int32_t* __restrict__ faces, int Fmax, int V,
                                                                   const float* __restrict__ closest_f, const float* __restrict__ closest_d,
                                                                   const float* __restrict__ g_loss, float* __restrict__ grad_pos, int gstride) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    size_t o = (size_t)b * S + i;
    int f = (int)closest_f[o];
    if (f < 0 || f >= Fmax) return;
    const float* face = soup + ((size_t)b * Fmax + f) * 9;
    float fc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) fc[k] = face[k];
    float p[3] = {points[o * 3], points[o * 3 + 1], points[o * 3 + 2]};
    float scale = g_loss[b] / ((float)S * 2.f * sqrtf(closest_d[o] + 1e-10f));
    float g9[9];
    tri_grad(fc, p, scale, g9);
    const int32_t* fi = faces + ((size_t)b * Fmax + f) * 3;
    float* gp = grad_pos + (size_t)b * V * gstride;
#pragma unroll
    for (int v = 0; v < 3; ++v)
        if (g9[v * 3] != 0.f || g9[v * 3 + 1] != 0.f || g9[v * 3 + 2] != 0.f) grad_add3(gp, (size_t)fi[v], gstride, g9[v * 3], g9[v * 3 + 1], g9[v * 3 + 2]);
}

__global__ void sqrt_mean_kernel(const float* __restrict__ d, int S, float eps, double* __restrict__ acc) {
    int b = blockIdx.y;
    double s = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) s += (double)sqrtf(d[(size_t)b * S + i] + eps);
    s = warp_sum(s);
    __shared__ double sw[8];
    if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sw[w];
        atomicAdd(acc + b, t);
    }
}
__global__ void sqrt_mean_finalize_kernel(const double* __restrict__ acc, const int32_t* __restrict__ counts, int S, int B, float* __restrict__ out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    out[b] = (counts && counts[b] == 0) ? 1.0f : (float)(acc[b] / (double)S);      // empty surface -> 1 (deftet.py:162-166)
}

}  // namespace dtb

using namespace dtb;

constexpr int PFD_ALWAYS_CAP = 256;

__global__ void pfd_fill_none_kernel(float* __restrict__ d, float* __restrict__ f, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { d[i] = 10000.0f; f[i] = -1.0f; }
}

extern "C" int dtb_point_face_distance_grid_res(int Fmax) {
    int g = (int)ceil(sqrt((double)(Fmax > 1 ? Fmax : 1)) * 0.375);      // res-70 sweep with cell-sorted queries (tools/r2_time.py a4): 48 at Fmax = 16384
    g = (g + 3) / 4 * 4;
    if (g < 4) g = 4;
    if (g > 128) g = 128;
    return g;
}
extern "C" size_t dtb_point_face_distance_workspace(int B, int S, int Fmax, int G) {
    if (G <= 0) G = dtb_point_face_distance_grid_res(Fmax);
    G = (G + 3) / 4 * 4;
    size_t nbr = (size_t)B * G * G * G;          // query bins: one per cell
    return pointgrid_workspace_bytes(B, Fmax, G, true, true) + align_up((size_t)B * PFD_ALWAYS_CAP * 4, 256) + 1024 +
           2 * align_up(nbr * 4, 256) + align_up((size_t)B * S * 4, 256) + align_up((size_t)B * S * 16, 256) + scan_workspace_bytes(nbr) + 256 +
           align_up((size_t)B * Fmax * 128, 256);
}

// counts (B,) i32: number of valid faces of each sample (the reference passes it as float n_face_b).
extern "C" int dtb_point_face_distance_forward(const float* points, const float* faces, const int32_t* counts, int B, int S, int Fmax,
                                               int G, float* closest_d, float* closest_f, void* workspace, size_t workspace_bytes,
                                               void* stream) {
    DTB_REQUIRE(points && counts && closest_d && closest_f, "point_face_distance_forward: null argument");
    DTB_REQUIRE(B > 0 && S >= 0 && Fmax >= 0, "point_face_distance_forward: bad sizes");
    if (S == 0) return DTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (Fmax == 0) {                       // no faces at all: min_d = 10000, min_idx = -1 (for.cu:278-279)
        pfd_fill_none_kernel<<<cdiv((long long)B * S, 256), 256, 0, st>>>(closest_d, closest_f, (size_t)B * S);
        DTB_LAUNCH_CHECK("pfd_fill_none");
        return DTB_OK;
    }
    if (G <= 0) G = dtb_point_face_distance_grid_res(Fmax);
    G = (G + 3) / 4 * 4;
    Workspace ws(workspace, workspace_bytes);
    PointGrid pg;
    pointgrid_carve(pg, B, Fmax, G, true, true, ws);
    int32_t* always = ws.take<int32_t>((size_t)B * PFD_ALWAYS_CAP);
    unsigned* rmax = ws.take<unsigned>(B);
    int32_t* n_always = ws.take<int32_t>(B);
    const size_t nbr = (size_t)B * G * G * G;    // query bins: one per cell of the face grid
    unsigned* qstart = ws.take<unsigned>(nbr);
    unsigned* qend = ws.take<unsigned>(nbr);
    unsigned* qbrick = ws.take<unsigned>((size_t)B * S);
    float4* qsorted = ws.take<float4>((size_t)B * S);
    size_t qsb = scan_workspace_bytes(nbr);
    void* qsws = ws.take<char>(qsb);
    float4* pre = ws.take<float4>((size_t)B * Fmax * 8);
    if (!ws.ok || !workspace) { set_error("point_face_distance: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    DTB_CUDA(cudaMemsetAsync(rmax, 0, B * sizeof(unsigned), st));
    DTB_CUDA(cudaMemsetAsync(n_always, 0, B * sizeof(int32_t), st));
    {
        int rc = pointgrid_build_ragged(pg, faces, true, counts, st);
        if (rc) return rc;
        dim3 gs(cdiv(Fmax, 256), B);
        face_stats_kernel<<<gs, 256, 0, st>>>(faces, counts, Fmax, rmax, always, n_always, PFD_ALWAYS_CAP, pre);
        DTB_LAUNCH_CHECK("face_stats");
    }
    {
        dim3 gq(cdiv(S, 256), B);
        DTB_CUDA(cudaMemsetAsync(qstart, 0, nbr * sizeof(unsigned), st));
        qbin_count_kernel<<<gq, 256, 0, st>>>(points, S, G, pg.bbox_ord, qstart, qbrick);
        DTB_LAUNCH_CHECK("qbin_count");
        int rc = exclusive_scan_u32_dup(qstart, qstart, qend, nbr, nullptr, qsws, qsb, st);
        if (rc) return rc;
        qbin_fill_kernel<<<gq, 256, 0, st>>>(points, S, qbrick, qend, qsorted);
        DTB_LAUNCH_CHECK("qbin_fill");
    }
    const int NB = G / 4;
    dim3 grid(NB * NB * NB, B);
    prof_begin(PROF_PFD_FORWARD, st);
    pfd_forward_tiled_kernel<<<grid, PFD_THREADS, 0, st>>>(S, faces, pre, counts, Fmax, G, pg.bbox_ord, pg.cell_start, pg.cell_end, pg.sorted,
                                                           pg.mask, rmax, always, n_always, PFD_ALWAYS_CAP, qstart, qend, qsorted,
                                                           closest_d, closest_f);
    DTB_LAUNCH_CHECK("pfd_forward_tiled");
    prof_end(PROF_PFD_FORWARD, st);
    return DTB_OK;
}

extern "C" int dtb_point_face_distance_backward(const float* points, const float* faces, const float* closest_f, const float* dl_dd, int B,
                                                int S, int Fmax, float* dldface, void* stream) {
    DTB_REQUIRE(points && closest_f && dl_dd && dldface, "point_face_distance_backward: null argument");
    if (B == 0 || S == 0 || Fmax == 0) return DTB_OK;
    dim3 grid(cdiv(S, 256), B);
    pfd_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, S, faces, Fmax, closest_f, dl_dd, dldface);
    DTB_LAUNCH_CHECK("pfd_backward");
    return DTB_OK;
}

extern "C" int dtb_point_face_distance_backward_indexed(const float* points, const float* soup, const int32_t* faces, const float* closest_f,
                                                        const float* closest_d, const float* g_loss, int B, int S, int Fmax, int V,
                                                        float* grad_pos, int grad_stride, void* stream) {
    DTB_REQUIRE(points && soup && faces && closest_f && closest_d && g_loss && grad_pos, "point_face_distance_backward_indexed: null argument");
    if (B == 0 || S == 0 || Fmax == 0) return DTB_OK;
    dim3 grid(cdiv(S, 256), B);
    DTB_REQUIRE(grad_stride == 3 || (grad_stride == 4 && (((size_t)grad_pos) & 15) == 0), "point_face_distance_backward_indexed: bad grad_stride / alignment");
    pfd_backward_indexed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, S, soup, faces, Fmax, V, closest_f, closest_d, g_loss, grad_pos, grad_stride);
    DTB_LAUNCH_CHECK("pfd_backward_indexed");
    return DTB_OK;
}

// out[b] = mean_i sqrt(d[b,i] + eps)   (counts may be NULL; counts[b]==0 -> 1)
extern "C" int dtb_sqrt_mean(const float* d, const int32_t* counts, int B, int S, float eps, double* acc, float* out, void* stream) {
    DTB_REQUIRE(d && acc && out, "sqrt_mean: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    DTB_CUDA(cudaMemsetAsync(acc, 0, B * sizeof(double), st));
    if (S > 0) {
        dim3 grid(min(cdiv(S, 256), 128), B);
        sqrt_mean_kernel<<<grid, 256, 0, st>>>(d, S, eps, acc);
        DTB_LAUNCH_CHECK("sqrt_mean");
    }
    sqrt_mean_finalize_kernel<<<cdiv(B, 64), 64, 0, st>>>(acc, counts, S, B, out);
    DTB_LAUNCH_CHECK("sqrt_mean_finalize");
    return DTB_OK;
}
