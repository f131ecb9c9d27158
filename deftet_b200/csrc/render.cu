// A15: per-pixel tetrahedral-face volume rasterizer -- stand-in for kal.render.mesh.deftet_sparse_render, the
// call at diff_render/diftet_6_subdiv/5_rendereq/deftetrneder.py:97-100 (Kaolin is third-party, un-vendored and
// un-pinned in the reference: "parity unpinned"; the semantics below are the call-site contract + SURVEY.md 8c):
//   for every pixel: all faces whose 2-D triangle contains the pixel (w1 = k1/(k3+eps), w2 = k2/(k3+eps),
//   w0 = 1-w1-w2, reject if any w < 0) and whose interpolated z lies inside the pixel's [min, max] range; the
//   first K such faces in ascending face id, sorted near-to-far (z descending: the camera looks down -z);
//   output barycentrically interpolated features (B,P,K,d) and face ids (B,P,K) int64, void slots = 0 / -1.
//   backward: gradients to face_features and face_vertices_image (not to z, not to the pixel).
// B200 design: faces are binned into a 2-D grid over the pixels' bounding box ((cell, face) pairs, radix-sorted so
// that every cell lists its faces in ascending id); one WARP per pixel streams its cell's list with coalesced
// loads, appends hits in order with ballot/popc, bitonic-sorts them by depth in shared memory and writes the K
// slots of the pixel with coalesced stores (the dense (B,P,K,d) output is the bandwidth hog: 4.6 GB per 800x800
// view at K=300).
#include "prims.cuh"
#include "deftet_b200.h"

namespace dtb {

typedef unsigned long long u64;

struct PixGrid { float ox, oy, inv; int R; };
__device__ __forceinline__ PixGrid pix_grid(const unsigned* bbox_ord, int b, int R) {
    const unsigned* q = bbox_ord + (size_t)b * 4;
    float x0 = ord2f(q[0]), y0 = ord2f(q[1]), x1 = ord2f(q[2]), y1 = ord2f(q[3]);
    float ext = fmaxf(fmaxf(x1 - x0, y1 - y0), 1e-20f) * (1.0f + 1e-6f);
    PixGrid g; g.ox = x0; g.oy = y0; g.inv = (float)R / ext; g.R = R;
    return g;
}
__device__ __forceinline__ int pix_cell(float x, float o, float inv, int R) {
    float f = floorf((x - o) * inv);
    f = fminf(fmaxf(f, 0.f), (float)(R - 1));
    return (int)f;
}

__global__ void rd_init_kernel(unsigned* bbox_ord, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 4) bbox_ord[i] = ((i % 4) < 2) ? 0xffffffffu : 0u;
}
__global__ void __launch_bounds__(256) rd_bbox_kernel(const float* __restrict__ pix, int P, unsigned* __restrict__ bbox_ord) {
    int b = blockIdx.y;
    float mn[2] = {3.4e38f, 3.4e38f}, mx[2] = {-3.4e38f, -3.4e38f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        float x = pix[((size_t)b * P + i) * 2], y = pix[((size_t)b * P + i) * 2 + 1];
        mn[0] = fminf(mn[0], x); mx[0] = fmaxf(mx[0], x); mn[1] = fminf(mn[1], y); mx[1] = fmaxf(mx[1], y);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (mn[k] <= mx[k]) { atomicMin(&bbox_ord[(size_t)b * 4 + k], f2ord(mn[k])); atomicMax(&bbox_ord[(size_t)b * 4 + 2 + k], f2ord(mx[k])); }
    }
}

// cells overlapped by a face's xy bounding box (clipped to the pixel grid); faces entirely outside are dropped
__device__ __forceinline__ bool face_cells(const float* xy, const PixGrid& g, const unsigned* bbox_ord, int b, int& cx0, int& cx1, int& cy0, int& cy1) {
    float x0 = fminf(xy[0], fminf(xy[2], xy[4])), x1 = fmaxf(xy[0], fmaxf(xy[2], xy[4]));
    float y0 = fminf(xy[1], fminf(xy[3], xy[5])), y1 = fmaxf(xy[1], fmaxf(xy[3], xy[5]));
    const unsigned* q = bbox_ord + (size_t)b * 4;
    if (!(x1 >= ord2f(q[0]) && x0 <= ord2f(q[2]) && y1 >= ord2f(q[1]) && y0 <= ord2f(q[3]))) return false;
    cx0 = pix_cell(x0, g.ox, g.inv, g.R); cx1 = pix_cell(x1, g.ox, g.inv, g.R);
    cy0 = pix_cell(y0, g.oy, g.inv, g.R); cy1 = pix_cell(y1, g.oy, g.inv, g.R);
    return true;
}
__global__ void __launch_bounds__(256) rd_count_kernel(const float* __restrict__ face_xy, int F, int R, const unsigned* __restrict__ bbox_ord,
                                                       unsigned* __restrict__ npairs) {
    int b = blockIdx.y;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    PixGrid g = pix_grid(bbox_ord, b, R);
    int cx0, cx1, cy0, cy1;
    unsigned n = 0;
    if (face_cells(face_xy + ((size_t)b * F + f) * 6, g, bbox_ord, b, cx0, cx1, cy0, cy1)) n = (unsigned)((cx1 - cx0 + 1) * (cy1 - cy0 + 1));
    npairs[(size_t)b * F + f] = n;
}
__global__ void __launch_bounds__(256) rd_pairs_kernel(const float* __restrict__ face_xy, int F, int R, const unsigned* __restrict__ bbox_ord,
                                                       const unsigned* __restrict__ pair_off, size_t cap, u64* __restrict__ keys,
                                                       unsigned* __restrict__ vals, unsigned* __restrict__ cell_cnt) {
    int b = blockIdx.y;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    PixGrid g = pix_grid(bbox_ord, b, R);
    int cx0, cx1, cy0, cy1;
    if (!face_cells(face_xy + ((size_t)b * F + f) * 6, g, bbox_ord, b, cx0, cx1, cy0, cy1)) return;
    size_t o = pair_off[(size_t)b * F + f];
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
            u64 cell = ((u64)b * R + cy) * R + cx;
            if (o < cap) { keys[o] = cell * (u64)F + (u64)f; vals[o] = (unsigned)f; }
            atomicAdd(cell_cnt + cell, 1u);
            ++o;
        }
}
// pad keys beyond the real pair count so that they sort to the end (pad_key = one past the largest valid key)
__global__ void rd_pad_kernel(u64* keys, unsigned* vals, const unsigned* total, size_t cap, u64 pad_key) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap && i >= *total) { keys[i] = pad_key; vals[i] = 0xffffffffu; }
}

// cell_end[c] = cell_start[c+1] (the pair list is sorted by cell); flags overflow of the pair capacity
__global__ void rd_cell_end_kernel(const unsigned* __restrict__ cstart, const unsigned* __restrict__ total, size_t cells, size_t cap,
                                   unsigned* __restrict__ cend, int32_t* __restrict__ overflow) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    unsigned t = *total;
    if (c == 0) *overflow = (t > cap) ? 1 : 0;
    unsigned e = (c + 1 < cells) ? cstart[c + 1] : t;
    unsigned lim = (unsigned)min((size_t)t, cap);
    cend[c] = min(e, lim);
}

struct Bary2 { float w0, w1, w2, depth; bool hit; };
__device__ __forceinline__ Bary2 face_test(const float* xy, const float* z, float px, float py, float zmin, float zmax, float eps) {
    Bary2 r; r.hit = false;
    float ax = xy[0], ay = xy[1], bx = xy[2], by = xy[3], cx = xy[4], cy = xy[5];
    if (px < fminf(ax, fminf(bx, cx)) || px > fmaxf(ax, fmaxf(bx, cx)) || py < fminf(ay, fminf(by, cy)) || py > fmaxf(ay, fmaxf(by, cy))) return r;
    // hit / miss and the depth order are decided with non-contracted fp32 (bit-identical to the CPU oracle)
    float m = xsub(bx, ax), pp = xsub(by, ay), n = xsub(cx, ax), q = xsub(cy, ay), s = xsub(px, ax), t = xsub(py, ay);
    float k1 = xsub(xmul(s, q), xmul(n, t)), k2 = xsub(xmul(m, t), xmul(s, pp)), k3 = xsub(xmul(m, q), xmul(n, pp));
    float den = xadd(k3, eps);
    r.w1 = xdiv(k1, den); r.w2 = xdiv(k2, den); r.w0 = xsub(xsub(1.f, r.w1), r.w2);
    if (r.w0 < 0.f || r.w1 < 0.f || r.w2 < 0.f) return r;
    r.depth = xadd(xadd(xmul(r.w0, z[0]), xmul(r.w1, z[1])), xmul(r.w2, z[2]));
    if (!(r.depth >= zmin && r.depth <= zmax)) return r;
    r.hit = true;
    return r;
}

constexpr int RD_WARPS = 4;

// Warp-cooperative hit collection for one pixel: scan the pixel's cell list in ascending face id, append hits with
// ballot/popc (order preserved, at most K), then bitonic-sort them by (depth descending, face id ascending) in the warp's
// shared-memory arrays.  Returns the number of hits.
__device__ __forceinline__ int rd_collect_sorted(float px, float py, float zmin, float zmax, const float* __restrict__ fxy,
                                                 const float* __restrict__ fz, const unsigned* __restrict__ cell_faces, unsigned j0,
                                                 unsigned j1, int K, float eps, float* s_depth, int* s_face, float* s_w1, float* s_w2) {
    const int lane = threadIdx.x & 31;
    int count = 0;
    for (unsigned jb = j0; jb < j1 && count < K; jb += 32) {
        unsigned j = jb + lane;
        Bary2 r; r.hit = false;
        int f = -1;
        if (j < j1) {
            f = (int)cell_faces[j];
            r = face_test(fxy + (size_t)f * 6, fz + (size_t)f * 3, px, py, zmin, zmax, eps);
        }
        unsigned hits = __ballot_sync(0xffffffffu, r.hit);
        if (r.hit) {
            int slot = count + __popc(hits & ((1u << lane) - 1u));
            if (slot < K) { s_depth[slot] = r.depth; s_face[slot] = f; s_w1[slot] = r.w1; s_w2[slot] = r.w2; }
        }
        count = min(count + __popc(hits), K);
    }
    __syncwarp();
    int n2 = 1;
    while (n2 < count) n2 <<= 1;
    for (int i = count + lane; i < n2; i += 32) { s_depth[i] = -3.4e38f; s_face[i] = -1; s_w1[i] = 0.f; s_w2[i] = 0.f; }
    __syncwarp();
    // every lane owns one compare-exchange per step: pair index p -> (i, l = i | j) with bit log2(j) of i cleared
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = lane; p < (n2 >> 1); p += 32) {
                int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                int l = i | j;
                bool up = (i & k) == 0;              // "up" = this half sorted in final (descending-depth) order
                float da = s_depth[i], db = s_depth[l];
                int fa = s_face[i], fb = s_face[l];
                // a comes first when its depth is larger, or equal depth and smaller face id (void = -1 sorts last)
                bool a_first = (da > db) || (da == db && (unsigned)fa < (unsigned)fb);
                if (a_first != up) {
                    s_depth[i] = db; s_depth[l] = da; s_face[i] = fb; s_face[l] = fa;
                    float t = s_w1[i]; s_w1[i] = s_w1[l]; s_w1[l] = t;
                    t = s_w2[i]; s_w2[i] = s_w2[l]; s_w2[l] = t;
                }
            }
            __syncwarp();
        }
    return count;
}

// one warp per pixel.  dynamic smem: RD_WARPS * Kpad * (depth f32, face i32, w1 f32, w2 f32)
__global__ void __launch_bounds__(RD_WARPS * 32) rd_forward_kernel(
    const float* __restrict__ pix, const float* __restrict__ ranges, const float* __restrict__ face_z, const float* __restrict__ face_xy,
    const float* __restrict__ face_feat, int P, int F, int D, int K, int Kpad, float eps, int R, const unsigned* __restrict__ bbox_ord,
    const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_end, const unsigned* __restrict__ cell_faces,
    float* __restrict__ out_feat, long long* __restrict__ out_idx) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const long long p = (long long)blockIdx.x * RD_WARPS + warp;
    if (p >= P) return;
    float* s_depth = smem + (size_t)warp * Kpad * 4;
    int* s_face = (int*)(s_depth + Kpad);
    float* s_w1 = s_depth + 2 * Kpad;
    float* s_w2 = s_depth + 3 * Kpad;
    const size_t po = (size_t)b * P + p;
    const float px = pix[po * 2], py = pix[po * 2 + 1];
    PixGrid g = pix_grid(bbox_ord, b, R);
    const size_t cell = ((size_t)b * R + pix_cell(py, g.oy, g.inv, R)) * R + pix_cell(px, g.ox, g.inv, R);
    const int count = rd_collect_sorted(px, py, ranges[po * 2], ranges[po * 2 + 1], face_xy + (size_t)b * F * 6, face_z + (size_t)b * F * 3,
                                        cell_faces, cell_start[cell], cell_end[cell], K, eps, s_depth, s_face, s_w1, s_w2);
    // write the K slots of this pixel
    long long* oi = out_idx + po * K;
    float* of = out_feat + po * (size_t)K * D;
    const float* ff = face_feat + (size_t)b * F * 3 * D;
    const bool vec4 = (D == 4) && ((((size_t)of) | ((size_t)ff)) & 15) == 0;
    if (vec4) {                                           // RGBA features: one 16-byte store per slot
        float4* of4 = reinterpret_cast<float4*>(of);
        for (int k = lane; k < K; k += 32) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            long long id = -1ll;
            if (k < count) {
                int f = s_face[k];
                id = f;
                float w1 = s_w1[k], w2 = s_w2[k], w0 = 1.f - w1 - w2;
                const float4* q = reinterpret_cast<const float4*>(ff + (size_t)f * 12);
                float4 a = __ldg(q), bq = __ldg(q + 1), c = __ldg(q + 2);
                o.x = w0 * a.x + w1 * bq.x + w2 * c.x; o.y = w0 * a.y + w1 * bq.y + w2 * c.y;
                o.z = w0 * a.z + w1 * bq.z + w2 * c.z; o.w = w0 * a.w + w1 * bq.w + w2 * c.w;
            }
            of4[k] = o;
            oi[k] = id;
        }
    } else {
        for (int k = lane; k < K; k += 32) oi[k] = (k < count) ? (long long)s_face[k] : -1ll;
        for (int e = lane; e < K * D; e += 32) {
            int k = e / D, c = e - k * D;
            float v = 0.f;
            if (k < count) {
                int f = s_face[k];
                float w1 = s_w1[k], w2 = s_w2[k], w0 = 1.f - w1 - w2;
                const float* q = ff + (size_t)f * 3 * D;
                v = w0 * q[c] + w1 * q[D + c] + w2 * q[2 * D + c];
            }
            of[e] = v;
        }
    }
}

// ---- fused render + composite ("peel2mask", 5_rendereq/deftetrneder.py:31-64) ------------------------------------------
// out_color (B,P,D-1) = sum_k vis_k c_k + (1 - sum_k vis_k) (white background), out_mask (B,P,1) = sum_k vis_k with
// alpha_k = clamp(f_k[0], 1e-10, 1-1e-10), vis_k = alpha_k prod_{i<k} (1 - alpha_i) over the K depth-sorted slots (void slots
// have f = 0, i.e. alpha = 1e-10: their tiny contribution is kept).  The (B,P,K,D) tensor is never materialised.
constexpr int RC_MAXD = 8;
constexpr float RC_EPS = 1e-10f;

__device__ __forceinline__ void rc_slot_features(const float* __restrict__ ff, int f, float w1, float w2, int D, float* out) {
    const float* q = ff + (size_t)f * 3 * D;
    const float w0 = 1.f - w1 - w2;
    for (int c = 0; c < D; ++c) out[c] = w0 * q[c] + w1 * q[D + c] + w2 * q[2 * D + c];
}

__global__ void __launch_bounds__(RD_WARPS * 32) rc_forward_kernel(
    const float* __restrict__ pix, const float* __restrict__ ranges, const float* __restrict__ face_z, const float* __restrict__ face_xy,
    const float* __restrict__ face_feat, int P, int F, int D, int K, int Kpad, float eps, int R, const unsigned* __restrict__ bbox_ord,
    const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_end, const unsigned* __restrict__ cell_faces,
    float* __restrict__ out_color, float* __restrict__ out_mask) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const long long p = (long long)blockIdx.x * RD_WARPS + warp;
    if (p >= P) return;
    float* s_depth = smem + (size_t)warp * Kpad * 4;
    int* s_face = (int*)(s_depth + Kpad);
    float* s_w1 = s_depth + 2 * Kpad;
    float* s_w2 = s_depth + 3 * Kpad;
    const size_t po = (size_t)b * P + p;
    const float px = pix[po * 2], py = pix[po * 2 + 1];
    PixGrid g = pix_grid(bbox_ord, b, R);
    const size_t cell = ((size_t)b * R + pix_cell(py, g.oy, g.inv, R)) * R + pix_cell(px, g.ox, g.inv, R);
    const int count = rd_collect_sorted(px, py, ranges[po * 2], ranges[po * 2 + 1], face_xy + (size_t)b * F * 6, face_z + (size_t)b * F * 3,
                                        cell_faces, cell_start[cell], cell_end[cell], K, eps, s_depth, s_face, s_w1, s_w2);
    const float* ff = face_feat + (size_t)b * F * 3 * D;
    float T = 1.f, mask = 0.f;
    float col[RC_MAXD];
    for (int c = 0; c < RC_MAXD; ++c) col[c] = 0.f;
    for (int base = 0; base < count; base += 32) {
        const int k = base + lane;
        const bool valid = k < count;
        float f[RC_MAXD];
        float alpha = 0.f;
        if (valid) { rc_slot_features(ff, s_face[k], s_w1[k], s_w2[k], D, f); alpha = fminf(fmaxf(f[0], RC_EPS), 1.0f - RC_EPS); }
        float prod = valid ? 1.f - alpha : 1.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { float y = __shfl_up_sync(0xffffffffu, prod, o); if (lane >= o) prod *= y; }
        float excl = __shfl_up_sync(0xffffffffu, prod, 1);
        if (lane == 0) excl = 1.f;
        const float vis = valid ? alpha * T * excl : 0.f;
        mask += vis;
        for (int c = 1; c < D; ++c) col[c] += valid ? vis * f[c] : 0.f;
        T *= __shfl_sync(0xffffffffu, prod, 31);
    }
    mask = warp_sum(mask);
    for (int c = 1; c < D; ++c) col[c] = warp_sum(col[c]);
    mask += (float)(K - count) * RC_EPS * T;                 // void slots: alpha = clamp(0) = 1e-10, (1 - 1e-10) == 1 in fp32
    if (lane == 0) {
        out_mask[po] = mask;
        for (int c = 1; c < D; ++c) out_color[po * (D - 1) + (c - 1)] = col[c] + (1.f - mask);
    }
}

// backward of the fused op: recollect the pixel's sorted hits, redo the front-to-back scan, run it backwards, scatter.
// dynamic smem: RD_WARPS * Kpad * 6 floats (depth->alpha, face, w1, w2, T, vis*G)
__global__ void __launch_bounds__(RD_WARPS * 32) rc_backward_kernel(
    const float* __restrict__ pix, const float* __restrict__ ranges, const float* __restrict__ face_z, const float* __restrict__ face_xy,
    const float* __restrict__ face_feat, int P, int F, int D, int K, int Kpad, float eps, int R, const unsigned* __restrict__ bbox_ord,
    const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_end, const unsigned* __restrict__ cell_faces,
    const float* __restrict__ g_color, const float* __restrict__ g_mask, float* __restrict__ g_xy, float* __restrict__ g_feat) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const long long p = (long long)blockIdx.x * RD_WARPS + warp;
    if (p >= P) return;
    float* s_depth = smem + (size_t)warp * Kpad * 6;
    int* s_face = (int*)(s_depth + Kpad);
    float* s_w1 = s_depth + 2 * Kpad;
    float* s_w2 = s_depth + 3 * Kpad;
    float* s_T = s_depth + 4 * Kpad;
    float* s_vg = s_depth + 5 * Kpad;
    const size_t po = (size_t)b * P + p;
    const float px = pix[po * 2], py = pix[po * 2 + 1];
    PixGrid g = pix_grid(bbox_ord, b, R);
    const size_t cell = ((size_t)b * R + pix_cell(py, g.oy, g.inv, R)) * R + pix_cell(px, g.ox, g.inv, R);
    const float* fxy = face_xy + (size_t)b * F * 6;
    const int count = rd_collect_sorted(px, py, ranges[po * 2], ranges[po * 2 + 1], fxy, face_z + (size_t)b * F * 3, cell_faces,
                                        cell_start[cell], cell_end[cell], K, eps, s_depth, s_face, s_w1, s_w2);
    if (count == 0) return;                                   // only void slots: no face receives a gradient
    const float* ff = face_feat + (size_t)b * F * 3 * D;
    float gc[RC_MAXD];
    float gsum = 0.f;
    for (int c = 1; c < D; ++c) { gc[c] = g_color[po * (D - 1) + (c - 1)]; gsum += gc[c]; }
    const float gM = g_mask[po] - gsum;                       // out_color = C + (1 - M): dL/dM = g_mask - sum_c g_color_c
    float* s_alpha = s_depth;                                 // depth is no longer needed after the sort
    // pass 1 (front to back): alpha_k, T_k and vis_k * G_k with G_k = dL/dvis_k = sum_c g_color_c c_k,c + gM
    float T = 1.f;
    for (int base = 0; base < count; base += 32) {
        const int k = base + lane;
        const bool valid = k < count;
        float f[RC_MAXD];
        float alpha = 0.f, G = gM;
        if (valid) {
            rc_slot_features(ff, s_face[k], s_w1[k], s_w2[k], D, f);
            alpha = fminf(fmaxf(f[0], RC_EPS), 1.0f - RC_EPS);
            for (int c = 1; c < D; ++c) G += gc[c] * f[c];
        }
        float prod = valid ? 1.f - alpha : 1.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { float y = __shfl_up_sync(0xffffffffu, prod, o); if (lane >= o) prod *= y; }
        float excl = __shfl_up_sync(0xffffffffu, prod, 1);
        if (lane == 0) excl = 1.f;
        if (valid) { s_alpha[k] = alpha; s_T[k] = T * excl; s_vg[k] = alpha * T * excl * G; }
        T *= __shfl_sync(0xffffffffu, prod, 31);
    }
    __syncwarp();
    // pass 2 (back to front): S_k = sum_{j>k} vis_j G_j (+ the void slots), d alpha_k = T_k G_k - S_k / (1 - alpha_k)
    float carry = (float)(K - count) * RC_EPS * T * gM;
    const int last_base = ((count - 1) / 32) * 32;
    for (int base = last_base; base >= 0; base -= 32) {
        const int k = base + lane;
        const bool valid = k < count;
        float v = valid ? s_vg[k] : 0.f;
        float suf = v;                                        // inclusive suffix sum across lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { float y = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += y; }
        const float S = carry + suf - v;                      // strictly after k
        carry += __shfl_sync(0xffffffffu, suf, 0);
        if (!valid) continue;
        const int f = s_face[k];
        const float w1 = s_w1[k], w2 = s_w2[k], w0 = 1.f - w1 - w2;
        float fv[RC_MAXD];
        rc_slot_features(ff, f, w1, w2, D, fv);
        const float alpha = s_alpha[k], Tk = s_T[k];
        float G = gM;
        for (int c = 1; c < D; ++c) G += gc[c] * fv[c];
        float df[RC_MAXD];
        const bool pass = (fv[0] >= RC_EPS) && (fv[0] <= 1.0f - RC_EPS);       // torch.clamp passes the gradient inside [min, max]
        df[0] = pass ? (Tk * G - S / (1.f - alpha)) : 0.f;
        const float vis = alpha * Tk;
        for (int c = 1; c < D; ++c) df[c] = gc[c] * vis;
        // scatter to the face's features and, through the weights, to its image-space vertices
        const float* q = ff + (size_t)f * 3 * D;
        float gw0 = 0.f, gw1 = 0.f, gw2 = 0.f;
        for (int c = 0; c < D; ++c) { gw0 += df[c] * q[c]; gw1 += df[c] * q[D + c]; gw2 += df[c] * q[2 * D + c]; }
        if (g_feat) {
            float* gf = g_feat + ((size_t)b * F + f) * 3 * D;
            if (D == 4 && (((size_t)gf) & 15) == 0) {
                float4* gf4 = reinterpret_cast<float4*>(gf);
                atomicAdd(gf4 + 0, make_float4(w0 * df[0], w0 * df[1], w0 * df[2], w0 * df[3]));
                atomicAdd(gf4 + 1, make_float4(w1 * df[0], w1 * df[1], w1 * df[2], w1 * df[3]));
                atomicAdd(gf4 + 2, make_float4(w2 * df[0], w2 * df[1], w2 * df[2], w2 * df[3]));
            } else {
                for (int c = 0; c < D; ++c) { atomicAdd(gf + c, w0 * df[c]); atomicAdd(gf + D + c, w1 * df[c]); atomicAdd(gf + 2 * D + c, w2 * df[c]); }
            }
        }
        if (g_xy) {
            const float* xy = fxy + (size_t)f * 6;
            float ax = xy[0], ay = xy[1], bx = xy[2], by = xy[3], cx = xy[4], cy = xy[5];
            float m = bx - ax, pp = by - ay, n = cx - ax, qq = cy - ay, s = px - ax, t = py - ay;
            float k1 = s * qq - n * t, k2 = m * t - s * pp, k3 = m * qq - n * pp;
            float inv = 1.f / (k3 + eps);
            float a1 = gw1 - gw0, a2 = gw2 - gw0;
            float gk1 = a1 * inv, gk2 = a2 * inv, gk3 = -(a1 * k1 + a2 * k2) * inv * inv;
            float g_s = gk1 * qq - gk2 * pp, g_t = -gk1 * n + gk2 * m;
            float g_m = gk2 * t + gk3 * qq, g_pp = -gk2 * s - gk3 * n, g_n = -gk1 * t - gk3 * pp, g_q = gk1 * s + gk3 * m;
            float2* gx2 = reinterpret_cast<float2*>(g_xy + ((size_t)b * F + f) * 6);
            atomicAdd(gx2 + 0, make_float2(-(g_m + g_n + g_s), -(g_pp + g_q + g_t)));
            atomicAdd(gx2 + 1, make_float2(g_m, g_pp));
            atomicAdd(gx2 + 2, make_float2(g_n, g_q));
        }
    }
}

// backward: one thread per (pixel, slot)
__global__ void __launch_bounds__(256) rd_backward_kernel(const float* __restrict__ pix, const float* __restrict__ face_xy,
                                                          const float* __restrict__ face_feat, const long long* __restrict__ idx,
                                                          const float* __restrict__ g_out, int P, int F, int D, int K, float eps,
                                                          float* __restrict__ g_xy, float* __restrict__ g_feat) {
    const int b = blockIdx.y;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)P * K) return;
    const size_t o = (size_t)b * P * K + e;
    long long f = idx[o];
    if (f < 0) return;
    const long long p = e / K;
    const float px = pix[((size_t)b * P + p) * 2], py = pix[((size_t)b * P + p) * 2 + 1];
    const float* xy = face_xy + ((size_t)b * F + f) * 6;
    float ax = xy[0], ay = xy[1], bx = xy[2], by = xy[3], cx = xy[4], cy = xy[5];
    float m = bx - ax, pp = by - ay, n = cx - ax, q = cy - ay, s = px - ax, t = py - ay;
    float k1 = s * q - n * t, k2 = m * t - s * pp, k3 = m * q - n * pp;
    float den = k3 + eps, inv = 1.f / den;
    float w1 = k1 * inv, w2 = k2 * inv, w0 = 1.f - w1 - w2;
    const float* go = g_out + o * D;
    const float* ff = face_feat + ((size_t)b * F + f) * 3 * D;
    float gw0 = 0.f, gw1 = 0.f, gw2 = 0.f;
    float* gfb = g_feat ? g_feat + ((size_t)b * F + f) * 3 * D : nullptr;
    if (D == 4 && ((((size_t)go) | ((size_t)ff) | ((size_t)gfb)) & 15) == 0) {       // RGBA: 16-byte vector loads / reductions
        float4 g4 = *reinterpret_cast<const float4*>(go);
        const float4* f4 = reinterpret_cast<const float4*>(ff);
        float4 a = f4[0], bq = f4[1], c = f4[2];
        gw0 = g4.x * a.x + g4.y * a.y + g4.z * a.z + g4.w * a.w;
        gw1 = g4.x * bq.x + g4.y * bq.y + g4.z * bq.z + g4.w * bq.w;
        gw2 = g4.x * c.x + g4.y * c.y + g4.z * c.z + g4.w * c.w;
        if (gfb) {
            float4* gf4 = reinterpret_cast<float4*>(gfb);
            atomicAdd(gf4 + 0, make_float4(w0 * g4.x, w0 * g4.y, w0 * g4.z, w0 * g4.w));
            atomicAdd(gf4 + 1, make_float4(w1 * g4.x, w1 * g4.y, w1 * g4.z, w1 * g4.w));
            atomicAdd(gf4 + 2, make_float4(w2 * g4.x, w2 * g4.y, w2 * g4.z, w2 * g4.w));
        }
    } else {
        for (int c = 0; c < D; ++c) {
            float gc = go[c];
            gw0 += gc * ff[c]; gw1 += gc * ff[D + c]; gw2 += gc * ff[2 * D + c];
            if (gfb) { atomicAdd(gfb + c, w0 * gc); atomicAdd(gfb + D + c, w1 * gc); atomicAdd(gfb + 2 * D + c, w2 * gc); }
        }
    }
    if (g_xy) {
        float a1 = gw1 - gw0, a2 = gw2 - gw0;                 // w0 = 1 - w1 - w2
        float gk1 = a1 * inv, gk2 = a2 * inv, gk3 = -(a1 * k1 + a2 * k2) * inv * inv;
        float g_s = gk1 * q - gk2 * pp, g_t = -gk1 * n + gk2 * m;
        float g_m = gk2 * t + gk3 * q, g_pp = -gk2 * s - gk3 * n, g_n = -gk1 * t - gk3 * pp, g_q = gk1 * s + gk3 * m;
        float* gx = g_xy + ((size_t)b * F + f) * 6;
        if ((((size_t)gx) & 7) == 0) {                        // one 8-byte vector reduction per vertex
            float2* gx2 = reinterpret_cast<float2*>(gx);
            atomicAdd(gx2 + 0, make_float2(-(g_m + g_n + g_s), -(g_pp + g_q + g_t)));
            atomicAdd(gx2 + 1, make_float2(g_m, g_pp));
            atomicAdd(gx2 + 2, make_float2(g_n, g_q));
        } else {
            atomicAdd(gx + 0, -(g_m + g_n + g_s)); atomicAdd(gx + 1, -(g_pp + g_q + g_t));
            atomicAdd(gx + 2, g_m); atomicAdd(gx + 3, g_pp); atomicAdd(gx + 4, g_n); atomicAdd(gx + 5, g_q);
        }
    }
}

}  // namespace dtb

using namespace dtb;

static int rd_kpad(int K) { int p = 32; while (p < K) p <<= 1; return p; }

extern "C" size_t dtb_sparse_render_workspace(int B, int P, int F, int R, long long pair_capacity) {
    (void)P;
    if (R <= 0) R = 64;
    if (pair_capacity <= 0) pair_capacity = (long long)B * F * 8;
    size_t cells = (size_t)B * R * R, n = (size_t)pair_capacity;
    return align_up((size_t)B * 16, 256) + 2 * align_up((size_t)B * F * 4, 256) + 2 * align_up(cells * 4, 256) + 2 * align_up(n * 8, 256) +
           2 * align_up(n * 4, 256) + scan_workspace_bytes((size_t)B * F) + scan_workspace_bytes(cells) + sort_workspace_bytes(n) + 2048;
}

struct RdBinning { unsigned* bbox; unsigned* cstart; unsigned* cend; unsigned* faces; int R; };

// carve the binning arrays out of the workspace (same layout for forward and for a backward that reuses them) and, when
// `build` is set, bin the faces: (cell, face) pairs sorted by cell then face id
static int rd_binning(const float* pixel_coords, const float* face_xy, int B, int P, int F, int R, long long pair_capacity, int32_t* overflow,
                      void* workspace, size_t workspace_bytes, bool build, cudaStream_t st, RdBinning& out) {
    if (R <= 0) R = 64;
    if (pair_capacity <= 0) pair_capacity = (long long)B * F * 8;
    size_t cells = (size_t)B * R * R, cap = (size_t)pair_capacity;
    Workspace ws(workspace, workspace_bytes);
    unsigned* bbox = ws.take<unsigned>((size_t)B * 4);
    unsigned* npairs = ws.take<unsigned>((size_t)B * F);
    unsigned* pair_off = ws.take<unsigned>((size_t)B * F);
    unsigned* cstart = ws.take<unsigned>(cells);
    unsigned* cend = ws.take<unsigned>(cells);
    u64* k0 = ws.take<u64>(cap); u64* k1 = ws.take<u64>(cap);
    unsigned* v0 = ws.take<unsigned>(cap); unsigned* v1 = ws.take<unsigned>(cap);
    size_t sb1 = scan_workspace_bytes((size_t)B * F), sb2 = scan_workspace_bytes(cells), sob = sort_workspace_bytes(cap);
    void* sws1 = ws.take<char>(sb1); void* sws2 = ws.take<char>(sb2); void* sows = ws.take<char>(sob);
    unsigned* total = ws.take<unsigned>(1);
    if (!ws.ok || !workspace) { set_error("sparse_render: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    out.bbox = bbox; out.cstart = cstart; out.cend = cend; out.faces = v1; out.R = R;
    if (!build) return DTB_OK;
    rd_init_kernel<<<cdiv(B * 4, 64), 64, 0, st>>>(bbox, B);
    dim3 gp(min(cdiv(P, 256), 128), B);
    rd_bbox_kernel<<<gp, 256, 0, st>>>(pixel_coords, P, bbox);
    DTB_LAUNCH_CHECK("rd_bbox");
    DTB_CUDA(cudaMemsetAsync(cstart, 0, cells * 4, st));
    DTB_CUDA(cudaMemsetAsync(total, 0, 4, st));
    if (F > 0) {
        dim3 gf(cdiv(F, 256), B);
        rd_count_kernel<<<gf, 256, 0, st>>>(face_xy, F, R, bbox, npairs);
        DTB_LAUNCH_CHECK("rd_count");
        int rc = exclusive_scan_u32(npairs, pair_off, (size_t)B * F, total, sws1, sb1, st);
        if (rc) return rc;
        rd_pairs_kernel<<<gf, 256, 0, st>>>(face_xy, F, R, bbox, pair_off, cap, k0, v0, cstart);
        DTB_LAUNCH_CHECK("rd_pairs");
        u64 maxkey = (u64)cells * (u64)F;
        int bits = 1; while (bits < 64 && (maxkey >> bits)) ++bits;
        rd_pad_kernel<<<cdiv((long long)cap, 256), 256, 0, st>>>(k0, v0, total, cap, maxkey);
        DTB_LAUNCH_CHECK("rd_pad");
        rc = radix_sort_pairs_u64(k0, v0, k1, v1, cap, bits, sows, sob, st);
        if (rc) return rc;
    }
    int rc = exclusive_scan_u32(cstart, cstart, cells, nullptr, sws2, sb2, st);
    if (rc) return rc;
    rd_cell_end_kernel<<<cdiv((long long)cells, 256), 256, 0, st>>>(cstart, total, cells, cap, cend, overflow);
    DTB_LAUNCH_CHECK("rd_cell_end");
    return DTB_OK;
}

// Number of (cell, face) pairs the binning will produce for these inputs: written to *n_pairs (device int64-free: u32).
// Lets the caller size pair_capacity exactly instead of guessing (one tiny kernel sequence + one host read).
__global__ void rd_sum_kernel(const unsigned* __restrict__ npairs, size_t n, unsigned* __restrict__ total) {
    unsigned s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += npairs[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}
extern "C" int dtb_sparse_render_pair_count(const float* pixel_coords, const float* face_xy, int B, int P, int F, int R, unsigned* n_pairs,
                                            void* workspace, size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(n_pairs != nullptr, "sparse_render_pair_count: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    DTB_CUDA(cudaMemsetAsync(n_pairs, 0, sizeof(unsigned), st));
    if (B == 0 || P == 0 || F == 0) return DTB_OK;
    DTB_REQUIRE(pixel_coords && face_xy, "sparse_render_pair_count: null argument");
    if (R <= 0) R = 64;
    Workspace ws(workspace, workspace_bytes);
    unsigned* bbox = ws.take<unsigned>((size_t)B * 4);
    unsigned* npairs = ws.take<unsigned>((size_t)B * F);
    if (!ws.ok || !workspace) { set_error("sparse_render_pair_count: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    rd_init_kernel<<<cdiv(B * 4, 64), 64, 0, st>>>(bbox, B);
    dim3 gp(min(cdiv(P, 256), 128), B);
    rd_bbox_kernel<<<gp, 256, 0, st>>>(pixel_coords, P, bbox);
    dim3 gf(cdiv(F, 256), B);
    rd_count_kernel<<<gf, 256, 0, st>>>(face_xy, F, R, bbox, npairs);
    rd_sum_kernel<<<min(cdiv((long long)B * F, 256), 256), 256, 0, st>>>(npairs, (size_t)B * F, n_pairs);
    DTB_LAUNCH_CHECK("rd_pair_count");
    return DTB_OK;
}

// pixel_coords (B,P,2), render_ranges (B,P,2), face_z (B,F,3), face_xy (B,F,3,2), face_feat (B,F,3,D) ->
// out_feat (B,P,K,D) f32, out_idx (B,P,K) i64.  R: cells per axis of the face-binning grid (<=0: 64);
// pair_capacity: room for (cell, face) pairs (<=0: 8 per face); *overflow (device int) is set if it was too small.
extern "C" int dtb_sparse_render_forward(const float* pixel_coords, const float* render_ranges, const float* face_z, const float* face_xy,
                                         const float* face_feat, int B, int P, int F, int D, int K, float eps, int R, long long pair_capacity,
                                         float* out_feat, long long* out_idx, int32_t* overflow, void* workspace, size_t workspace_bytes,
                                         void* stream) {
    if (B == 0 || P == 0 || K == 0) return DTB_OK;
    DTB_REQUIRE(pixel_coords && render_ranges && out_feat && out_idx && overflow && D > 0, "sparse_render_forward: bad argument");
    DTB_REQUIRE(K <= 1024, "sparse_render_forward: K=%d > 1024 not supported", K);
    DTB_REQUIRE(F == 0 || (face_z && face_xy && face_feat), "sparse_render_forward: null faces");
    cudaStream_t st = (cudaStream_t)stream;
    RdBinning bn;
    int rc = rd_binning(pixel_coords, face_xy, B, P, F, R, pair_capacity, overflow, workspace, workspace_bytes, true, st, bn);
    if (rc) return rc;
    int Kpad = rd_kpad(K);
    size_t smem = (size_t)RD_WARPS * Kpad * 16;
    if (smem > 48 * 1024) DTB_CUDA(cudaFuncSetAttribute(rd_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(P, RD_WARPS), B);
    rd_forward_kernel<<<grid, RD_WARPS * 32, smem, st>>>(pixel_coords, render_ranges, face_z, face_xy, face_feat, P, F, D, K, Kpad, eps, bn.R,
                                                         bn.bbox, bn.cstart, bn.cend, bn.faces, out_feat, out_idx);
    DTB_LAUNCH_CHECK("rd_forward");
    return DTB_OK;
}

// Fused render + front-to-back composite: out_color (B,P,D-1), out_mask (B,P,1); the workspace keeps the face binning for
// dtb_render_composite_backward (call it with the SAME workspace, sizes, R and pair_capacity).
extern "C" int dtb_render_composite_forward(const float* pixel_coords, const float* render_ranges, const float* face_z, const float* face_xy,
                                            const float* face_feat, int B, int P, int F, int D, int K, float eps, int R, long long pair_capacity,
                                            float* out_color, float* out_mask, int32_t* overflow, void* workspace, size_t workspace_bytes,
                                            void* stream) {
    if (B == 0 || P == 0) return DTB_OK;
    DTB_REQUIRE(pixel_coords && render_ranges && out_color && out_mask && overflow, "render_composite_forward: null argument");
    DTB_REQUIRE(D >= 2 && D <= RC_MAXD, "render_composite_forward: D=%d must be in [2, %d] (alpha + colour channels)", D, RC_MAXD);
    DTB_REQUIRE(K >= 1 && K <= 1024, "render_composite_forward: K=%d out of range", K);
    DTB_REQUIRE(F == 0 || (face_z && face_xy && face_feat), "render_composite_forward: null faces");
    cudaStream_t st = (cudaStream_t)stream;
    RdBinning bn;
    int rc = rd_binning(pixel_coords, face_xy, B, P, F, R, pair_capacity, overflow, workspace, workspace_bytes, true, st, bn);
    if (rc) return rc;
    int Kpad = rd_kpad(K);
    size_t smem = (size_t)RD_WARPS * Kpad * 16;
    if (smem > 48 * 1024) DTB_CUDA(cudaFuncSetAttribute(rc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(P, RD_WARPS), B);
    rc_forward_kernel<<<grid, RD_WARPS * 32, smem, st>>>(pixel_coords, render_ranges, face_z, face_xy, face_feat, P, F, D, K, Kpad, eps, bn.R,
                                                         bn.bbox, bn.cstart, bn.cend, bn.faces, out_color, out_mask);
    DTB_LAUNCH_CHECK("rc_forward");
    return DTB_OK;
}

// g_color (B,P,D-1), g_mask (B,P,1) -> ACCUMULATES into g_xy (B,F,3,2) and g_feat (B,F,3,D) (either may be NULL)
extern "C" int dtb_render_composite_backward(const float* pixel_coords, const float* render_ranges, const float* face_z, const float* face_xy,
                                             const float* face_feat, const float* g_color, const float* g_mask, int B, int P, int F, int D, int K,
                                             float eps, int R, long long pair_capacity, float* g_xy, float* g_feat, void* workspace,
                                             size_t workspace_bytes, void* stream) {
    if (B == 0 || P == 0 || F == 0) return DTB_OK;
    DTB_REQUIRE(pixel_coords && render_ranges && face_z && face_xy && face_feat && g_color && g_mask, "render_composite_backward: null argument");
    DTB_REQUIRE(D >= 2 && D <= RC_MAXD && K >= 1 && K <= 1024, "render_composite_backward: D / K out of range");
    cudaStream_t st = (cudaStream_t)stream;
    RdBinning bn;
    int rc = rd_binning(pixel_coords, face_xy, B, P, F, R, pair_capacity, nullptr, workspace, workspace_bytes, false, st, bn);
    if (rc) return rc;
    int Kpad = rd_kpad(K);
    size_t smem = (size_t)RD_WARPS * Kpad * 24;
    if (smem > 48 * 1024) DTB_CUDA(cudaFuncSetAttribute(rc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(P, RD_WARPS), B);
    rc_backward_kernel<<<grid, RD_WARPS * 32, smem, st>>>(pixel_coords, render_ranges, face_z, face_xy, face_feat, P, F, D, K, Kpad, eps, bn.R,
                                                          bn.bbox, bn.cstart, bn.cend, bn.faces, g_color, g_mask, g_xy, g_feat);
    DTB_LAUNCH_CHECK("rc_backward");
    return DTB_OK;
}

// g_out (B,P,K,D), idx (B,P,K) from forward -> ACCUMULATES into g_xy (B,F,3,2) and g_feat (B,F,3,D) (either may be NULL)
extern "C" int dtb_sparse_render_backward(const float* pixel_coords, const float* face_xy, const float* face_feat, const long long* idx,
                                          const float* g_out, int B, int P, int F, int D, int K, float eps, float* g_xy, float* g_feat,
                                          void* stream) {
    if (B == 0 || P == 0 || K == 0 || F == 0) return DTB_OK;
    DTB_REQUIRE(pixel_coords && face_xy && face_feat && idx && g_out, "sparse_render_backward: null argument");
    dim3 grid(cdiv((long long)P * K, 256), B);
    rd_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pixel_coords, face_xy, face_feat, idx, g_out, P, F, D, K, eps, g_xy, g_feat);
    DTB_LAUNCH_CHECK("rd_backward");
    return DTB_OK;
}
