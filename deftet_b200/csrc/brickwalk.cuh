// Exact nearest-item search over a brick-layout PointGrid: walk cubic shells of 4x4x4-cell bricks outward
// from the query, skip empty bricks with one 64-bit occupancy word, prune bricks and cells with a
// conservative lower bound, and hand every surviving item to the visitor.  The visitor keeps the running
// lexicographic minimum (value, index), so the result equals a brute-force scan in index order with a
// strict '<' -- the tie rule of every reference search kernel (SURVEY.md section 7 "hard parts").
#pragma once
#include "pointgrid.cuh"

namespace dtb {

// squared distance from q to the axis-aligned box [lo, lo + w]^3-ish (per-axis lo, common width w)
__device__ __forceinline__ float box_dist2(float qx, float qy, float qz, float lx, float ly, float lz, float w, float shrink) {
    float dx = fmaxf(fmaxf(lx - qx, qx - (lx + w)), 0.f);
    float dy = fmaxf(fmaxf(ly - qy, qy - (ly + w)), 0.f);
    float dz = fmaxf(fmaxf(lz - qz, qz - (lz + w)), 0.f);
    float d = sqrtf(dx * dx + dy * dy + dz * dz);
    d = fmaxf(d - shrink, 0.f) * 0.9999f;
    return d * d;
}

// occupancy-word masks of the cells of a 4x4x4 brick whose x / y / z index lies in [lo, hi] (0 <= lo <= hi <= 3)
__device__ __forceinline__ unsigned long long brick_xmask(int lo, int hi) {
    return (unsigned long long)(((1u << (hi - lo + 1)) - 1u) << lo) * 0x1111111111111111ull;
}
__device__ __forceinline__ unsigned long long brick_ymask(int lo, int hi) {
    unsigned long long m16 = ((1ull << (4 * (hi - lo + 1))) - 1ull) << (4 * lo);
    return m16 * 0x0001000100010001ull;
}
__device__ __forceinline__ unsigned long long brick_zmask(int lo, int hi) {
    int n = 16 * (hi - lo + 1);
    unsigned long long m = (n >= 64) ? ~0ull : ((1ull << n) - 1ull);
    return m << (16 * lo);
}
// cells of the brick at (lx,ly,lz) that the ball (q, r) can reach; 0 if none
__device__ __forceinline__ unsigned long long brick_reach_mask(float qx, float qy, float qz, float r, float lx, float ly, float lz, float inv_h) {
    float fx0 = floorf((qx - r - lx) * inv_h), fx1 = floorf((qx + r - lx) * inv_h);
    float fy0 = floorf((qy - r - ly) * inv_h), fy1 = floorf((qy + r - ly) * inv_h);
    float fz0 = floorf((qz - r - lz) * inv_h), fz1 = floorf((qz + r - lz) * inv_h);
    if (fx1 < 0.f || fx0 > 3.f || fy1 < 0.f || fy0 > 3.f || fz1 < 0.f || fz0 > 3.f) return 0ull;
    int x0 = (int)fmaxf(fx0, 0.f), x1 = (int)fminf(fx1, 3.f), y0 = (int)fmaxf(fy0, 0.f), y1 = (int)fminf(fy1, 3.f);
    int z0 = (int)fmaxf(fz0, 0.f), z1 = (int)fminf(fz1, 3.f);
    return brick_xmask(x0, x1) & brick_ymask(y0, y1) & brick_zmask(z0, z1);
}

// V must provide:  float bound() const   -- current best value (squared distance) for pruning
//                  static float no_hit()  -- the value bound() has while nothing has been accepted (initial best)
//                  void item(const float4& it)  -- evaluate one item (x,y,z = binned position, w = index bits)
// `inflate`: radius by which an item may extend beyond the position it was binned with (0 for points).
// visit one brick: skip if empty / out of reach, else scan the cells the current search ball can reach
template <typename V>
__device__ __forceinline__ void brick_visit(float qx, float qy, float qz, const GridParams& g, int NB, int bx, int by, int bz, float bw,
                                            float shrink, unsigned long long skip_bits, const unsigned* __restrict__ cell_start,
                                            const unsigned* __restrict__ cell_end, const float4* __restrict__ sorted,
                                            const unsigned long long* __restrict__ mask, size_t cell_base, V& vis) {
    const size_t brick = ((size_t)bz * NB + by) * NB + bx;
    unsigned long long m = __ldg(mask + (cell_base >> 6) + brick) & ~skip_bits;
    if (!m) return;
    const float lx = g.ox + (float)bx * bw, ly = g.oy + (float)by * bw, lz = g.oz + (float)bz * bw;
    if (box_dist2(qx, qy, qz, lx, ly, lz, bw, shrink) > vis.bound()) return;
    // keep only the cells inside the bounding cube of the current search ball (radius sqrt(best) + shrink, padded by
    // 1 % of a cell against the rounding of the cell assignment)
    m &= brick_reach_mask(qx, qy, qz, sqrtf(vis.bound()) * 1.0001f + shrink + 0.01f * g.h, lx, ly, lz, g.inv_h);
    const size_t c0 = cell_base + brick * 64;
    while (m) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        const float cxl = lx + (float)(k & 3) * g.h, cyl = ly + (float)((k >> 2) & 3) * g.h, czl = lz + (float)(k >> 4) * g.h;
        if (box_dist2(qx, qy, qz, cxl, cyl, czl, g.h, shrink) > vis.bound()) continue;
        const unsigned j0 = __ldg(cell_start + c0 + k), j1 = __ldg(cell_end + c0 + k);
        for (unsigned j = j0; j < j1; ++j) vis.item(__ldg(sorted + j));
    }
}

template <typename V>
__device__ __forceinline__ void brick_walk(float qx, float qy, float qz, const GridParams& g, int G, float inflate,
                                           const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_end,
                                           const float4* __restrict__ sorted, const unsigned long long* __restrict__ mask,
                                           size_t cell_base, V& vis) {
    const int NB = G >> 2;
    const float bw = 4.0f * g.h;
    // rounding slack of the cell assignment floor((x-o)*inv_h): a few ulps of the coordinates, far below 1e-3 h
    const float slack = 1e-3f * g.h + 1e-6f * (fabsf(g.ox) + fabsf(g.oy) + fabsf(g.oz) + (float)G * g.h);
    const float shrink = inflate + slack;
    const int cx = cell_coord(qx, g.ox, g.inv_h, G), cy = cell_coord(qy, g.oy, g.inv_h, G), cz = cell_coord(qz, g.oz, g.inv_h, G);
    const int bx0 = cx >> 2, by0 = cy >> 2, bz0 = cz >> 2;
    const unsigned home_cell = cell_index(cx, cy, cz, G, true);
    {   // the query's own cell first, so that the bound is finite before anything is filtered
        const size_t hc = cell_base + home_cell;
        const unsigned j0 = __ldg(cell_start + hc), j1 = __ldg(cell_end + hc);
        for (unsigned j = j0; j < j1; ++j) vis.item(__ldg(sorted + j));
    }
    // phase 1: brick shells outward until something has been found (shell R = bricks at Chebyshev distance R)
    const float no_hit = V::no_hit();                       // the visitor's bound while nothing has been accepted yet
    int Rv = -1;                                            // shells 0..Rv have been visited
    for (int R = 0; R < NB; ++R) {
        if (R >= 1 && vis.bound() < no_hit) break;          // a candidate exists: switch to the bounded enumeration below
        const int z0 = max(bz0 - R, 0), z1 = min(bz0 + R, NB - 1), y0 = max(by0 - R, 0), y1 = min(by0 + R, NB - 1);
        for (int bz = z0; bz <= z1; ++bz) {
            const bool zface = (bz == bz0 - R) || (bz == bz0 + R);
            for (int by = y0; by <= y1; ++by) {
                const bool full = zface || (by == by0 - R) || (by == by0 + R);
                const int xa = max(bx0 - R, 0), xb = min(bx0 + R, NB - 1);
                const int step = full ? 1 : max(2 * R, 1);
                for (int bx = bx0 - R; bx <= bx0 + R; bx += step) {
                    if (bx < xa || bx > xb) continue;
                    brick_visit(qx, qy, qz, g, NB, bx, by, bz, bw, shrink, R == 0 ? (1ull << (home_cell & 63u)) : 0ull, cell_start, cell_end,
                                sorted, mask, cell_base, vis);
                }
            }
        }
        Rv = R;
    }
    if (!(vis.bound() < no_hit)) return;                    // the grid holds nothing the visitor accepts
    // phase 2: only the bricks that the bounding cube of the search ball touches and that phase 1 has not visited.
    // The ball only shrinks from here on, so the range computed now is a superset of what is needed.
    const float r = sqrtf(vis.bound()) * 1.0001f + shrink + 0.01f * g.h;
    const float ibw = g.inv_h * 0.25f;
    const int xlo = max((int)floorf((qx - r - g.ox) * ibw), 0), xhi = min((int)floorf((qx + r - g.ox) * ibw), NB - 1);
    const int ylo = max((int)floorf((qy - r - g.oy) * ibw), 0), yhi = min((int)floorf((qy + r - g.oy) * ibw), NB - 1);
    const int zlo = max((int)floorf((qz - r - g.oz) * ibw), 0), zhi = min((int)floorf((qz + r - g.oz) * ibw), NB - 1);
    for (int bz = zlo; bz <= zhi; ++bz)
        for (int by = ylo; by <= yhi; ++by)
            for (int bx = xlo; bx <= xhi; ++bx) {
                if (max(max(abs(bx - bx0), abs(by - by0)), abs(bz - bz0)) <= Rv) continue;
                brick_visit(qx, qy, qz, g, NB, bx, by, bz, bw, shrink, 0ull, cell_start, cell_end, sorted, mask, cell_base, vis);
            }
}

}  // namespace dtb
