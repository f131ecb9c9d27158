// Uniform-grid binning of a batched point set (counting sort by cell), shared by the point-in-tet query
// (points binned, tets traverse cells), the 1-NN search and the point->face distance search.
//
// Layout (all in the caller's workspace), per batch sample b and cubic grid of G^3 cells:
//   bbox_ord [B][6]            order-preserving-encoded min xyz / max xyz of the items
//   cell_start/cell_end [B*G^3] item range of each cell inside `sorted`
//   sorted  [B*N] float4       (x, y, z, original index as int bits), grouped by cell
//   mask    [B*(G/4)^3] u64    brick layout only: bit (z&3)*16+(y&3)*4+(x&3) of the brick's word = cell occupied
// Cell of x on one axis: clamp(floor((x - min) * inv_h), 0, G-1), inv_h = G / (max extent * (1 + 2^-20)):
// monotone in x, so conservative cell ranges of boxes are obtained by mapping their corners.
#pragma once
#include "prims.cuh"

namespace dtb {

struct GridParams {     // per-sample, computed on device from bbox_ord
    float ox, oy, oz;   // origin (bbox min)
    float inv_h;        // cells per unit length
    float h;            // cell edge
};

struct PointGrid {
    int B, N, G;
    int xmult;                // row-major layout only: cells along x are xmult times finer (G * xmult per row); 1 = cubic cells
    bool brick;               // cell order: false = row-major (z,y,x); true = 4x4x4 bricks (G % 4 == 0), 64 cells per brick
    unsigned* bbox_ord;       // [B][6]
    unsigned* cell_start;     // [B*G^3]
    unsigned* cell_end;       // [B*G^3]
    float4* sorted;           // [B*N]
    unsigned* cell_of;        // [B*N] scratch (cell id per item)
    unsigned long long* mask; // [B*G*G*W] or nullptr
    void* scan_ws; size_t scan_ws_bytes;
    size_t clear_bytes;       // bbox_ord .. end of cell_start (zeroed by one memset)
};

__device__ __forceinline__ GridParams grid_params(const unsigned* bbox_ord, int b, int G) {
    // bbox_ord is zero-initialised (one memset together with the cell counters): slots 0-2 hold ~ord(min) so that
    // atomicMax works for the minimum too; an untouched box (no finite item) reads as the unit box at the origin
    const unsigned* q = bbox_ord + (size_t)b * 6;
    GridParams g;
    if (q[3] == 0u || q[0] == 0u) { g.ox = g.oy = g.oz = 0.f; g.inv_h = (float)G; g.h = 1.0f / (float)G; return g; }
    g.ox = ord2f(~q[0]); g.oy = ord2f(~q[1]); g.oz = ord2f(~q[2]);
    float ex = ord2f(q[3]) - g.ox, ey = ord2f(q[4]) - g.oy, ez = ord2f(q[5]) - g.oz;
    float ext = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-20f));
    ext = ext * (1.0f + 9.5367431640625e-7f);
    g.inv_h = (float)G / ext;
    g.h = ext / (float)G;
    return g;
}
// linear cell id inside one sample
__device__ __forceinline__ unsigned cell_index(int cx, int cy, int cz, int G, bool brick) {
    if (!brick) return ((unsigned)cz * G + cy) * G + cx;
    unsigned nb = (unsigned)G >> 2;
    unsigned bid = (((unsigned)cz >> 2) * nb + ((unsigned)cy >> 2)) * nb + ((unsigned)cx >> 2);
    return bid * 64u + ((cz & 3) << 4) + ((cy & 3) << 2) + (cx & 3);
}
__device__ __forceinline__ int cell_coord(float x, float o, float inv_h, int G) {
    float f = floorf((x - o) * inv_h);
    f = fminf(fmaxf(f, 0.0f), (float)(G - 1));      // NaN -> 0
    return (int)f;
}

size_t pointgrid_workspace_bytes(int B, int N, int G, bool with_mask, bool brick, int xmult = 1);
// carve a PointGrid out of ws (returns false if it does not fit)
bool pointgrid_carve(PointGrid& pg, int B, int N, int G, bool with_mask, bool brick, Workspace& ws, int xmult = 1);
// enqueue: bbox -> count -> scan -> fill (-> mask).  items: (B,N,3) f32 contiguous, or, when
// `tri_centroid` is true, (B,N,3,3) triangles binned by centroid.
int pointgrid_build(PointGrid& pg, const float* items, bool tri_centroid, cudaStream_t st);
// same, but only the first counts[b] items of sample b exist (padded-ragged batches)
int pointgrid_build_ragged(PointGrid& pg, const float* items, bool tri_centroid, const int32_t* counts, cudaStream_t st);

}  // namespace dtb
