#include "pointgrid.cuh"

namespace dtb {

__device__ __forceinline__ void load_item(const float* items, size_t i, bool tri, float& x, float& y, float& z) {
    if (!tri) {
        x = items[i * 3]; y = items[i * 3 + 1]; z = items[i * 3 + 2];
    } else {
        const float* f = items + i * 9;
        x = (f[0] + f[3] + f[6]) * (1.0f / 3.0f);
        y = (f[1] + f[4] + f[7]) * (1.0f / 3.0f);
        z = (f[2] + f[5] + f[8]) * (1.0f / 3.0f);
    }
}

__global__ void __launch_bounds__(256) pg_bbox_kernel(const float* __restrict__ items, bool tri, int N, const int32_t* __restrict__ counts, unsigned* __restrict__ bbox_ord) {
    int b = blockIdx.y;
    const int n = counts ? min(counts[b], N) : N;
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float x, y, z;
        load_item(items, (size_t)b * N + i, tri, x, y, z);
        if (fabsf(x) <= 3.0e38f && fabsf(y) <= 3.0e38f && fabsf(z) <= 3.0e38f) {       // finite items only: one inf would blow the grid up to a single cell
            mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
            mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
    __shared__ float s_mn[8][3], s_mx[8][3];
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { s_mn[threadIdx.x >> 5][k] = mn[k]; s_mx[threadIdx.x >> 5][k] = mx[k]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        int k = threadIdx.x;
        float a = s_mn[0][k], c = s_mx[0][k];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a = fminf(a, s_mn[w][k]); c = fmaxf(c, s_mx[w][k]); }
        if (a <= c) {
            atomicMax(&bbox_ord[(size_t)b * 6 + k], ~f2ord(a));
            atomicMax(&bbox_ord[(size_t)b * 6 + 3 + k], f2ord(c));
        }
    }
}

__global__ void __launch_bounds__(256) pg_count_kernel(const float* __restrict__ items, bool tri, int N, int G, int xmult, bool brick,
                                                       const int32_t* __restrict__ counts, const unsigned* __restrict__ bbox_ord, unsigned* __restrict__ cell_count,
                                                       unsigned* __restrict__ cell_of) {
    int b = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || (counts && i >= counts[b])) return;
    GridParams g = grid_params(bbox_ord, b, G);
    float x, y, z;
    load_item(items, (size_t)b * N + i, tri, x, y, z);
    int cy = cell_coord(y, g.oy, g.inv_h, G), cz = cell_coord(z, g.oz, g.inv_h, G);
    unsigned c;
    if (xmult > 1) {          // row-major, x refined: (z, y) rows of G * xmult cells
        const int Gx = G * xmult;
        c = ((unsigned)b * G * G + (unsigned)cz * G + cy) * Gx + cell_coord(x, g.ox, g.inv_h * (float)xmult, Gx);
    } else {
        c = (unsigned)b * G * G * G + cell_index(cell_coord(x, g.ox, g.inv_h, G), cy, cz, G, brick);
    }
    cell_of[(size_t)b * N + i] = c;
    atomicAdd(&cell_count[c], 1u);
}

__global__ void __launch_bounds__(256) pg_fill_kernel(const float* __restrict__ items, bool tri, int N, const int32_t* __restrict__ counts,
                                                      const unsigned* __restrict__ cell_of, unsigned* __restrict__ cell_end,
                                                      float4* __restrict__ sorted) {
    int b = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || (counts && i >= counts[b])) return;
    float x, y, z;
    load_item(items, (size_t)b * N + i, tri, x, y, z);
    unsigned c = cell_of[(size_t)b * N + i];
    unsigned dst = atomicAdd(&cell_end[c], 1u);
    sorted[dst] = make_float4(x, y, z, __int_as_float(i));
}

__global__ void pg_brick_mask_kernel(const unsigned* __restrict__ cell_start, const unsigned* __restrict__ cell_end, size_t n_bricks,
                                     unsigned long long* __restrict__ mask) {
    size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_bricks) return;
    unsigned long long m = 0;
    const uint4* s4 = reinterpret_cast<const uint4*>(cell_start + w * 64);
    const uint4* e4 = reinterpret_cast<const uint4*>(cell_end + w * 64);
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        uint4 s = s4[k], e = e4[k];
        if (e.x > s.x) m |= 1ull << (4 * k);
        if (e.y > s.y) m |= 1ull << (4 * k + 1);
        if (e.z > s.z) m |= 1ull << (4 * k + 2);
        if (e.w > s.w) m |= 1ull << (4 * k + 3);
    }
    mask[w] = m;
}

size_t pointgrid_workspace_bytes(int B, int N, int G, bool with_mask, bool brick, int xmult) {
    Workspace ws(nullptr, 0);
    PointGrid pg;
    pointgrid_carve(pg, B, N, G, with_mask, brick, ws, xmult);
    return ws.off;
}

bool pointgrid_carve(PointGrid& pg, int B, int N, int G, bool with_mask, bool brick, Workspace& ws, int xmult) {
    pg.B = B; pg.N = N; pg.G = G; pg.brick = brick;
    pg.xmult = (brick || xmult < 1) ? 1 : xmult;
    size_t cells = (size_t)B * G * G * G * pg.xmult;
    // bbox_ord and cell_start are adjacent (bbox padded to 256 B) so that one memset clears both
    pg.bbox_ord = ws.take<unsigned>((size_t)B * 6);
    pg.cell_start = ws.take<unsigned>(cells);
    pg.clear_bytes = ws.base ? (size_t)((char*)(pg.cell_start + cells) - (char*)pg.bbox_ord) : 0;
    pg.cell_end = ws.take<unsigned>(cells);
    pg.sorted = ws.take<float4>((size_t)B * N);
    pg.cell_of = ws.take<unsigned>((size_t)B * N);
    pg.mask = (with_mask && brick) ? ws.take<unsigned long long>(cells / 64) : nullptr;
    pg.scan_ws_bytes = scan_workspace_bytes(cells);
    pg.scan_ws = ws.take<char>(pg.scan_ws_bytes);
    return ws.ok;
}

int pointgrid_build(PointGrid& pg, const float* items, bool tri, cudaStream_t st) { return pointgrid_build_ragged(pg, items, tri, nullptr, st); }

int pointgrid_build_ragged(PointGrid& pg, const float* items, bool tri, const int32_t* counts, cudaStream_t st) {
    const int B = pg.B, N = pg.N, G = pg.G;
    size_t cells = (size_t)B * G * G * G * pg.xmult;
    if (cells >= (1ull << 31) || (size_t)B * N >= (1ull << 31)) { set_error("pointgrid: problem too large for 32-bit cell ids"); return DTB_EOVERFLOW; }
    DTB_CUDA(cudaMemsetAsync(pg.bbox_ord, 0, pg.clear_bytes, st));
    if (N > 0) {
        dim3 gb(min(cdiv(N, 256 * 8), 64), B);
        pg_bbox_kernel<<<gb, 256, 0, st>>>(items, tri, N, counts, pg.bbox_ord);
        DTB_LAUNCH_CHECK("pg_bbox");
    }
    if (N > 0) {
        dim3 gc(cdiv(N, 256), B);
        pg_count_kernel<<<gc, 256, 0, st>>>(items, tri, N, G, pg.xmult, pg.brick, counts, pg.bbox_ord, pg.cell_start, pg.cell_of);
        DTB_LAUNCH_CHECK("pg_count");
    }
    int rc = exclusive_scan_u32_dup(pg.cell_start, pg.cell_start, pg.cell_end, cells, nullptr, pg.scan_ws, pg.scan_ws_bytes, st);
    if (rc) return rc;
    if (N > 0) {
        dim3 gc(cdiv(N, 256), B);
        pg_fill_kernel<<<gc, 256, 0, st>>>(items, tri, N, counts, pg.cell_of, pg.cell_end, pg.sorted);
        DTB_LAUNCH_CHECK("pg_fill");
    }
    if (pg.brick && (G % 4) != 0) { set_error("pointgrid: brick layout needs G %% 4 == 0 (G=%d)", G); return DTB_EINVAL; }
    if (pg.mask && pg.brick) {
        size_t n_bricks = cells / 64;
        pg_brick_mask_kernel<<<cdiv((long long)n_bricks, 128), 128, 0, st>>>(pg.cell_start, pg.cell_end, n_bricks, pg.mask);
        DTB_LAUNCH_CHECK("pg_brick_mask");
    } else if (pg.mask) {
        set_error("pointgrid: occupancy masks exist for the brick layout only");
        return DTB_EINVAL;
    }
    return DTB_OK;
}

}  // namespace dtb
