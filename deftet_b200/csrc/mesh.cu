// A16: inside/outside labels of points w.r.t. a watertight triangle mesh (stand-in for kal.ops.mesh.check_sign,
//      called at layers/DefTet/deftet.py:46 on the tet centroids; Kaolin is third-party, un-vendored and un-pinned,
//      so parity is defined by the exact ray-parity semantics restated in oracle/render_oracle.c -- "parity unpinned").
// A17: Laplacian smoothness of vertex offsets, sum ||(D^-1 A) d - d||^2 (layers/DefTet/deftet.py:340-343 through
//      utils/matrix_utils.py:22-33 sparse_batch_matmul), on the sorted edge list of the A10 builder.
#include "prims.cuh"
#include "deftet_b200.h"

namespace dtb {

// ---- A16 ----------------------------------------------------------------------------------------------
// +z ray parity.  Triangles are binned by the xy bounding box into an R x R grid over the mesh's xy extent; a
// point tests only the triangles of its cell.  The xy containment test is evaluated in double precision with a
// half-open edge rule (an edge belongs to the triangle on its left/top), so a ray through a shared edge or
// vertex of a consistently tessellated surface is counted exactly once.
struct XYGrid { double ox, oy, inv; int R; };

__device__ __forceinline__ XYGrid xy_grid(const unsigned* bbox_ord, int b, int R) {
    const unsigned* q = bbox_ord + (size_t)b * 6;
    XYGrid g;
    double x0 = ord2f(q[0]), y0 = ord2f(q[1]), x1 = ord2f(q[3]), y1 = ord2f(q[4]);
    double ext = fmax(fmax(x1 - x0, y1 - y0), 1e-30) * (1.0 + 1e-9);
    g.ox = x0; g.oy = y0; g.inv = (double)R / ext; g.R = R;
    return g;
}
__device__ __forceinline__ int xy_cell(double x, double o, double inv, int R) {
    double f = floor((x - o) * inv);
    if (!(f > 0.0)) f = 0.0;
    if (f > (double)(R - 1)) f = (double)(R - 1);
    return (int)f;
}

__global__ void __launch_bounds__(256) cs_bbox_kernel(const float* __restrict__ verts, int n, unsigned* __restrict__ bbox_ord) {
    int b = blockIdx.y;
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* v = verts + ((size_t)b * n + i) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], v[k]); mx[k] = fmaxf(mx[k], v[k]); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (mn[k] <= mx[k]) { atomicMin(&bbox_ord[(size_t)b * 6 + k], f2ord(mn[k])); atomicMax(&bbox_ord[(size_t)b * 6 + 3 + k], f2ord(mx[k])); }
        }
    }
}
__global__ void cs_init_kernel(unsigned* bbox_ord, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 6) bbox_ord[i] = ((i % 6) < 3) ? 0xffffffffu : 0u;
}

// pass 0: count (cell, tri) pairs per cell; pass 1: fill
__global__ void __launch_bounds__(256) cs_bin_kernel(const float* __restrict__ verts, int n, const int32_t* __restrict__ faces, int m, int R,
                                                     const unsigned* __restrict__ bbox_ord, unsigned* __restrict__ cell_cnt,
                                                     unsigned* __restrict__ cell_cur, int32_t* __restrict__ cell_list, int pass,
                                                     const unsigned* __restrict__ total = nullptr, size_t cap = 0) {
    int b = blockIdx.y;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= m) return;
    if (pass == 1 && total && (size_t)*total > cap) return;          // list would overflow: the query falls back to all faces
    XYGrid g = xy_grid(bbox_ord, b, R);
    const float* vb = verts + (size_t)b * n * 3;
    int i0 = faces[f * 3], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
    double x0 = vb[(size_t)i0 * 3], y0 = vb[(size_t)i0 * 3 + 1], x1 = vb[(size_t)i1 * 3], y1 = vb[(size_t)i1 * 3 + 1];
    double x2 = vb[(size_t)i2 * 3], y2 = vb[(size_t)i2 * 3 + 1];
    int cx0 = xy_cell(fmin(x0, fmin(x1, x2)), g.ox, g.inv, R), cx1 = xy_cell(fmax(x0, fmax(x1, x2)), g.ox, g.inv, R);
    int cy0 = xy_cell(fmin(y0, fmin(y1, y2)), g.oy, g.inv, R), cy1 = xy_cell(fmax(y0, fmax(y1, y2)), g.oy, g.inv, R);
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
            size_t c = ((size_t)b * R + cy) * R + cx;
            if (pass == 0) atomicAdd(cell_cnt + c, 1u);
            else cell_list[atomicAdd(cell_cur + c, 1u)] = f;
        }
}

// orientation of (a->b) w.r.t. p, with the half-open tie rule
__device__ __forceinline__ bool edge_inside(double ax, double ay, double bx, double by, double px, double py, bool ccw) {
    double e = (bx - ax) * (py - ay) - (by - ay) * (px - ax);
    if (!ccw) e = -e;
    if (e > 0.0) return true;
    if (e < 0.0) return false;
    // on the edge: owned if the (ccw-oriented) edge is a "left" edge (going down) or a horizontal "top" edge (going left)
    double dx = bx - ax, dy = by - ay;
    if (!ccw) { dx = -dx; dy = -dy; }
    return (dy < 0.0) || (dy == 0.0 && dx < 0.0);
}

__global__ void __launch_bounds__(256) cs_query_kernel(const float* __restrict__ verts, int n, const int32_t* __restrict__ faces, int R,
                                                       const unsigned* __restrict__ bbox_ord, const unsigned* __restrict__ cell_start,
                                                       const unsigned* __restrict__ cell_end, const int32_t* __restrict__ cell_list,
                                                       const float* __restrict__ points, int p, unsigned char* __restrict__ out,
                                                       int m = 0, const unsigned* __restrict__ total = nullptr, size_t cap = 0) {
    int b = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p) return;
    const float* q = points + ((size_t)b * p + i) * 3;
    double px = q[0], py = q[1], pz = q[2];
    const unsigned* bb = bbox_ord + (size_t)b * 6;
    bool inside = false;
    if (px >= (double)ord2f(bb[0]) && px <= (double)ord2f(bb[3]) && py >= (double)ord2f(bb[1]) && py <= (double)ord2f(bb[4])) {
        XYGrid g = xy_grid(bbox_ord, b, R);
        size_t c = ((size_t)b * R + xy_cell(py, g.oy, g.inv, R)) * R + xy_cell(px, g.ox, g.inv, R);
        const float* vb = verts + (size_t)b * n * 3;
        unsigned crossings = 0;
        auto crosses = [&](int f) -> bool {
            const float* A = vb + (size_t)faces[f * 3] * 3; const float* Bv = vb + (size_t)faces[f * 3 + 1] * 3; const float* Cv = vb + (size_t)faces[f * 3 + 2] * 3;
            double ax = A[0], ay = A[1], bx = Bv[0], by = Bv[1], cx = Cv[0], cy = Cv[1];
            double area = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
            if (area == 0.0) return false;                           // projects to a segment: never crossed
            bool ccw = area > 0.0;
            if (!edge_inside(ax, ay, bx, by, px, py, ccw) || !edge_inside(bx, by, cx, cy, px, py, ccw) || !edge_inside(cx, cy, ax, ay, px, py, ccw))
                return false;
            // z of the triangle's plane at (px, py)
            double w0 = ((bx - px) * (cy - py) - (by - py) * (cx - px)) / area;
            double w1 = ((cx - px) * (ay - py) - (cy - py) * (ax - px)) / area;
            double z = w0 * (double)A[2] + w1 * (double)Bv[2] + (1.0 - w0 - w1) * (double)Cv[2];
            return z > pz;
        };
        if (total && (size_t)*total > cap) {                         // the (cell, triangle) list did not fit: every face (same answer)
            for (int f = 0; f < m; ++f) crossings += crosses(f) ? 1u : 0u;
        } else {
            for (unsigned j = cell_start[c]; j < cell_end[c]; ++j) crossings += crosses(cell_list[j]) ? 1u : 0u;
        }
        inside = (crossings & 1u) != 0u;
    }
    out[(size_t)b * p + i] = inside ? 1 : 0;
}

// ---- A17 ----------------------------------------------------------------------------------------------
// edges (E,2) sorted by row (a, b), weight (E,) (= 1/deg(a)); row_start (V+1) built on the fly by binary search
__global__ void __launch_bounds__(256) lap_rows_kernel(const int32_t* __restrict__ edges, int E, int V, unsigned* __restrict__ row_start) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > V) return;
    int lo = 0, hi = E;                      // first edge with a >= v
    while (lo < hi) { int mid = (lo + hi) >> 1; if (edges[(size_t)mid * 2] < v) lo = mid + 1; else hi = mid; }
    row_start[v] = (unsigned)lo;
}
// r[b,v,:] = sum_j w_vj d[b,j,:] - d[b,v,:];  loss[b] += |r|^2
__global__ void __launch_bounds__(256) lap_fwd_kernel(const float* __restrict__ d, const int32_t* __restrict__ edges, const float* __restrict__ weight,
                                                      const unsigned* __restrict__ row_start, int V, float* __restrict__ r, double* __restrict__ acc) {
    int b = blockIdx.y;
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (v < V) {
        const float* db = d + (size_t)b * V * 3;
        float n0 = 0.f, n1 = 0.f, n2 = 0.f;
        for (unsigned e = row_start[v]; e < row_start[v + 1]; ++e) {
            int j = edges[(size_t)e * 2 + 1];
            float w = weight[e];
            n0 += w * db[(size_t)j * 3]; n1 += w * db[(size_t)j * 3 + 1]; n2 += w * db[(size_t)j * 3 + 2];
        }
        n0 -= db[(size_t)v * 3]; n1 -= db[(size_t)v * 3 + 1]; n2 -= db[(size_t)v * 3 + 2];
        float* rb = r + ((size_t)b * V + v) * 3;
        rb[0] = n0; rb[1] = n1; rb[2] = n2;
        s = (double)(n0 * n0 + n1 * n1 + n2 * n2);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(acc + b, s);
}
// grad d = 2 g (A^T r - r)  with A = D^-1 Adj
__global__ void __launch_bounds__(256) lap_bwd_kernel(const float* __restrict__ r, const int32_t* __restrict__ edges, const float* __restrict__ weight,
                                                      const unsigned* __restrict__ row_start, int V, const float* __restrict__ g_loss,
                                                      float* __restrict__ grad) {
    int b = blockIdx.y;
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float g2 = 2.f * g_loss[b];
    const float* rb = r + ((size_t)b * V + v) * 3;
    float* gb = grad + (size_t)b * V * 3;
    float r0 = g2 * rb[0], r1 = g2 * rb[1], r2 = g2 * rb[2];
    atomicAdd(gb + (size_t)v * 3, -r0); atomicAdd(gb + (size_t)v * 3 + 1, -r1); atomicAdd(gb + (size_t)v * 3 + 2, -r2);
    for (unsigned e = row_start[v]; e < row_start[v + 1]; ++e) {
        int j = edges[(size_t)e * 2 + 1];
        float w = weight[e];
        atomicAdd(gb + (size_t)j * 3, w * r0); atomicAdd(gb + (size_t)j * 3 + 1, w * r1); atomicAdd(gb + (size_t)j * 3 + 2, w * r2);
    }
}
__global__ void lap_finalize_kernel(const double* acc, int B, float* loss) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) loss[b] = (float)acc[b];
}

}  // namespace dtb

using namespace dtb;

extern "C" size_t dtb_check_sign_workspace(int B, int m, int R) {
    size_t cells = (size_t)B * R * R;
    // worst case every triangle spans a handful of cells; the pair list is sized after counting, so reserve generously
    return align_up((size_t)B * 6 * 4, 256) + 2 * align_up(cells * 4, 256) + align_up((size_t)B * m * 16 * 4, 256) + scan_workspace_bytes(cells) + 1024;
}
// verts (B,n,3) f32, faces (m,3) i32 shared by the batch, points (B,p,3) f32 -> out (B,p) u8 (1 = inside).
// R = cells per axis of the xy hash grid (kaolin's hash_resolution; <= 0 -> 256).
static int check_sign_impl(const float* verts, const int32_t* faces, const float* points, int B, int n, int m, int p, int R, bool fixed,
                           int* r_used, unsigned char* out, void* workspace, size_t workspace_bytes, void* stream) {
    if ((p == 0 && !r_used) || B == 0) return DTB_OK;
    DTB_REQUIRE((points && out) || p == 0, "check_sign: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0 || n == 0) { if (p) DTB_CUDA(cudaMemsetAsync(out, 0, (size_t)B * p, st)); if (r_used) *r_used = R; return DTB_OK; }
    DTB_REQUIRE(verts && faces, "check_sign: null mesh");
    if (R <= 0) R = 256;
    if (R > 1024) R = 1024;
    const int R_ws = R;                                   // the workspace was sized for this resolution (only ever shrinks below)
    size_t cells = (size_t)B * R_ws * R_ws;
    Workspace ws(workspace, workspace_bytes);
    unsigned* bbox = ws.take<unsigned>((size_t)B * 6);
    unsigned* cstart = ws.take<unsigned>(cells);
    unsigned* cend = ws.take<unsigned>(cells);
    size_t cap = (size_t)B * m * 16;
    int32_t* list = ws.take<int32_t>(cap);
    size_t sb = scan_workspace_bytes(cells);
    void* sws = ws.take<char>(sb);
    unsigned* total = ws.take<unsigned>(1);
    if (!ws.ok || !workspace) { set_error("check_sign: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    cs_init_kernel<<<cdiv(B * 6, 64), 64, 0, st>>>(bbox, B);
    dim3 gv(min(cdiv(n, 256), 64), B);
    cs_bbox_kernel<<<gv, 256, 0, st>>>(verts, n, bbox);
    DTB_LAUNCH_CHECK("cs_bbox");
    dim3 gf(cdiv(m, 256), B);
    // the (cell, triangle) list has a fixed capacity.  Blocking form: halve the grid resolution until it fits (one small blocking
    // copy per attempt).  Fixed form (a resolution found earlier for a mesh of this size): never synchronises -- should the list
    // overflow after all, the fill is skipped on the device and the query tests every face instead (same answer, slower).
    for (;;) {
        cells = (size_t)B * R * R;
        DTB_CUDA(cudaMemsetAsync(cstart, 0, cells * 4, st));
        cs_bin_kernel<<<gf, 256, 0, st>>>(verts, n, faces, m, R, bbox, cstart, nullptr, nullptr, 0);
        DTB_LAUNCH_CHECK("cs_bin_count");
        int rc = exclusive_scan_u32(cstart, cstart, cells, total, sws, sb, st);
        if (rc) return rc;
        if (fixed) break;
        unsigned h_total = 0;
        DTB_CUDA(cudaMemcpyAsync(&h_total, total, 4, cudaMemcpyDeviceToHost, st));
        DTB_CUDA(cudaStreamSynchronize(st));
        if ((size_t)h_total <= cap) break;
        if (R <= 1) { set_error("check_sign: %u (cell, triangle) pairs exceed the workspace capacity %zu", h_total, cap); return DTB_EOVERFLOW; }
        R = R > 2 ? R / 2 : 1;
    }
    if (r_used) *r_used = R;
    if (p == 0) return DTB_OK;
    DTB_CUDA(cudaMemcpyAsync(cend, cstart, cells * 4, cudaMemcpyDeviceToDevice, st));
    cs_bin_kernel<<<gf, 256, 0, st>>>(verts, n, faces, m, R, bbox, nullptr, cend, list, 1, total, cap);
    DTB_LAUNCH_CHECK("cs_bin_fill");
    dim3 gp(cdiv(p, 256), B);
    cs_query_kernel<<<gp, 256, 0, st>>>(verts, n, faces, R, bbox, cstart, cend, list, points, p, out, m, total, cap);
    DTB_LAUNCH_CHECK("cs_query");
    (void)R_ws;
    return DTB_OK;
}

extern "C" int dtb_check_sign(const float* verts, const int32_t* faces, const float* points, int B, int n, int m, int p, int R,
                              unsigned char* out, void* workspace, size_t workspace_bytes, void* stream) {
    return check_sign_impl(verts, faces, points, B, n, m, p, R, false, nullptr, out, workspace, workspace_bytes, stream);
}
// Same, and reports the grid resolution that was finally used (<= R): feed it to dtb_check_sign_fixed for later meshes of this size.
extern "C" int dtb_check_sign_probe(const float* verts, const int32_t* faces, const float* points, int B, int n, int m, int p, int R,
                                    unsigned char* out, int* r_used, void* workspace, size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(r_used, "check_sign_probe: null r_used");
    return check_sign_impl(verts, faces, points, B, n, m, p, R, false, r_used, out, workspace, workspace_bytes, stream);
}
// Non-blocking form: resolution R as given, no host synchronisation (capturable); a list overflow is handled on the device.
extern "C" int dtb_check_sign_fixed(const float* verts, const int32_t* faces, const float* points, int B, int n, int m, int p, int R,
                                    unsigned char* out, void* workspace, size_t workspace_bytes, void* stream) {
    return check_sign_impl(verts, faces, points, B, n, m, p, R, true, nullptr, out, workspace, workspace_bytes, stream);
}

// A17.  d (B,V,3) offsets; edges (E,2) i32 sorted by first column with weight (E,) from dtb_tet_point_adj.
// resid_ws (B,V,3) f32 and rows_ws (V+1) u32 are scratch kept for backward; acc (B,) f64.
extern "C" int dtb_laplacian_forward(const float* d, const int32_t* edges, const float* weight, int B, int V, int E, float* resid_ws,
                                     unsigned* rows_ws, double* acc, float* loss, void* stream) {
    DTB_REQUIRE(d && edges && weight && resid_ws && rows_ws && acc && loss, "laplacian_forward: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    lap_rows_kernel<<<cdiv(V + 1, 256), 256, 0, st>>>(edges, E, V, rows_ws);
    DTB_LAUNCH_CHECK("lap_rows");
    DTB_CUDA(cudaMemsetAsync(acc, 0, B * sizeof(double), st));
    dim3 g(cdiv(V, 256), B);
    lap_fwd_kernel<<<g, 256, 0, st>>>(d, edges, weight, rows_ws, V, resid_ws, acc);
    DTB_LAUNCH_CHECK("lap_fwd");
    lap_finalize_kernel<<<cdiv(B, 64), 64, 0, st>>>(acc, B, loss);
    DTB_LAUNCH_CHECK("lap_finalize");
    return DTB_OK;
}
extern "C" int dtb_laplacian_backward(const float* resid_ws, const int32_t* edges, const float* weight, const unsigned* rows_ws,
                                      const float* g_loss, int B, int V, float* grad, void* stream) {
    DTB_REQUIRE(resid_ws && edges && weight && rows_ws && g_loss && grad, "laplacian_backward: null argument");
    dim3 g(cdiv(V, 256), B);
    lap_bwd_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(resid_ws, edges, weight, rows_ws, V, g_loss, grad);
    DTB_LAUNCH_CHECK("lap_bwd");
    return DTB_OK;
}
