// A5: boundary-face <-> boundary-face adjacency through a shared edge, and the normal-consistency loss on it.
//   reference kernel   layers/DefTet/tet_face_adj_m_idx/tet_face_adj_m_for.cu:15-108 (O(F^2) scan, equality of
//                      vertex *coordinates*, at most n_max_nei = 30 neighbours in ascending face id)
//   reference consumer utils/mesh_utils.py:16-39 get_surface_normal_loss, :42-53 get_normal
// Here vertices are grouped first -- by exact coordinate value through an open-addressing hash table (soup
// form, what the drop-in receives) or simply by vertex id (indexed form, what the engine has) -- then a
// per-group incidence list (count / scan / fill) turns the search into O(F * degree).  Two faces are
// neighbours iff they share two vertex groups, which for non-degenerate triangles is exactly
// check_share() (:38-69).  Deviation, documented in DESIGN.md: the reference compares with sum|diff| <= 1e-15,
// which differs from exact equality only for coordinates below ~1e-8 in magnitude.
#include "prims.cuh"
#include "deftet_b200.h"

namespace dtb {

constexpr int ADJ_MAX = 30;      // n_max_nei (tet_face_adj_m_idx/utils.py:45)

__device__ __forceinline__ unsigned canon_bits(float f) {
    unsigned u = __float_as_uint(f);
    return (u == 0x80000000u) ? 0u : u;         // -0.0 == +0.0
}
__device__ __forceinline__ unsigned hash3(unsigned a, unsigned b, unsigned c) {
    unsigned h = a * 0x9E3779B1u;
    h ^= (h >> 15); h += b * 0x85EBCA77u; h ^= (h >> 13); h += c * 0xC2B2AE3Du; h ^= (h >> 16);
    h *= 0x27D4EB2Fu; h ^= (h >> 15);
    return h;
}

// ---- grouping by coordinates (soup form) --------------------------------------------------------------
__global__ void __launch_bounds__(256) fa_hash_insert_kernel(const float* __restrict__ soup, const int32_t* __restrict__ counts, int Fmax,
                                                             int H, int* __restrict__ slots) {
    int b = blockIdx.y;
    int c = blockIdx.x * blockDim.x + threadIdx.x;       // corner
    int n = (counts ? counts[b] : Fmax) * 3;
    if (c >= n) return;
    const float* sb = soup + (size_t)b * Fmax * 9;
    unsigned kx = canon_bits(sb[c * 3]), ky = canon_bits(sb[c * 3 + 1]), kz = canon_bits(sb[c * 3 + 2]);
    unsigned h = hash3(kx, ky, kz) & (H - 1);
    int* tb = slots + (size_t)b * H;
    for (int probe = 0; probe < H; ++probe) {
        int cur = tb[h];
        if (cur < 0) {
            int old = atomicCAS(tb + h, -1, c);
            if (old < 0) return;
            cur = old;
        }
        if (canon_bits(sb[cur * 3]) == kx && canon_bits(sb[cur * 3 + 1]) == ky && canon_bits(sb[cur * 3 + 2]) == kz) {
            atomicMin(tb + h, c);
            return;
        }
        h = (h + 1) & (H - 1);
    }
}
__global__ void __launch_bounds__(256) fa_hash_lookup_kernel(const float* __restrict__ soup, const int32_t* __restrict__ counts, int Fmax,
                                                             int H, const int* __restrict__ slots, int* __restrict__ group,
                                                             unsigned* __restrict__ gcount) {
    int b = blockIdx.y;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int n = (counts ? counts[b] : Fmax) * 3;
    if (c >= n) return;
    const float* sb = soup + (size_t)b * Fmax * 9;
    unsigned kx = canon_bits(sb[c * 3]), ky = canon_bits(sb[c * 3 + 1]), kz = canon_bits(sb[c * 3 + 2]);
    unsigned h = hash3(kx, ky, kz) & (H - 1);
    const int* tb = slots + (size_t)b * H;
    int g = -1;
    for (int probe = 0; probe < H; ++probe) {
        int cur = tb[h];
        if (cur < 0) break;
        if (canon_bits(sb[cur * 3]) == kx && canon_bits(sb[cur * 3 + 1]) == ky && canon_bits(sb[cur * 3 + 2]) == kz) { g = (int)h; break; }
        h = (h + 1) & (H - 1);
    }
    group[(size_t)b * Fmax * 3 + c] = g;
    if (g >= 0) atomicAdd(gcount + (size_t)b * H + g, 1u);
}
// ---- grouping by vertex id (indexed form) ----------------------------------------------------------------
__global__ void __launch_bounds__(256) fa_index_group_kernel(const int32_t* __restrict__ faces, const int32_t* __restrict__ counts, int Fmax,
                                                             int V, int* __restrict__ group, unsigned* __restrict__ gcount) {
    int b = blockIdx.y;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int n = (counts ? counts[b] : Fmax) * 3;
    if (c >= n) return;
    int v = faces[(size_t)b * Fmax * 3 + c];
    group[(size_t)b * Fmax * 3 + c] = v;
    atomicAdd(gcount + (size_t)b * V + v, 1u);
}
__global__ void __launch_bounds__(256) fa_fill_kernel(const int* __restrict__ group, const int32_t* __restrict__ counts, int Fmax, int NG,
                                                      unsigned* __restrict__ gcursor, int* __restrict__ glist) {
    int b = blockIdx.y;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int n = (counts ? counts[b] : Fmax) * 3;
    if (c >= n) return;
    int g = group[(size_t)b * Fmax * 3 + c];
    if (g < 0) return;
    unsigned dst = atomicAdd(gcursor + (size_t)b * NG + g, 1u);
    glist[dst] = c;
}

// keep the ADJ_MAX smallest distinct ids, sorted ascending
__device__ __forceinline__ void insert_sorted(int* arr, int& n, int g) {
    int lo = 0;
    while (lo < n && arr[lo] < g) ++lo;
    if (lo < n && arr[lo] == g) return;
    if (lo >= ADJ_MAX) return;
    int last = (n < ADJ_MAX) ? n : ADJ_MAX - 1;
    for (int k = last; k > lo; --k) arr[k] = arr[k - 1];
    arr[lo] = g;
    if (n < ADJ_MAX) ++n;
}

__global__ void __launch_bounds__(128) fa_neighbour_kernel(const int* __restrict__ group, const int32_t* __restrict__ counts, int Fmax, int NG,
                                                           const unsigned* __restrict__ gstart, const unsigned* __restrict__ gend,
                                                           const int* __restrict__ glist, float* __restrict__ adj_f32,
                                                           int32_t* __restrict__ adj_i32, int32_t* __restrict__ deg) {
    int b = blockIdx.y;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    int nf = counts ? counts[b] : Fmax;
    if (f >= nf) return;
    const int* gb = group + (size_t)b * Fmax * 3;
    int g[3] = {gb[f * 3], gb[f * 3 + 1], gb[f * 3 + 2]};
    unsigned s[3], e[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (g[k] >= 0) { s[k] = gstart[(size_t)b * NG + g[k]]; e[k] = gend[(size_t)b * NG + g[k]]; }
        else { s[k] = e[k] = 0; }
    }
    int nb[ADJ_MAX];
    int n = 0;
    // a face sharing group(k) and group(m), k < m, shares the edge (k, m)
#pragma unroll
    for (int k = 0; k < 2; ++k)
        for (unsigned i = s[k]; i < e[k]; ++i) {
            int of = glist[i] / 3;
            if (of == f) continue;
            bool shared = false;
            for (int m = k + 1; m < 3 && !shared; ++m) {
                if (g[m] == g[k]) continue;              // degenerate: the same vertex twice is not an edge
                for (unsigned j = s[m]; j < e[m]; ++j)
                    if (glist[j] / 3 == of) { shared = true; break; }
            }
            if (shared) insert_sorted(nb, n, of);
        }
    if (adj_f32) {
        float* o = adj_f32 + ((size_t)b * Fmax + f) * ADJ_MAX;
        for (int k = 0; k < ADJ_MAX; ++k) o[k] = k < n ? (float)nb[k] : -1.0f;
    }
    if (adj_i32) {
        int32_t* o = adj_i32 + ((size_t)b * Fmax + f) * ADJ_MAX;
        for (int k = 0; k < ADJ_MAX; ++k) o[k] = k < n ? nb[k] : -1;
    }
    if (deg) deg[(size_t)b * Fmax + f] = n;
}

// ---- normal-consistency loss on the adjacency (engine form) ---------------------------------------------
// n = cross(b-a, c-a) / sqrt(|cross|^2 + 1e-12)  (mesh_utils.py:42-53); loss_b = mean_pairs (1 - n_i . n_j)
__device__ __forceinline__ void face_normal(const float* p, const int32_t* fi, float* c, float& inv) {
    const float* a = p + (size_t)fi[0] * 3; const float* b = p + (size_t)fi[1] * 3; const float* d = p + (size_t)fi[2] * 3;
    float u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, v[3] = {d[0] - a[0], d[1] - a[1], d[2] - a[2]};
    c[0] = u[1] * v[2] - u[2] * v[1]; c[1] = u[2] * v[0] - u[0] * v[2]; c[2] = u[0] * v[1] - u[1] * v[0];
    inv = 1.0f / sqrtf(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + 1e-12f);
}
__global__ void __launch_bounds__(256) nl_normals_kernel(const float* __restrict__ pos, int V, const int32_t* __restrict__ faces,
                                                         const int32_t* __restrict__ counts, int Fmax, float* __restrict__ normals) {
    int b = blockIdx.y;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= counts[b]) return;
    float c[3], inv;
    face_normal(pos + (size_t)b * V * 3, faces + ((size_t)b * Fmax + f) * 3, c, inv);
    float* o = normals + ((size_t)b * Fmax + f) * 3;
    o[0] = c[0] * inv; o[1] = c[1] * inv; o[2] = c[2] * inv;
}
// acc[b*2] += sum (1 - ni.nj), acc[b*2+1] += number of pairs
__global__ void __launch_bounds__(256) nl_forward_kernel(const float* __restrict__ normals, const int32_t* __restrict__ adj,
                                                         const int32_t* __restrict__ counts, int Fmax, double* __restrict__ acc) {
    int b = blockIdx.y;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0, cnt = 0.0;
    if (f < counts[b]) {
        const float* nb = normals + (size_t)b * Fmax * 3;
        float n0 = nb[f * 3], n1 = nb[f * 3 + 1], n2 = nb[f * 3 + 2];
        const int32_t* a = adj + ((size_t)b * Fmax + f) * ADJ_MAX;
        for (int k = 0; k < ADJ_MAX; ++k) {
            int j = a[k];
            if (j < 0) break;
            s += (double)(1.0f - (n0 * nb[j * 3] + n1 * nb[j * 3 + 1] + n2 * nb[j * 3 + 2]));
            cnt += 1.0;
        }
    }
    s = warp_sum(s); cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0 && cnt > 0.0) { atomicAdd(acc + b * 2, s); atomicAdd(acc + b * 2 + 1, cnt); }
}
__global__ void nl_finalize_kernel(const double* __restrict__ acc, const int32_t* __restrict__ counts, int B, float* __restrict__ loss) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (counts[b] == 0) loss[b] = 1.0f;                          // empty surface (deftet.py:162-166)
    else loss[b] = acc[b * 2 + 1] > 0.0 ? (float)(acc[b * 2] / acc[b * 2 + 1]) : 0.0f;   // no pairs -> 0 (mesh_utils.py:29-33)
}
// d loss / d n_i = -(w) * sum over pairs containing i; pairs are directed (i,j) rows of the adjacency
__global__ void __launch_bounds__(256) nl_backward_pairs_kernel(const float* __restrict__ normals, const int32_t* __restrict__ adj,
                                                                const int32_t* __restrict__ counts, int Fmax, const double* __restrict__ acc,
                                                                const float* __restrict__ g_loss, float* __restrict__ gn) {
    int b = blockIdx.y;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= counts[b]) return;
    double E = acc[b * 2 + 1];
    if (!(E > 0.0)) return;
    float w = -g_loss[b] / (float)E;
    const float* nb = normals + (size_t)b * Fmax * 3;
    float* gb = gn + (size_t)b * Fmax * 3;
    float n0 = nb[f * 3], n1 = nb[f * 3 + 1], n2 = nb[f * 3 + 2];
    const int32_t* a = adj + ((size_t)b * Fmax + f) * ADJ_MAX;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < ADJ_MAX; ++k) {
        int j = a[k];
        if (j < 0) break;
        s0 += nb[j * 3]; s1 += nb[j * 3 + 1]; s2 += nb[j * 3 + 2];
        atomicAdd(gb + j * 3, w * n0); atomicAdd(gb + j * 3 + 1, w * n1); atomicAdd(gb + j * 3 + 2, w * n2);
    }
    atomicAdd(gb + f * 3, w * s0); atomicAdd(gb + f * 3 + 1, w * s1); atomicAdd(gb + f * 3 + 2, w * s2);
}
// chain rule n = c * (|c|^2 + eps)^-1/2, c = (b-a) x (d-a)  -> vertex gradients
__global__ void __launch_bounds__(256) nl_backward_vertices_kernel(const float* __restrict__ pos, int V, const int32_t* __restrict__ faces,
                                                                   const int32_t* __restrict__ counts, int Fmax, const float* __restrict__ gn,
                                                                   float* __restrict__ grad_pos, int gstride) {
    int b = blockIdx.y;
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= counts[b]) return;
    const float* p = pos + (size_t)b * V * 3;
    const int32_t* fi = faces + ((size_t)b * Fmax + f) * 3;
    float c[3], inv;
    face_normal(p, fi, c, inv);
    const float* g = gn + ((size_t)b * Fmax + f) * 3;
    float gd = g[0] * c[0] + g[1] * c[1] + g[2] * c[2];
    float inv3 = inv * inv * inv;
    float gc[3] = {g[0] * inv - gd * inv3 * c[0], g[1] * inv - gd * inv3 * c[1], g[2] * inv - gd * inv3 * c[2]};
    const float* a = p + (size_t)fi[0] * 3; const float* bb = p + (size_t)fi[1] * 3; const float* d = p + (size_t)fi[2] * 3;
    float u[3] = {bb[0] - a[0], bb[1] - a[1], bb[2] - a[2]}, v[3] = {d[0] - a[0], d[1] - a[1], d[2] - a[2]};
    // c = u x v : dL/du = v x gc, dL/dv = gc x u
    float gu[3] = {v[1] * gc[2] - v[2] * gc[1], v[2] * gc[0] - v[0] * gc[2], v[0] * gc[1] - v[1] * gc[0]};
    float gv[3] = {gc[1] * u[2] - gc[2] * u[1], gc[2] * u[0] - gc[0] * u[2], gc[0] * u[1] - gc[1] * u[0]};
    float* gp = grad_pos + (size_t)b * V * gstride;
    grad_add3(gp, (size_t)fi[1], gstride, gu[0], gu[1], gu[2]);
    grad_add3(gp, (size_t)fi[2], gstride, gv[0], gv[1], gv[2]);
    grad_add3(gp, (size_t)fi[0], gstride, -gu[0] - gv[0], -gu[1] - gv[1], -gu[2] - gv[2]);
}

}  // namespace dtb

using namespace dtb;

static int next_pow2(long long x) { int p = 16; while (p < x) p <<= 1; return p; }

// NG = groups per sample: hash slots (soup) or V (indexed)
static size_t face_adj_ws(int B, int Fmax, long long NG, bool soup) {
    Workspace ws(nullptr, 0);
    if (soup) ws.take<int>((size_t)B * NG);
    ws.take<int>((size_t)B * Fmax * 3);            // group
    ws.take<unsigned>((size_t)B * NG);             // gstart
    ws.take<unsigned>((size_t)B * NG);             // gend
    ws.take<int>((size_t)B * Fmax * 3);            // glist
    ws.take<char>(scan_workspace_bytes((size_t)B * NG));
    return ws.off + 256;
}

extern "C" size_t dtb_face_adjacency_workspace(int B, int Fmax, int V) {
    long long NG = V > 0 ? V : next_pow2(6LL * Fmax);
    return face_adj_ws(B, Fmax, NG, V <= 0);
}

// soup != NULL: group by coordinates (drop-in); else faces (B,Fmax,3) i32 + V: group by vertex id.
// counts may be NULL (all Fmax faces valid).  Outputs (any may be NULL): adj_f32 / adj_i32 (B,Fmax,30) padded
// with -1, ascending; deg (B,Fmax).
extern "C" int dtb_face_adjacency(const float* soup, const int32_t* faces, const int32_t* counts, int B, int Fmax, int V, float* adj_f32,
                                  int32_t* adj_i32, int32_t* deg, void* workspace, size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(soup || (faces && V > 0), "face_adjacency: need a face soup or (faces, V)");
    DTB_REQUIRE(B > 0 && Fmax >= 0, "face_adjacency: bad sizes");
    if (Fmax == 0) return DTB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool use_soup = soup != nullptr;
    long long NGl = use_soup ? next_pow2(6LL * Fmax) : V;
    DTB_REQUIRE((long long)B * NGl < (1LL << 31), "face_adjacency: too many groups");
    int NG = (int)NGl;
    Workspace ws(workspace, workspace_bytes);
    int* slots = use_soup ? ws.take<int>((size_t)B * NG) : nullptr;
    int* group = ws.take<int>((size_t)B * Fmax * 3);
    unsigned* gstart = ws.take<unsigned>((size_t)B * NG);
    unsigned* gend = ws.take<unsigned>((size_t)B * NG);
    int* glist = ws.take<int>((size_t)B * Fmax * 3);
    size_t sb = scan_workspace_bytes((size_t)B * NG);
    void* sws = ws.take<char>(sb);
    if (!ws.ok || !workspace) { set_error("face_adjacency: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    dim3 gc(cdiv((long long)Fmax * 3, 256), B);
    DTB_CUDA(cudaMemsetAsync(gstart, 0, (size_t)B * NG * sizeof(unsigned), st));
    if (use_soup) {
        DTB_CUDA(cudaMemsetAsync(slots, 0xff, (size_t)B * NG * sizeof(int), st));
        fa_hash_insert_kernel<<<gc, 256, 0, st>>>(soup, counts, Fmax, NG, slots);
        DTB_LAUNCH_CHECK("fa_hash_insert");
        fa_hash_lookup_kernel<<<gc, 256, 0, st>>>(soup, counts, Fmax, NG, slots, group, gstart);
        DTB_LAUNCH_CHECK("fa_hash_lookup");
    } else {
        fa_index_group_kernel<<<gc, 256, 0, st>>>(faces, counts, Fmax, V, group, gstart);
        DTB_LAUNCH_CHECK("fa_index_group");
    }
    int rc = exclusive_scan_u32_dup(gstart, gstart, gend, (size_t)B * NG, nullptr, sws, sb, st);
    if (rc) return rc;
    fa_fill_kernel<<<gc, 256, 0, st>>>(group, counts, Fmax, NG, gend, glist);
    DTB_LAUNCH_CHECK("fa_fill");
    dim3 gf(cdiv(Fmax, 128), B);
    fa_neighbour_kernel<<<gf, 128, 0, st>>>(group, counts, Fmax, NG, gstart, gend, glist, adj_f32, adj_i32, deg);
    DTB_LAUNCH_CHECK("fa_neighbour");
    return DTB_OK;
}

// normals_ws: (B,Fmax,3) f32 scratch kept for backward; acc: (B,2) f64
extern "C" int dtb_normal_loss_forward(const float* pos, const int32_t* faces, const int32_t* counts, const int32_t* adj, int B, int V,
                                       int Fmax, float* normals_ws, double* acc, float* loss, void* stream) {
    DTB_REQUIRE(pos && faces && counts && adj && normals_ws && acc && loss, "normal_loss_forward: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    DTB_CUDA(cudaMemsetAsync(acc, 0, (size_t)B * 2 * sizeof(double), st));
    if (Fmax > 0) {
        dim3 g(cdiv(Fmax, 256), B);
        nl_normals_kernel<<<g, 256, 0, st>>>(pos, V, faces, counts, Fmax, normals_ws);
        DTB_LAUNCH_CHECK("nl_normals");
        nl_forward_kernel<<<g, 256, 0, st>>>(normals_ws, adj, counts, Fmax, acc);
        DTB_LAUNCH_CHECK("nl_forward");
    }
    nl_finalize_kernel<<<cdiv(B, 64), 64, 0, st>>>(acc, counts, B, loss);
    DTB_LAUNCH_CHECK("nl_finalize");
    return DTB_OK;
}

// gn_ws: (B,Fmax,3) f32 scratch (zero-filled here); accumulates into grad_pos
extern "C" int dtb_normal_loss_backward(const float* pos, const int32_t* faces, const int32_t* counts, const int32_t* adj, const float* normals_ws,
                                        const double* acc, const float* g_loss, int B, int V, int Fmax, float* gn_ws, float* grad_pos,
                                        int grad_stride, void* stream) {
    DTB_REQUIRE(pos && faces && counts && adj && normals_ws && acc && g_loss && gn_ws && grad_pos, "normal_loss_backward: null argument");
    if (Fmax == 0) return DTB_OK;
    DTB_REQUIRE(grad_stride == 3 || (grad_stride == 4 && (((size_t)grad_pos) & 15) == 0), "normal_loss_backward: bad grad_stride / alignment");
    cudaStream_t st = (cudaStream_t)stream;
    DTB_CUDA(cudaMemsetAsync(gn_ws, 0, (size_t)B * Fmax * 3 * sizeof(float), st));
    dim3 g(cdiv(Fmax, 256), B);
    nl_backward_pairs_kernel<<<g, 256, 0, st>>>(normals_ws, adj, counts, Fmax, acc, g_loss, gn_ws);
    DTB_LAUNCH_CHECK("nl_backward_pairs");
    nl_backward_vertices_kernel<<<g, 256, 0, st>>>(pos, V, faces, counts, Fmax, gn_ws, grad_pos, grad_stride);
    DTB_LAUNCH_CHECK("nl_backward_vertices");
    return DTB_OK;
}
