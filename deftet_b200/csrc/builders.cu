// A10-A14: topology builders on the GPU (sort / scan / compact instead of std::map and Python dicts).
//   A10 vertex adjacency   utils/lib/tet_point_adj/run.cpp:20-56   (unique directed vertex pairs)
//   A11 face table         utils/tet_utils.py:208-256 tet_to_face  (first-occurrence order; + boundary list)
//   A12 tet-tet sharing    utils/lib/tet_adj_share/run.cpp:40-97   (ascending face-key order)
//   A13 face-face adjacency utils/lib/tet_face_adj/run.cpp:18-92   (wrapped int32 edge key, ordered pairs per edge)
//   A14 vertex collapse    utils/lib/colaps_v/run.cpp:18-59        ("%.5f" string key, first-occurrence ids)
// Output orders are the reference's, except A10 whose reference order is libstdc++'s unordered_set iteration
// order (unspecified): edges are emitted sorted by (a, b).
#include "prims.cuh"
#include "deftet_b200.h"

namespace dtb {

typedef unsigned long long u64;

__device__ __constant__ int LOCAL_FACE[4][3] = {{0, 1, 2}, {1, 0, 3}, {2, 3, 0}, {3, 2, 1}};   // idx_array (tet_utils.py:213-217)

__device__ __forceinline__ void tet_face_verts(const int32_t* __restrict__ tet, int slot, int& a, int& b, int& c) {
    int t = slot >> 2, i = slot & 3;
    const int32_t* q = tet + (size_t)t * 4;
    a = q[LOCAL_FACE[i][0]]; b = q[LOCAL_FACE[i][1]]; c = q[LOCAL_FACE[i][2]];
}
// key = min*n^2 + max*n + mid, "mid" = the last vertex that is neither min nor max (defaults as the reference does)
__device__ __forceinline__ u64 face_key(int a, int b, int c, u64 n, int mid_default) {
    int lo = min(a, min(b, c)), hi = max(a, max(b, c));
    int mid = mid_default;
    if (a != lo && a != hi) mid = a;
    if (b != lo && b != hi) mid = b;
    if (c != lo && c != hi) mid = c;
    return (u64)lo * n * n + (u64)hi * n + (u64)mid;
}

// ---- shared: sorted tet-face keys -----------------------------------------------------------------------
__global__ void __launch_bounds__(256) face_keys_kernel(const int32_t* __restrict__ tet, int T, u64 n, u64* __restrict__ keys,
                                                        unsigned* __restrict__ vals) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T * 4) return;
    int a, b, c;
    tet_face_verts(tet, s, a, b, c);
    keys[s] = face_key(a, b, c, n, c);       // tet_to_face / tet_adj_share leave `c` at the third vertex when undefined
    vals[s] = (unsigned)s;
}

// run bookkeeping over sorted keys: for every sorted position the length of its run if it is the run start, else 0
__global__ void __launch_bounds__(256) run_lengths_kernel(const u64* __restrict__ keys, size_t n, unsigned* __restrict__ run_len) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    u64 k = keys[p];
    if (p > 0 && keys[p - 1] == k) { run_len[p] = 0; return; }
    unsigned len = 1;
    while (p + len < n && keys[p + len] == k) ++len;
    run_len[p] = len;
}

// ---- A11 --------------------------------------------------------------------------------------------------
// slot-indexed marks: pair_flag[s0] = 1 and partner[s0] = s1 for faces shared by exactly two tets (s0 < s1, the
// first occurrence); single_flag[s] = 1 for faces seen once.
__global__ void __launch_bounds__(256) face_mark_kernel(const unsigned* __restrict__ run_len, const unsigned* __restrict__ sorted_slot, size_t n,
                                                        unsigned* __restrict__ pair_flag, unsigned* __restrict__ partner,
                                                        unsigned* __restrict__ single_flag) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    unsigned len = run_len[p];
    if (len == 2) {
        unsigned s0 = sorted_slot[p], s1 = sorted_slot[p + 1];       // stable sort: s0 < s1
        pair_flag[s0] = 1u; partner[s0] = s1;
    } else if (len == 1) {
        single_flag[sorted_slot[p]] = 1u;
    }
}
__global__ void __launch_bounds__(256) face_emit_kernel(const int32_t* __restrict__ tet, size_t n, const unsigned* __restrict__ pair_flag,
                                                        const unsigned* __restrict__ pair_pos, const unsigned* __restrict__ partner,
                                                        const unsigned* __restrict__ single_flag, const unsigned* __restrict__ single_pos,
                                                        int32_t* __restrict__ face_fx3, int32_t* __restrict__ face_tet_fx2,
                                                        int32_t* __restrict__ face_slot_fx2, int32_t* __restrict__ boundary_fx3) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    if (pair_flag[s]) {
        unsigned o = pair_pos[s], s1 = partner[s];
        int a, b, c;
        tet_face_verts(tet, (int)s, a, b, c);
        if (face_fx3) { face_fx3[o * 3] = a; face_fx3[o * 3 + 1] = b; face_fx3[o * 3 + 2] = c; }
        if (face_tet_fx2) { face_tet_fx2[o * 2] = (int)(s >> 2); face_tet_fx2[o * 2 + 1] = (int)(s1 >> 2); }
        if (face_slot_fx2) { face_slot_fx2[o * 2] = (int)(s & 3); face_slot_fx2[o * 2 + 1] = (int)(s1 & 3); }
    }
    if (single_flag[s] && boundary_fx3) {
        unsigned o = single_pos[s];
        int a, b, c;
        tet_face_verts(tet, (int)s, a, b, c);
        boundary_fx3[o * 3] = a; boundary_fx3[o * 3 + 1] = b; boundary_fx3[o * 3 + 2] = c;
    }
}

// ---- A12 --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) share_flag_kernel(const unsigned* __restrict__ run_len, size_t n, unsigned* __restrict__ flag) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) flag[p] = run_len[p] == 2 ? 1u : 0u;
}
__global__ void __launch_bounds__(256) share_emit_kernel(const unsigned* __restrict__ flag, const unsigned* __restrict__ pos,
                                                         const unsigned* __restrict__ sorted_slot, size_t n, int32_t* __restrict__ out) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || !flag[p]) return;
    unsigned s0 = sorted_slot[p], s1 = sorted_slot[p + 1];
    int32_t* o = out + (size_t)pos[p] * 6;
    o[0] = (int)(s0 >> 2); o[1] = (int)(s1 >> 2); o[2] = (int)(s0 & 3);       // run.cpp:83-88
    o[3] = (int)(s1 >> 2); o[4] = (int)(s0 >> 2); o[5] = (int)(s1 & 3);
}

// ---- A10 --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) edge_keys_kernel(const int32_t* __restrict__ tet, int T, u64 n, u64* __restrict__ keys,
                                                        unsigned* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;      // 12 directed pairs per tet
    if (i >= T * 12) return;
    int t = i / 12, e = i % 12;
    int a = e / 3, b = e % 3;
    if (b >= a) ++b;
    keys[i] = (u64)tet[(size_t)t * 4 + a] * n + (u64)tet[(size_t)t * 4 + b];
    vals[i] = (unsigned)i;
}
__global__ void __launch_bounds__(256) unique_flag_kernel(const u64* __restrict__ keys, size_t n, unsigned* __restrict__ flag) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) flag[p] = (p == 0 || keys[p] != keys[p - 1]) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) edge_emit_kernel(const u64* __restrict__ keys, const unsigned* __restrict__ flag,
                                                        const unsigned* __restrict__ pos, size_t n, u64 nv, int32_t* __restrict__ edges,
                                                        unsigned* __restrict__ degree) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || !flag[p]) return;
    u64 k = keys[p];
    int a = (int)(k / nv), b = (int)(k % nv);
    edges[(size_t)pos[p] * 2] = a; edges[(size_t)pos[p] * 2 + 1] = b;
    if (degree) atomicAdd(degree + a, 1u);
}
__global__ void __launch_bounds__(256) edge_weight_kernel(const int32_t* __restrict__ edges, const unsigned* __restrict__ n_edge,
                                                          const unsigned* __restrict__ degree, float* __restrict__ weight, size_t cap) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= cap || p >= *n_edge) return;
    weight[p] = (float)(1.0 / (double)degree[edges[p * 2]]);         // D^-1 A (tet_point_adj/interface.py:42-54)
}

// ---- A13 --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fedge_keys_kernel(const int32_t* __restrict__ tet, int T, unsigned n, u64* __restrict__ keys,
                                                         unsigned* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;      // 4 faces x 3 edges per tet, in push order
    if (i >= T * 12) return;
    int slot = i / 3, e = i % 3;
    int tri[3];
    tet_face_verts(tet, slot, tri[0], tri[1], tri[2]);
    int p = tri[e], q = tri[(e + 1) % 3];
    unsigned a = (unsigned)min(p, q), b = (unsigned)max(p, q);
    unsigned k = a * n + b;                             // int e = point_a * n_point + point_b (wraps, run.cpp:39)
    keys[i] = (u64)(k ^ 0x80000000u);                   // std::map<int,...> iterates in signed order
    vals[i] = (unsigned)i;
}
__device__ __forceinline__ u64 abs_face_key(const int32_t* tet, int slot, u64 n) {
    int a, b, c;
    tet_face_verts(tet, slot, a, b, c);
    return face_key(a, b, c, n, a);                     // face_p_c defaults to triangle[0] (run.cpp:44-58)
}
// per sorted entry: number of (entry, other) pairs it emits
__global__ void __launch_bounds__(256) fadj_count_kernel(const int32_t* __restrict__ tet, const u64* __restrict__ keys,
                                                         const unsigned* __restrict__ seq, size_t n, u64 nv, unsigned* __restrict__ cnt,
                                                         int32_t* __restrict__ out, const unsigned* __restrict__ pos, size_t cap) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    u64 k = keys[p];
    size_t lo = p, hi = p;
    while (lo > 0 && keys[lo - 1] == k) --lo;
    while (hi + 1 < n && keys[hi + 1] == k) ++hi;
    int fa = (int)(seq[p] / 3);
    u64 ka = abs_face_key(tet, fa, nv);
    unsigned c = 0;
    size_t o = pos ? pos[p] : 0;
    for (size_t j = lo; j <= hi; ++j) {
        int fb = (int)(seq[j] / 3);
        if (fb == fa) continue;
        if (abs_face_key(tet, fb, nv) == ka) continue;
        if (out) {
            if (o + c < cap) { out[(o + c) * 2] = fa; out[(o + c) * 2 + 1] = fb; }
        }
        ++c;
    }
    if (cnt) cnt[p] = c;
}

// ---- A14 --------------------------------------------------------------------------------------------------
// key of one coordinate = what "%.5f" prints: sign character + round-half-even(|x| * 1e5), computed exactly
__device__ __forceinline__ u64 decimal5_key(float x) {
    unsigned u = __float_as_uint(x);
    u64 sign = (u64)(u >> 31) << 63;
    unsigned ex = (u >> 23) & 0xffu;
    unsigned man = u & 0x7fffffu;
    if (ex == 0xffu) return sign | 0x7ff0000000000000ull | man;        // inf / nan print as text
    u64 m = ex ? (u64)(man | 0x800000u) : (u64)man;
    int e = (ex ? (int)ex : 1) - 150;                                   // |x| = m * 2^e
    if (e >= 0) return sign | 0x4000000000000000ull | (u64)(u & 0x7fffffffu);   // integers >= 2^23: distinct floats, distinct strings
    u64 N = m * 100000ull;                                              // < 2^41
    int sh = -e;
    u64 q;
    if (sh > 62) q = 0;
    else {
        q = N >> sh;
        u64 rem = N & ((1ull << sh) - 1ull), half = 1ull << (sh - 1);
        if (rem > half || (rem == half && (q & 1ull))) ++q;
    }
    return sign | q;
}
__device__ __forceinline__ unsigned hash_u64x3(u64 a, u64 b, u64 c) {
    u64 h = a * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h += b * 0xBF58476D1CE4E5B9ull; h ^= h >> 31; h += c * 0x94D049BB133111EBull; h ^= h >> 30;
    h *= 0xD6E8FEB86659FD93ull; h ^= h >> 32;
    return (unsigned)h;
}
__global__ void __launch_bounds__(256) collapse_insert_kernel(const float* __restrict__ pts, int N, unsigned H, int* __restrict__ slots) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    u64 kx = decimal5_key(pts[(size_t)i * 3]), ky = decimal5_key(pts[(size_t)i * 3 + 1]), kz = decimal5_key(pts[(size_t)i * 3 + 2]);
    unsigned h = hash_u64x3(kx, ky, kz) & (H - 1);
    for (unsigned probe = 0; probe < H; ++probe) {
        int cur = slots[h];
        if (cur < 0) {
            int old = atomicCAS(slots + h, -1, i);
            if (old < 0) return;
            cur = old;
        }
        if (decimal5_key(pts[(size_t)cur * 3]) == kx && decimal5_key(pts[(size_t)cur * 3 + 1]) == ky && decimal5_key(pts[(size_t)cur * 3 + 2]) == kz) {
            atomicMin(slots + h, i);
            return;
        }
        h = (h + 1) & (H - 1);
    }
}
__global__ void __launch_bounds__(256) collapse_lookup_kernel(const float* __restrict__ pts, int N, unsigned H, const int* __restrict__ slots,
                                                              int32_t* __restrict__ rep, unsigned* __restrict__ is_first) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    u64 kx = decimal5_key(pts[(size_t)i * 3]), ky = decimal5_key(pts[(size_t)i * 3 + 1]), kz = decimal5_key(pts[(size_t)i * 3 + 2]);
    unsigned h = hash_u64x3(kx, ky, kz) & (H - 1);
    int r = i;
    for (unsigned probe = 0; probe < H; ++probe) {
        int cur = slots[h];
        if (cur < 0) break;
        if (decimal5_key(pts[(size_t)cur * 3]) == kx && decimal5_key(pts[(size_t)cur * 3 + 1]) == ky && decimal5_key(pts[(size_t)cur * 3 + 2]) == kz) { r = cur; break; }
        h = (h + 1) & (H - 1);
    }
    rep[i] = r;
    is_first[i] = (r == i) ? 1u : 0u;
}
__global__ void __launch_bounds__(256) collapse_emit_kernel(const int32_t* __restrict__ rep, const unsigned* __restrict__ is_first,
                                                            const unsigned* __restrict__ uid, int N, int32_t* __restrict__ map_array,
                                                            int32_t* __restrict__ inverse_idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    map_array[i] = (int)uid[rep[i]];
    if (is_first[i]) inverse_idx[uid[i]] = i;
}

// ---- N3: face table with boundary faces in one first-occurrence order, and the tet neighbour table ---------------
//   3_model/prepare_for_wz.py:49-108 tet_to_face_idx(with_boundary=True);  3_model/utils_tetsv.py:16-62 tet_neighbour_idx
__global__ void __launch_bounds__(256) face_mark_all_kernel(const unsigned* __restrict__ run_len, const unsigned* __restrict__ sorted_slot,
                                                            size_t n, unsigned* __restrict__ flag, unsigned* __restrict__ partner) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    unsigned len = run_len[p];
    if (len == 1 || len == 2) {
        unsigned s0 = sorted_slot[p];
        flag[s0] = 1u;
        partner[s0] = len == 2 ? sorted_slot[p + 1] : 0xffffffffu;
    }
}
__global__ void __launch_bounds__(256) face_emit_all_kernel(const int32_t* __restrict__ tet, size_t n, const unsigned* __restrict__ flag,
                                                            const unsigned* __restrict__ pos, const unsigned* __restrict__ partner,
                                                            int32_t* __restrict__ face_fx3, int32_t* __restrict__ face_tet_fx2,
                                                            int32_t* __restrict__ face_slot_fx2) {
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || !flag[s]) return;
    unsigned o = pos[s], s1 = partner[s];
    int a, b, c;
    tet_face_verts(tet, (int)s, a, b, c);
    face_fx3[(size_t)o * 3] = a; face_fx3[(size_t)o * 3 + 1] = b; face_fx3[(size_t)o * 3 + 2] = c;
    bool two = s1 != 0xffffffffu;
    if (face_tet_fx2) { face_tet_fx2[(size_t)o * 2] = (int)(s >> 2); face_tet_fx2[(size_t)o * 2 + 1] = two ? (int)(s1 >> 2) : -1; }
    if (face_slot_fx2) { face_slot_fx2[(size_t)o * 2] = (int)(s & 3); face_slot_fx2[(size_t)o * 2 + 1] = two ? (int)(s1 & 3) : -1; }
}
// nbr[4t+i] = the tet across local face i of tet t, -1 on the boundary (pre-filled by the caller)
__global__ void __launch_bounds__(256) neighbour_kernel(const unsigned* __restrict__ run_len, const unsigned* __restrict__ sorted_slot, size_t n,
                                                        int32_t* __restrict__ nbr) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || run_len[p] != 2) return;
    unsigned s0 = sorted_slot[p], s1 = sorted_slot[p + 1];
    nbr[s0] = (int)(s1 >> 2);
    nbr[s1] = (int)(s0 >> 2);
}

static int key_bits_for(u64 max_key) { int b = 1; while (b < 64 && (max_key >> b)) ++b; return b; }

}  // namespace dtb

using namespace dtb;

// ====================================================================================================
// device-pointer API
// ====================================================================================================
struct FaceSort { u64* keys; unsigned* slot; unsigned* run_len; size_t n; };

static size_t face_sort_ws(size_t n) {
    Workspace ws(nullptr, 0);
    ws.take<u64>(n); ws.take<u64>(n); ws.take<unsigned>(n); ws.take<unsigned>(n); ws.take<unsigned>(n);
    ws.take<char>(sort_workspace_bytes(n));
    return ws.off;
}
// sorts the 4T tet-face keys; returns sorted keys / slots / run lengths carved from ws
static int face_sort(const int32_t* tet, int n_point, int T, Workspace& ws, FaceSort& fs, cudaStream_t st) {
    size_t n = (size_t)T * 4;
    u64* k0 = ws.take<u64>(n); u64* k1 = ws.take<u64>(n);
    unsigned* v0 = ws.take<unsigned>(n); unsigned* v1 = ws.take<unsigned>(n);
    unsigned* rl = ws.take<unsigned>(n);
    size_t sb = sort_workspace_bytes(n);
    void* sws = ws.take<char>(sb);
    if (!ws.ok) { set_error("builders: workspace too small (need >= %zu)", ws.off); return DTB_EWORKSPACE; }
    u64 nv = (u64)n_point;
    if (nv >= (1ull << 21)) { set_error("builders: n_point %d too large for a 63-bit face key", n_point); return DTB_EOVERFLOW; }
    face_keys_kernel<<<cdiv((long long)n, 256), 256, 0, st>>>(tet, T, nv, k0, v0);
    DTB_LAUNCH_CHECK("face_keys");
    int bits = key_bits_for(nv * nv * nv);
    int rc = radix_sort_pairs_u64(k0, v0, k1, v1, n, bits, sws, sb, st);
    if (rc) return rc;
    run_lengths_kernel<<<cdiv((long long)n, 256), 256, 0, st>>>(k1, n, rl);
    DTB_LAUNCH_CHECK("run_lengths");
    fs.keys = k1; fs.slot = v1; fs.run_len = rl; fs.n = n;
    return DTB_OK;
}

extern "C" size_t dtb_tet_to_face_workspace(int T) {
    size_t n = (size_t)T * 4;
    return face_sort_ws(n) + 5 * align_up(n * 4, 256) + scan_workspace_bytes(n) + 1024;
}
// counts[0] = number of interior (2-tet) faces, counts[1] = number of boundary (1-tet) faces.
extern "C" int dtb_tet_to_face(const int32_t* tet, int n_point, int T, int32_t* face_fx3, int32_t* face_tet_fx2, int32_t* face_slot_fx2,
                               int32_t* boundary_fx3, int32_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(tet && counts, "tet_to_face: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) { DTB_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), st)); return DTB_OK; }
    Workspace ws(workspace, workspace_bytes);
    FaceSort fs;
    int rc = face_sort(tet, n_point, T, ws, fs, st);
    if (rc) return rc;
    size_t n = fs.n;
    unsigned* pair_flag = ws.take<unsigned>(n); unsigned* partner = ws.take<unsigned>(n); unsigned* single_flag = ws.take<unsigned>(n);
    unsigned* pair_pos = ws.take<unsigned>(n); unsigned* single_pos = ws.take<unsigned>(n);
    size_t sb = scan_workspace_bytes(n);
    void* sws = ws.take<char>(sb);
    if (!ws.ok || !workspace) { set_error("tet_to_face: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    DTB_CUDA(cudaMemsetAsync(pair_flag, 0, n * 4, st));
    DTB_CUDA(cudaMemsetAsync(single_flag, 0, n * 4, st));
    int blocks = cdiv((long long)n, 256);
    face_mark_kernel<<<blocks, 256, 0, st>>>(fs.run_len, fs.slot, n, pair_flag, partner, single_flag);
    DTB_LAUNCH_CHECK("face_mark");
    rc = exclusive_scan_u32(pair_flag, pair_pos, n, (unsigned*)counts, sws, sb, st);
    if (rc) return rc;
    rc = exclusive_scan_u32(single_flag, single_pos, n, (unsigned*)counts + 1, sws, sb, st);
    if (rc) return rc;
    face_emit_kernel<<<blocks, 256, 0, st>>>(tet, n, pair_flag, pair_pos, partner, single_flag, single_pos, face_fx3, face_tet_fx2,
                                             face_slot_fx2, boundary_fx3);
    DTB_LAUNCH_CHECK("face_emit");
    return DTB_OK;
}

extern "C" size_t dtb_tet_to_face_idx_workspace(int T) {
    size_t n = (size_t)T * 4;
    return face_sort_ws(n) + 3 * align_up(n * 4, 256) + scan_workspace_bytes(n) + 1024;
}
// Interior AND boundary faces in one list, in first-occurrence order; boundary faces carry -1 in the second column of
// face_tet_fx2 / face_slot_fx2.  *n_face = number of rows written (capacity needed: 4T rows).
extern "C" int dtb_tet_to_face_idx(const int32_t* tet, int n_point, int T, int32_t* face_fx3, int32_t* face_tet_fx2, int32_t* face_slot_fx2,
                                   int32_t* n_face, void* workspace, size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(tet && face_fx3 && n_face, "tet_to_face_idx: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) { DTB_CUDA(cudaMemsetAsync(n_face, 0, sizeof(int32_t), st)); return DTB_OK; }
    if (!workspace) { set_error("tet_to_face_idx: null workspace"); return DTB_EWORKSPACE; }
    Workspace ws(workspace, workspace_bytes);
    FaceSort fs;
    int rc = face_sort(tet, n_point, T, ws, fs, st);
    if (rc) return rc;
    size_t n = fs.n;
    unsigned* flag = ws.take<unsigned>(n); unsigned* partner = ws.take<unsigned>(n); unsigned* pos = ws.take<unsigned>(n);
    size_t sb = scan_workspace_bytes(n);
    void* sws = ws.take<char>(sb);
    if (!ws.ok || !workspace) { set_error("tet_to_face_idx: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    DTB_CUDA(cudaMemsetAsync(flag, 0, n * 4, st));
    int blocks = cdiv((long long)n, 256);
    face_mark_all_kernel<<<blocks, 256, 0, st>>>(fs.run_len, fs.slot, n, flag, partner);
    DTB_LAUNCH_CHECK("face_mark_all");
    rc = exclusive_scan_u32(flag, pos, n, (unsigned*)n_face, sws, sb, st);
    if (rc) return rc;
    face_emit_all_kernel<<<blocks, 256, 0, st>>>(tet, n, flag, pos, partner, face_fx3, face_tet_fx2, face_slot_fx2);
    DTB_LAUNCH_CHECK("face_emit_all");
    return DTB_OK;
}

extern "C" size_t dtb_tet_neighbours_workspace(int T) { return face_sort_ws((size_t)T * 4) + 1024; }
extern "C" int dtb_tet_neighbours(const int32_t* tet, int n_point, int T, int32_t* neighbour_tx4, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    DTB_REQUIRE(T == 0 || (tet && neighbour_tx4), "tet_neighbours: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) return DTB_OK;
    if (!workspace) { set_error("tet_neighbours: null workspace"); return DTB_EWORKSPACE; }
    Workspace ws(workspace, workspace_bytes);
    FaceSort fs;
    int rc = face_sort(tet, n_point, T, ws, fs, st);
    if (rc) return rc;
    DTB_CUDA(cudaMemsetAsync(neighbour_tx4, 0xff, fs.n * 4, st));
    neighbour_kernel<<<cdiv((long long)fs.n, 256), 256, 0, st>>>(fs.run_len, fs.slot, fs.n, neighbour_tx4);
    DTB_LAUNCH_CHECK("neighbour");
    return DTB_OK;
}

extern "C" size_t dtb_tet_adj_share_workspace(int T) {
    size_t n = (size_t)T * 4;
    return face_sort_ws(n) + 2 * align_up(n * 4, 256) + scan_workspace_bytes(n) + 1024;
}
// out: rows (t0,t1,f0),(t1,t0,f1) -- int32 (2*n_shared, 3); n_shared written to *n_out (the reference's n_face_edge_p[0])
extern "C" int dtb_tet_adj_share(const int32_t* tet, int n_point, int T, int32_t* out, int32_t* n_out, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(tet && out && n_out, "tet_adj_share: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) { DTB_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), st)); return DTB_OK; }
    Workspace ws(workspace, workspace_bytes);
    FaceSort fs;
    int rc = face_sort(tet, n_point, T, ws, fs, st);
    if (rc) return rc;
    size_t n = fs.n;
    unsigned* flag = ws.take<unsigned>(n); unsigned* pos = ws.take<unsigned>(n);
    size_t sb = scan_workspace_bytes(n);
    void* sws = ws.take<char>(sb);
    if (!ws.ok || !workspace) { set_error("tet_adj_share: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    int blocks = cdiv((long long)n, 256);
    share_flag_kernel<<<blocks, 256, 0, st>>>(fs.run_len, n, flag);
    DTB_LAUNCH_CHECK("share_flag");
    rc = exclusive_scan_u32(flag, pos, n, (unsigned*)n_out, sws, sb, st);
    if (rc) return rc;
    share_emit_kernel<<<blocks, 256, 0, st>>>(flag, pos, fs.slot, n, out);
    DTB_LAUNCH_CHECK("share_emit");
    return DTB_OK;
}

extern "C" size_t dtb_tet_point_adj_workspace(int n_point, int T) {
    size_t n = (size_t)T * 12;
    Workspace ws(nullptr, 0);
    ws.take<u64>(n); ws.take<u64>(n); ws.take<unsigned>(n); ws.take<unsigned>(n); ws.take<unsigned>(n); ws.take<unsigned>(n);
    ws.take<unsigned>((size_t)n_point);
    ws.take<char>(sort_workspace_bytes(n)); ws.take<char>(scan_workspace_bytes(n));
    return ws.off + 1024;
}
// edges (<=12T, 2) sorted by (a,b); weight (optional, same length) = 1/deg(a); *n_edge = count
extern "C" int dtb_tet_point_adj(const int32_t* tet, int n_point, int T, int32_t* edges, float* weight, int32_t* n_edge, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(tet && edges && n_edge, "tet_point_adj: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) { DTB_CUDA(cudaMemsetAsync(n_edge, 0, sizeof(int32_t), st)); return DTB_OK; }
    size_t n = (size_t)T * 12;
    Workspace ws(workspace, workspace_bytes);
    u64* k0 = ws.take<u64>(n); u64* k1 = ws.take<u64>(n);
    unsigned* v0 = ws.take<unsigned>(n); unsigned* v1 = ws.take<unsigned>(n);
    unsigned* flag = ws.take<unsigned>(n); unsigned* pos = ws.take<unsigned>(n);
    unsigned* degree = ws.take<unsigned>((size_t)n_point);
    size_t sob = sort_workspace_bytes(n), scb = scan_workspace_bytes(n);
    void* sows = ws.take<char>(sob); void* scws = ws.take<char>(scb);
    if (!ws.ok || !workspace) { set_error("tet_point_adj: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    u64 nv = (u64)n_point;
    int blocks = cdiv((long long)n, 256);
    edge_keys_kernel<<<blocks, 256, 0, st>>>(tet, T, nv, k0, v0);
    DTB_LAUNCH_CHECK("edge_keys");
    int rc = radix_sort_pairs_u64(k0, v0, k1, v1, n, key_bits_for(nv * nv), sows, sob, st);
    if (rc) return rc;
    unique_flag_kernel<<<blocks, 256, 0, st>>>(k1, n, flag);
    DTB_LAUNCH_CHECK("unique_flag");
    rc = exclusive_scan_u32(flag, pos, n, (unsigned*)n_edge, scws, scb, st);
    if (rc) return rc;
    DTB_CUDA(cudaMemsetAsync(degree, 0, (size_t)n_point * 4, st));
    edge_emit_kernel<<<blocks, 256, 0, st>>>(k1, flag, pos, n, nv, edges, degree);
    DTB_LAUNCH_CHECK("edge_emit");
    if (weight) {
        edge_weight_kernel<<<blocks, 256, 0, st>>>(edges, (const unsigned*)n_edge, degree, weight, n);
        DTB_LAUNCH_CHECK("edge_weight");
    }
    return DTB_OK;
}

extern "C" size_t dtb_tet_face_adj_workspace(int T) {
    size_t n = (size_t)T * 12;
    Workspace ws(nullptr, 0);
    ws.take<u64>(n); ws.take<u64>(n); ws.take<unsigned>(n); ws.take<unsigned>(n); ws.take<unsigned>(n); ws.take<unsigned>(n);
    ws.take<char>(sort_workspace_bytes(n)); ws.take<char>(scan_workspace_bytes(n));
    return ws.off + 1024;
}
// pairs (capacity rows, 2) of tet-face ids 4*t+i; *n_pairs = total number (may exceed capacity: then pairs is truncated and
// DTB_EOVERFLOW is NOT raised -- the caller compares n_pairs with capacity after synchronising).
extern "C" int dtb_tet_face_adj(const int32_t* tet, int n_point, int T, int32_t* pairs, long long capacity, int32_t* n_pairs,
                                void* workspace, size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(tet && pairs && n_pairs, "tet_face_adj: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (T == 0) { DTB_CUDA(cudaMemsetAsync(n_pairs, 0, sizeof(int32_t), st)); return DTB_OK; }
    size_t n = (size_t)T * 12;
    Workspace ws(workspace, workspace_bytes);
    u64* k0 = ws.take<u64>(n); u64* k1 = ws.take<u64>(n);
    unsigned* v0 = ws.take<unsigned>(n); unsigned* v1 = ws.take<unsigned>(n);
    unsigned* cnt = ws.take<unsigned>(n); unsigned* pos = ws.take<unsigned>(n);
    size_t sob = sort_workspace_bytes(n), scb = scan_workspace_bytes(n);
    void* sows = ws.take<char>(sob); void* scws = ws.take<char>(scb);
    if (!ws.ok || !workspace) { set_error("tet_face_adj: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    int blocks = cdiv((long long)n, 256);
    fedge_keys_kernel<<<blocks, 256, 0, st>>>(tet, T, (unsigned)n_point, k0, v0);
    DTB_LAUNCH_CHECK("fedge_keys");
    int rc = radix_sort_pairs_u64(k0, v0, k1, v1, n, 32, sows, sob, st);
    if (rc) return rc;
    fadj_count_kernel<<<blocks, 256, 0, st>>>(tet, k1, v1, n, (u64)n_point, cnt, nullptr, nullptr, 0);
    DTB_LAUNCH_CHECK("fadj_count");
    rc = exclusive_scan_u32(cnt, pos, n, (unsigned*)n_pairs, scws, scb, st);
    if (rc) return rc;
    fadj_count_kernel<<<blocks, 256, 0, st>>>(tet, k1, v1, n, (u64)n_point, nullptr, pairs, pos, (size_t)capacity);
    DTB_LAUNCH_CHECK("fadj_emit");
    return DTB_OK;
}

static unsigned pow2_at_least(size_t x) { unsigned p = 64; while (p < x) p <<= 1; return p; }
extern "C" size_t dtb_collapse_vertices_workspace(int N) {
    Workspace ws(nullptr, 0);
    ws.take<int>(pow2_at_least((size_t)N * 2)); ws.take<int32_t>(N); ws.take<unsigned>(N); ws.take<unsigned>(N);
    ws.take<char>(scan_workspace_bytes(N));
    return ws.off + 1024;
}
// map_array (N,), inverse_idx (N,) (first *n_unique entries valid)
extern "C" int dtb_collapse_vertices(const float* points, int N, int32_t* map_array, int32_t* inverse_idx, int32_t* n_unique,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    DTB_REQUIRE(points && map_array && inverse_idx && n_unique, "collapse_vertices: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) { DTB_CUDA(cudaMemsetAsync(n_unique, 0, sizeof(int32_t), st)); return DTB_OK; }
    unsigned H = pow2_at_least((size_t)N * 2);
    Workspace ws(workspace, workspace_bytes);
    int* slots = ws.take<int>(H);
    int32_t* rep = ws.take<int32_t>(N);
    unsigned* is_first = ws.take<unsigned>(N); unsigned* uid = ws.take<unsigned>(N);
    size_t sb = scan_workspace_bytes(N);
    void* sws = ws.take<char>(sb);
    if (!ws.ok || !workspace) { set_error("collapse_vertices: workspace too small (%zu < %zu)", workspace_bytes, ws.off); return DTB_EWORKSPACE; }
    DTB_CUDA(cudaMemsetAsync(slots, 0xff, (size_t)H * 4, st));
    int blocks = cdiv(N, 256);
    collapse_insert_kernel<<<blocks, 256, 0, st>>>(points, N, H, slots);
    DTB_LAUNCH_CHECK("collapse_insert");
    collapse_lookup_kernel<<<blocks, 256, 0, st>>>(points, N, H, slots, rep, is_first);
    DTB_LAUNCH_CHECK("collapse_lookup");
    int rc = exclusive_scan_u32(is_first, uid, N, (unsigned*)n_unique, sws, sb, st);
    if (rc) return rc;
    collapse_emit_kernel<<<blocks, 256, 0, st>>>(rep, is_first, uid, N, map_array, inverse_idx);
    DTB_LAUNCH_CHECK("collapse_emit");
    return DTB_OK;
}

// ====================================================================================================
// host-pointer API: exact argument lists of the reference's ctypes `run(...)` builders (blocking)
// ====================================================================================================
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { return dtb::check_cuda(cudaMalloc(&p, bytes ? bytes : 16), "cudaMalloc"); }
};
#define HOST_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
}

// utils/lib/tet_point_adj/run.cpp:20  run(int* tet_list, int* edge_p, int* n_edge, int n_point, int n_tet)
extern "C" int dtb_host_tet_point_adj(const int32_t* tet_list, int32_t* edge_p, int32_t* n_edge, int n_point, int n_tet) {
    DevBuf dt, de, dn, dw;
    size_t wsz = dtb_tet_point_adj_workspace(n_point, n_tet);
    HOST_TRY(dt.alloc((size_t)n_tet * 16)); HOST_TRY(de.alloc((size_t)n_tet * 12 * 8)); HOST_TRY(dn.alloc(16)); HOST_TRY(dw.alloc(wsz));
    DTB_CUDA(cudaMemcpy(dt.p, tet_list, (size_t)n_tet * 16, cudaMemcpyHostToDevice));
    HOST_TRY(dtb_tet_point_adj((const int32_t*)dt.p, n_point, n_tet, (int32_t*)de.p, nullptr, (int32_t*)dn.p, dw.p, wsz, nullptr));
    DTB_CUDA(cudaMemcpy(n_edge, dn.p, 4, cudaMemcpyDeviceToHost));
    DTB_CUDA(cudaMemcpy(edge_p, de.p, (size_t)n_edge[0] * 8, cudaMemcpyDeviceToHost));
    return DTB_OK;
}
// utils/lib/tet_adj_share/run.cpp:40  run(int* tet_list, int* face_edge_p, int* n_face_edge_p, int n_point, int n_tet)
extern "C" int dtb_host_tet_adj_share(const int32_t* tet_list, int32_t* face_edge_p, int32_t* n_face_edge_p, int n_point, int n_tet) {
    DevBuf dt, de, dn, dw;
    size_t wsz = dtb_tet_adj_share_workspace(n_tet);
    HOST_TRY(dt.alloc((size_t)n_tet * 16)); HOST_TRY(de.alloc((size_t)n_tet * 8 * 12)); HOST_TRY(dn.alloc(16)); HOST_TRY(dw.alloc(wsz));
    DTB_CUDA(cudaMemcpy(dt.p, tet_list, (size_t)n_tet * 16, cudaMemcpyHostToDevice));
    HOST_TRY(dtb_tet_adj_share((const int32_t*)dt.p, n_point, n_tet, (int32_t*)de.p, (int32_t*)dn.p, dw.p, wsz, nullptr));
    DTB_CUDA(cudaMemcpy(n_face_edge_p, dn.p, 4, cudaMemcpyDeviceToHost));
    DTB_CUDA(cudaMemcpy(face_edge_p, de.p, (size_t)n_face_edge_p[0] * 24, cudaMemcpyDeviceToHost));
    return DTB_OK;
}
// utils/lib/tet_face_adj/run.cpp:18  run(int* tet_list, int* face_edge_p, int* n_face_edge_p, int n_point, int n_tet)
// the caller's buffer holds 4*n_tet*50 rows (tet_face_adj/interface.py:27-28)
extern "C" int dtb_host_tet_face_adj(const int32_t* tet_list, int32_t* face_edge_p, int32_t* n_face_edge_p, int n_point, int n_tet) {
    DevBuf dt, de, dn, dw;
    size_t wsz = dtb_tet_face_adj_workspace(n_tet);
    long long cap = (long long)n_tet * 200;
    HOST_TRY(dt.alloc((size_t)n_tet * 16)); HOST_TRY(de.alloc((size_t)cap * 8)); HOST_TRY(dn.alloc(16)); HOST_TRY(dw.alloc(wsz));
    DTB_CUDA(cudaMemcpy(dt.p, tet_list, (size_t)n_tet * 16, cudaMemcpyHostToDevice));
    HOST_TRY(dtb_tet_face_adj((const int32_t*)dt.p, n_point, n_tet, (int32_t*)de.p, cap, (int32_t*)dn.p, dw.p, wsz, nullptr));
    DTB_CUDA(cudaMemcpy(n_face_edge_p, dn.p, 4, cudaMemcpyDeviceToHost));
    long long n = n_face_edge_p[0];
    if (n > cap) { set_error("tet_face_adj: %lld pairs exceed the caller's 200*n_tet rows", n); return DTB_EOVERFLOW; }
    DTB_CUDA(cudaMemcpy(face_edge_p, de.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return DTB_OK;
}
// utils/lib/colaps_v/run.cpp:39  run(float* point_p, int* map_array_p, int* inverse_idx_p, int* n_colaps_v_p, int n_point)
extern "C" int dtb_host_colaps_v(const float* point_p, int32_t* map_array_p, int32_t* inverse_idx_p, int32_t* n_colaps_v_p, int n_point) {
    DevBuf dp, dm, di, dn, dw;
    size_t wsz = dtb_collapse_vertices_workspace(n_point);
    HOST_TRY(dp.alloc((size_t)n_point * 12)); HOST_TRY(dm.alloc((size_t)n_point * 4)); HOST_TRY(di.alloc((size_t)n_point * 4));
    HOST_TRY(dn.alloc(16)); HOST_TRY(dw.alloc(wsz));
    DTB_CUDA(cudaMemcpy(dp.p, point_p, (size_t)n_point * 12, cudaMemcpyHostToDevice));
    HOST_TRY(dtb_collapse_vertices((const float*)dp.p, n_point, (int32_t*)dm.p, (int32_t*)di.p, (int32_t*)dn.p, dw.p, wsz, nullptr));
    DTB_CUDA(cudaMemcpy(n_colaps_v_p, dn.p, 4, cudaMemcpyDeviceToHost));
    DTB_CUDA(cudaMemcpy(map_array_p, dm.p, (size_t)n_point * 4, cudaMemcpyDeviceToHost));
    DTB_CUDA(cudaMemcpy(inverse_idx_p, di.p, (size_t)n_colaps_v_p[0] * 4, cudaMemcpyDeviceToHost));
    return DTB_OK;
}
