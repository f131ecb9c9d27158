/* CPU oracle for the brute-force search kernels of the DefTet hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference has no CPU implementation of these (they exist only as CUDA kernels), so this is a
 * restatement ("port") of the cited kernel bodies in plain C, one IEEE-754 fp32 rounding per source
 * operator (build with -ffp-contract=off; never -ffast-math).  Every function takes an [i0, i1) range of its outer (per point / per face) loop so that
 * the Python wrapper can spread it over host threads (ctypes releases the GIL; libgomp is not in this
 * image); results do not depend on the split.  Nothing in deftet_b200/ links or loads this file.
 * PINNED: checked bit for bit against the reference's own __host__ __device__ functions compiled for the host
 * (oracle/build_ref_kernels.sh -> oracle/_ref/kernels/*.so; tests/test_golden.py), except the trivial A2 body.
 *
 *   orc_point_in_tet            layers/DefTet/check_condition_tetrahedron_base/check_condition_tet_for.cu:105-189
 *   orc_nearest_neighbor        layers/nearest_neighbor/nearest_neighbor_cuda.cu:17-55
 *   orc_point_face_distance     layers/DefTet/tet_analytic_distance_batch/tet_analytic_distance_for.cu:139-307
 *   orc_point_face_distance_bwd layers/DefTet/tet_analytic_distance_batch/tet_analytic_distance_back.cu:222-317,348-483,592-686
 *   orc_face_adjacency          layers/DefTet/tet_face_adj_m_idx/tet_face_adj_m_for.cu:15-108
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EPS_D 1e-10
#define FWD_MAX_DIS 10000.0f   /* tet_analytic_distance_for.cu:17 */
#define BWD_MAX_DIS 9999999.0f /* tet_analytic_distance_back.cu:19 */

/* ---------------------------------------------------------------------------------------------- A1 */
static void v_minus(const float* a, const float* b, float* r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static void v_cross(const float* a, const float* b, float* n) {
    n[0] = a[1] * b[2] - a[2] * b[1];
    n[1] = a[2] * b[0] - a[0] * b[2];
    n[2] = a[0] * b[1] - a[1] * b[0];
}
static float v_dot(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* check_condition_tet_for.cu:105-121 */
static int same_side(const float* a, const float* b, const float* c, const float* d, const float* p) {
    float r1[3], r2[3], n[3];
    v_minus(b, a, r1);
    v_minus(c, a, r2);
    v_cross(r1, r2, n);
    v_minus(d, a, r1);
    float dotv4 = v_dot(n, r1);
    v_minus(p, a, r1);
    float dotp = v_dot(n, r1);
    return (dotp > 0) == (dotv4 > 0);
}

/* check_condition_tet_for.cu:124-189: first tet (ascending id) whose four predicates agree, else -1. */
void orc_point_in_tet(const float* tet_bxfx4x3, const float* pts_bxnx3, float* cond_bxn, int B, int P, int T, long long i0,
                      long long i1) {
    (void)B;
    for (long long i = i0; i < i1; ++i) {
        int b = (int)(i / P);
        const float* p = pts_bxnx3 + i * 3;
        const float* tb = tet_bxfx4x3 + (size_t)b * T * 12;
        float hit = -1.0f;
        for (int t = 0; t < T; ++t) {
            const float *A = tb + (size_t)t * 12, *Bv = A + 3, *Cv = A + 6, *D = A + 9;
            int s1 = same_side(A, Bv, Cv, D, p);
            int s2 = same_side(Bv, A, D, Cv, p);
            int s3 = same_side(Cv, D, A, Bv, p);
            int s4 = same_side(D, Cv, Bv, A, p);
            if (s1 == s2 && s2 == s3 && s3 == s4) { hit = (float)t; break; }
        }
        cond_bxn[i] = hit;
    }
}

/* ---------------------------------------------------------------------------------------------- A2 */
/* nearest_neighbor_cuda.cu:17-55: argmin of squared distance, strict <, init 1e20 / index 0. */
void orc_nearest_neighbor(const float* queries, const float* points, int32_t* result, int B, int Q, int M, long long i0,
                          long long i1) {
    (void)B;
    for (long long i = i0; i < i1; ++i) {
        int b = (int)(i / Q);
        const float* bp = points + (size_t)b * M * 3;
        float qx = queries[i * 3], qy = queries[i * 3 + 1], qz = queries[i * 3 + 2];
        float best = 1e20f;
        int bi = 0;
        for (int j = 0; j < M; ++j) {
            float d = 0;
            float e = bp[j * 3] - qx;
            d += e * e;
            e = bp[j * 3 + 1] - qy;
            d += e * e;
            e = bp[j * 3 + 2] - qz;
            d += e * e;
            if (d < best) { bi = j; best = d; }
        }
        result[i] = bi;
    }
}

/* ---------------------------------------------------------------------------------------------- A4 */
/* cuda_divide_non_zero: `a + eps` is evaluated in double (eps is a double literal) and narrowed on return. */
static float div_nz(float a) {
    if (a == 0) return (float)EPS_D;
    if (a < 0) return (float)((double)a - EPS_D);
    if (a > 0) return (float)((double)a + EPS_D);
    return (float)EPS_D;
}
static float f_abs(float a) { return a > 0.0 ? a : -a; }
static float min3(float a, float b, float c) { float m = a; if (b < m) m = b; if (c < m) m = c; return m; }
static float min3_idx(float a, float b, float c) { float m = a, i = 0; if (b < m) { m = b; i = 1; } if (c < m) { m = c; i = 2; } return i; }
static void v_normalize(float* a) {
    float len = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    len = div_nz(len);
    a[0] = a[0] / len; a[1] = a[1] / len; a[2] = a[2] / len;
}
static float pt_dist2(const float* a, const float* b) {
    return (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
}
/* distance_line_square: signed -- negative when the foot point is outside the segment */
static float line_dist2(const float* A, const float* Bp, const float* P) {
    float PA[3], BA[3], d[3], tmp[3];
    v_minus(P, A, PA);
    v_minus(Bp, A, BA);
    float t = v_dot(PA, BA) / div_nz(v_dot(BA, BA));
    tmp[0] = BA[0] * t; tmp[1] = BA[1] * t; tmp[2] = BA[2] * t;
    v_minus(PA, tmp, d);
    float dist = v_dot(d, d);
    if (t >= 0 && t <= 1) return dist;
    return -dist;
}
/* cuda_line_distance (xy-projected inside test): ret[0] type (-1 degenerate, 0 inside, 1 edge, 2 vertex),
 * ret[1] in-plane distance^2, ret[2] index of the closest edge / vertex (backward variant). */
static void tri_classify(const float* a, const float* b, const float* c, const float* p, float max_dis, float* ret) {
    float k1 = (b[1] - c[1]) * (p[0] - c[0]) + (c[0] - b[0]) * (p[1] - c[1]);
    float k2 = (a[0] - c[0]) * (p[1] - c[1]) + (c[1] - a[1]) * (p[0] - c[0]);
    float k3 = (b[1] - c[1]) * (a[0] - c[0]) + (c[0] - b[0]) * (a[1] - c[1]);
    if (k3 == 0) { ret[0] = -1; return; }
    float l1 = k1 / k3, l2 = k2 / k3, l3 = 1 - l1 - l2;
    float d12 = line_dist2(a, b, p), d23 = line_dist2(b, c, p), d13 = line_dist2(a, c, p);
    if (l1 >= 0 && l2 >= 0 && l3 >= 0) {
        ret[0] = 0;
        ret[1] = min3(f_abs(d12), f_abs(d23), f_abs(d13));
        ret[2] = min3_idx(f_abs(d12), f_abs(d23), f_abs(d13));
        return;
    }
    if (d12 <= 0) d12 = max_dis;
    if (d23 <= 0) d23 = max_dis;
    if (d13 <= 0) d13 = max_dis;
    float ml = min3(d12, d23, d13), mli = min3_idx(d12, d23, d13);
    float e1 = pt_dist2(a, p), e2 = pt_dist2(b, p), e3 = pt_dist2(c, p);
    float mp = min3(e1, e2, e3), mpi = min3_idx(e1, e2, e3);
    if (ml < mp) { ret[0] = 1; ret[1] = ml; ret[2] = mli; }
    else { ret[0] = 2; ret[1] = mp; ret[2] = mpi; }
}
/* cuda_min_triangle_distance */
static float tri_dist(const float* a, const float* b, const float* c, const float* p, float max_dis, float* ret, float* ip) {
    float r1[3], r2[3], n[3];
    v_minus(b, a, r1);
    v_minus(c, a, r2);
    v_cross(r1, r2, n);
    v_normalize(n);
    float t = v_dot(n, a) - v_dot(n, p);
    r1[0] = n[0] * t; r1[1] = n[1] * t; r1[2] = n[2] * t;
    ip[0] = p[0] + r1[0]; ip[1] = p[1] + r1[1]; ip[2] = p[2] + r1[2];
    float d1 = t * t;
    ret[0] = ret[1] = ret[2] = 0;
    tri_classify(a, b, c, ip, max_dis, ret);
    if (ret[0] == 0) return d1;
    if (ret[0] < 0) return max_dis;
    return d1 + ret[1];
}

/* tet_analytic_distance_for.cu:257-307: min over the first n_face_b[b] faces, first strict minimum wins. */
void orc_point_face_distance(const float* pts_bxpx3, const float* face_bxfx3x3, const float* n_face_b, float* closest_d,
                             float* closest_f, int B, int P, int F, long long i0, long long i1) {
    (void)B;
    for (long long i = i0; i < i1; ++i) {
        int b = (int)(i / P);
        const float* p = pts_bxpx3 + i * 3;
        const float* fb = face_bxfx3x3 + (size_t)b * F * 9;
        float min_d = 10000.0f;
        int min_i = -1;
        int nf = (int)n_face_b[b];
        for (int f = 0; f < nf; ++f) {
            float ret[3], ip[3];
            float d = tri_dist(fb + (size_t)f * 9, fb + (size_t)f * 9 + 3, fb + (size_t)f * 9 + 6, p, FWD_MAX_DIS, ret, ip);
            if (min_d > d) { min_d = d; min_i = f; }
        }
        closest_d[i] = min_d;
        closest_f[i] = (float)min_i;
    }
}

/* tet_analytic_distance_back.cu:592-686.  Serial accumulation in point order (the reference uses float
 * atomics, i.e. an unspecified order); faces with closest_f < 0 are skipped (the reference would read out
 * of bounds).  dldface must be zero-filled by the caller, like utils.py:65. */
void orc_point_face_distance_bwd(const float* pts_bxpx3, const float* face_bxfx3x3, const float* closest_f, const float* dl_dd,
                                 float* dldface, int B, int P, int F) {
    for (long long i = 0; i < (long long)B * P; ++i) {
        int b = (int)(i / P);
        const float* p = pts_bxpx3 + i * 3;
        int fi = (int)closest_f[i];
        if (fi < 0) continue;
        const float* face = face_bxfx3x3 + ((size_t)b * F + fi) * 9;
        float* g = dldface + ((size_t)b * F + fi) * 9;
        float ret[3], ip[3];
        tri_dist(face, face + 3, face + 6, p, BWD_MAX_DIS, ret, ip);
        float gp = dl_dd[i];
        if (ret[0] == 0) { /* cuda_gradient_triangle_distance :440-483 */
            const float *a = face, *bq = face + 3, *c = face + 6;
            float k1 = (bq[1] - c[1]) * (ip[0] - c[0]) + (c[0] - bq[0]) * (ip[1] - c[1]);
            float k2 = (a[0] - c[0]) * (ip[1] - c[1]) + (c[1] - a[1]) * (ip[0] - c[0]);
            float k3 = (bq[1] - c[1]) * (a[0] - c[0]) + (c[0] - bq[0]) * (a[1] - c[1]);
            if (k3 != 0) {
                float l1 = k1 / k3, l2 = k2 / k3, l3 = 1 - l1 - l2;
                float l[3] = {l1, l2, l3};
                for (int v = 0; v < 3; ++v)
                    for (int k = 0; k < 3; ++k) g[v * 3 + k] += gp * (2 * (ip[k] - p[k]) * l[v]);
            }
        }
        if (ret[0] == 1) { /* cuda_gradient_line_distance :291-317: only grad[0..2] survives, with factor t */
            int i1 = (int)ret[2], i2 = (i1 + 1) % 3;
            const float *A = face + i1 * 3, *Bq = face + i2 * 3;
            float PA[3], BA[3];
            v_minus(p, A, PA);
            v_minus(Bq, A, BA);
            float t = v_dot(PA, BA) / div_nz(v_dot(BA, BA));
            float q[3];
            for (int k = 0; k < 3; ++k) q[k] = A[k] * (1 - t) + Bq[k] * t;
            for (int k = 0; k < 3; ++k) g[i1 * 3 + k] += gp * (2 * (q[k] - p[k]) * t);
            (void)i2;
        }
        if (ret[0] == 2) {
            int iv = (int)ret[2];
            for (int k = 0; k < 3; ++k) g[iv * 3 + k] += 2 * gp * ((face[iv * 3 + k] - p[k]) * 1.0f);
        }
    }
}

/* ---------------------------------------------------------------------------------------------- A5 */
static int pt_equal(const float* a, const float* b) { /* tet_face_adj_m_for.cu:26-35, EPS 1e-15 (double) */
    float diff = 0.0f;
    for (int i = 0; i < 3; ++i) { float e = a[i] - b[i]; diff += (e < 0 ? -e : e); }
    return diff <= 1e-15;
}
static int share_edge(const float* fa, const float* fb) { /* :38-69 */
    int find = 0;
    for (int ia = 0; ia < 3; ++ia) {
        const float *aa = fa + ia * 3, *ab = fa + ((ia + 1) % 3) * 3;
        for (int ib = 0; ib < 3; ++ib) {
            const float *ba = fb + ib * 3, *bb = fb + ((ib + 1) % 3) * 3;
            if (pt_equal(aa, ba) && pt_equal(ab, bb)) find = 1;
            if (pt_equal(aa, bb) && pt_equal(ab, ba)) find = 1;
        }
    }
    return find;
}
/* tet_face_adj_m_for.cu:72-108: first n_max_nei neighbours in ascending face id; adj pre-filled with -1. */
void orc_face_adjacency(const float* face_fx3x3, float* adj_fxn, int F, int n_max_nei, long long i0, long long i1) {
    for (int f = (int)i0; f < (int)i1; ++f) {
        int found = 0;
        for (int j = 0; j < F; ++j) {
            if (j == f) continue;
            if (share_edge(face_fx3x3 + (size_t)f * 9, face_fx3x3 + (size_t)j * 9)) {
                adj_fxn[(size_t)f * n_max_nei + found] = (float)j;
                found += 1;
            }
            if (found >= n_max_nei) break;
        }
    }
}
