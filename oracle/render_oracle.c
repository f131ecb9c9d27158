/* CPU oracle for the two Kaolin operations on the DefTet hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: kaolin (NVIDIAGameWorks/kaolin) is neither vendored nor pinned by the reference
 * (README.md:30 "Install Kaolin following official Link") and is not installed here, and the reference has no
 * test or golden vector at these two boundaries.  What follows restates the call-site contract
 *   kal.render.mesh.deftet_sparse_render  diff_render/diftet_6_subdiv/5_rendereq/deftetrneder.py:97-100
 *   kal.ops.mesh.check_sign               layers/DefTet/deftet.py:46
 * plus the published semantics (paper docs/files/main.pdf section 3.2.2), and is the parity
 * definition for deftet_b200's kernels until a Kaolin build can be diffed against it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* per pixel: faces in ascending id whose 2-D triangle contains the pixel (w1=k1/(k3+eps), w2=k2/(k3+eps), w0=1-w1-w2,
 * all >= 0, pixel inside the face's bounding box) and whose interpolated z is inside [zmin, zmax]; first K of them;
 * sorted by z descending (ties: ascending face id); interpolated features; void slots: idx -1, features 0. */
void orc_sparse_render(const float* pix, const float* ranges, const float* face_z, const float* face_xy, const float* face_feat, int B, int P,
                       int F, int D, int K, float eps, float* out_feat, long long* out_idx, long long p0, long long p1) {
    float* depth = (float*)malloc(sizeof(float) * (K > 0 ? K : 1));
    float* w1s = (float*)malloc(sizeof(float) * (K > 0 ? K : 1));
    float* w2s = (float*)malloc(sizeof(float) * (K > 0 ? K : 1));
    int* fid = (int*)malloc(sizeof(int) * (K > 0 ? K : 1));
    (void)B;
    for (long long i = p0; i < p1; ++i) {
        int b = (int)(i / P);
        float px = pix[i * 2], py = pix[i * 2 + 1], zmin = ranges[i * 2], zmax = ranges[i * 2 + 1];
        int cnt = 0;
        for (int f = 0; f < F && cnt < K; ++f) {
            const float* xy = face_xy + ((size_t)b * F + f) * 6;
            const float* z = face_z + ((size_t)b * F + f) * 3;
            float ax = xy[0], ay = xy[1], bx = xy[2], by = xy[3], cx = xy[4], cy = xy[5];
            float xmin = fminf(ax, fminf(bx, cx)), xmax = fmaxf(ax, fmaxf(bx, cx)), ymin = fminf(ay, fminf(by, cy)), ymax = fmaxf(ay, fmaxf(by, cy));
            if (px < xmin || px > xmax || py < ymin || py > ymax) continue;
            float m = bx - ax, pp = by - ay, n = cx - ax, q = cy - ay, s = px - ax, t = py - ay;
            float k1 = s * q - n * t, k2 = m * t - s * pp, k3 = m * q - n * pp;
            float den = k3 + eps;
            float w1 = k1 / den, w2 = k2 / den, w0 = 1.0f - w1 - w2;
            if (w0 < 0.f || w1 < 0.f || w2 < 0.f) continue;
            float d = w0 * z[0] + w1 * z[1] + w2 * z[2];
            if (!(d >= zmin && d <= zmax)) continue;
            depth[cnt] = d; w1s[cnt] = w1; w2s[cnt] = w2; fid[cnt] = f; ++cnt;
        }
        /* insertion sort: depth descending, face id ascending on ties (stable since hits arrive in ascending id) */
        for (int a = 1; a < cnt; ++a) {
            float d = depth[a], u1 = w1s[a], u2 = w2s[a]; int f = fid[a]; int c = a - 1;
            while (c >= 0 && depth[c] < d) { depth[c + 1] = depth[c]; w1s[c + 1] = w1s[c]; w2s[c + 1] = w2s[c]; fid[c + 1] = fid[c]; --c; }
            depth[c + 1] = d; w1s[c + 1] = u1; w2s[c + 1] = u2; fid[c + 1] = f;
        }
        for (int k = 0; k < K; ++k) {
            out_idx[i * K + k] = k < cnt ? fid[k] : -1;
            for (int c = 0; c < D; ++c) {
                float v = 0.f;
                if (k < cnt) {
                    const float* ff = face_feat + ((size_t)b * F + fid[k]) * 3 * D;
                    float w0 = 1.0f - w1s[k] - w2s[k];
                    v = w0 * ff[c] + w1s[k] * ff[D + c] + w2s[k] * ff[2 * D + c];
                }
                out_feat[(i * K + k) * D + c] = v;
            }
        }
    }
    free(depth); free(w1s); free(w2s); free(fid);
}

/* +z ray parity against a closed triangle mesh, double precision, half-open edge rule (an edge on the ray belongs to
 * the triangle for which the counter-clockwise-oriented edge runs downwards, or leftwards when horizontal). */
static int edge_inside(double ax, double ay, double bx, double by, double px, double py, int ccw) {
    double e = (bx - ax) * (py - ay) - (by - ay) * (px - ax);
    if (!ccw) e = -e;
    if (e > 0.0) return 1;
    if (e < 0.0) return 0;
    double dx = bx - ax, dy = by - ay;
    if (!ccw) { dx = -dx; dy = -dy; }
    return (dy < 0.0) || (dy == 0.0 && dx < 0.0);
}
void orc_check_sign(const float* verts, const int32_t* faces, const float* points, int B, int n, int m, int p, unsigned char* out, long long i0,
                    long long i1) {
    (void)B;
    for (long long i = i0; i < i1; ++i) {
        int b = (int)(i / p);
        const float* vb = verts + (size_t)b * n * 3;
        double px = points[i * 3], py = points[i * 3 + 1], pz = points[i * 3 + 2];
        unsigned crossings = 0;
        for (int f = 0; f < m; ++f) {
            const float *A = vb + (size_t)faces[f * 3] * 3, *Bv = vb + (size_t)faces[f * 3 + 1] * 3, *Cv = vb + (size_t)faces[f * 3 + 2] * 3;
            double ax = A[0], ay = A[1], bx = Bv[0], by = Bv[1], cx = Cv[0], cy = Cv[1];
            double area = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
            if (area == 0.0) continue;
            int ccw = area > 0.0;
            if (!edge_inside(ax, ay, bx, by, px, py, ccw) || !edge_inside(bx, by, cx, cy, px, py, ccw) || !edge_inside(cx, cy, ax, ay, px, py, ccw)) continue;
            double w0 = ((bx - px) * (cy - py) - (by - py) * (cx - px)) / area;
            double w1 = ((cx - px) * (ay - py) - (cy - py) * (ax - px)) / area;
            double z = w0 * (double)A[2] + w1 * (double)Bv[2] + (1.0 - w0 - w1) * (double)Cv[2];
            if (z > pz) ++crossings;
        }
        out[i] = (unsigned char)(crossings & 1u);
    }
}
