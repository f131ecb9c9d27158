// Per-tetrahedron energy math shared by energies.cu (direct-gather / soup kernels) and energies_tiled.cu (tile-local kernels).
//   AMIPS            layers/DefTet/deftet.py:266-298
//   edge length      layers/DefTet/deftet.py:320-338
//   volume variance  layers/DefTet/deftet.py:239-263
#pragma once
#include "common.cuh"

namespace dtb {

constexpr int E_TILE = 256;
constexpr int E_MAXB = 32;     // samples handled per launch chunk (smem accumulators)
constexpr float AMIPS_SCALE = 20.0f;
constexpr float AMIPS_EPS = 1e-10f;

struct Tet12 { float a[3], b[3], c[3], d[3]; };

__device__ __forceinline__ void cross3(const float* u, const float* v, float* r) {
    r[0] = u[1] * v[2] - u[2] * v[1];
    r[1] = u[2] * v[0] - u[0] * v[2];
    r[2] = u[0] * v[1] - u[1] * v[0];
}
__device__ __forceinline__ float dot3(const float* u, const float* v) { return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]; }

// ---- per-tet math ---------------------------------------------------------------------------------
// AMIPS: J = 20*[B-A;C-A;D-A] * inv_v, E = ||J||_F^2 (det^2+1e-10)^(-1/3) [det>=0]   (deftet.py:266-298)
__device__ __forceinline__ float amips_energy(const Tet12& t, const float* M, float* J, float& det, float& tr, float& g) {
    float O[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float ak = t.a[k] * AMIPS_SCALE;
        O[0 + k] = t.b[k] * AMIPS_SCALE - ak;
        O[3 + k] = t.c[k] * AMIPS_SCALE - ak;
        O[6 + k] = t.d[k] * AMIPS_SCALE - ak;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) J[i * 3 + j] = O[i * 3 + 0] * M[0 + j] + O[i * 3 + 1] * M[3 + j] + O[i * 3 + 2] * M[6 + j];
    tr = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) tr += J[i] * J[i];
    float bc[3];
    cross3(J + 3, J + 6, bc);
    det = dot3(J, bc);
    g = rcbrtf(det * det + AMIPS_EPS);
    return (det >= 0.f) ? tr * g : 0.f;
}

// gradient of w * E_amips w.r.t. the four vertices
__device__ __forceinline__ void amips_grad(const float* M, const float* J, float det, float tr, float g, float w,
                                           float* ga, float* gb, float* gc, float* gd) {
    if (!(det >= 0.f)) return;
    float cof[9];
    cross3(J + 3, J + 6, cof + 0);
    cross3(J + 6, J + 0, cof + 3);
    cross3(J + 0, J + 3, cof + 6);
    float g2 = g * g;
    float c1 = 2.f * g;
    float c2 = (2.f / 3.f) * tr * det * g2 * g2;
    float dJ[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) dJ[i] = c1 * J[i] - c2 * cof[i];
    float ws = w * AMIPS_SCALE;
    float rows[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k)
            rows[i * 3 + k] = ws * (dJ[i * 3 + 0] * M[k * 3 + 0] + dJ[i * 3 + 1] * M[k * 3 + 1] + dJ[i * 3 + 2] * M[k * 3 + 2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        gb[k] += rows[0 + k];
        gc[k] += rows[3 + k];
        gd[k] += rows[6 + k];
        ga[k] -= rows[0 + k] + rows[3 + k] + rows[6 + k];
    }
}

// Edge energy: sum over 6 edges and xyz of (20*delta)^4                              (deftet.py:320-338)
__device__ __forceinline__ float edge_energy(const Tet12& t) {
    float A[3], B[3], C[3], D[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { A[k] = t.a[k] * 20.f; B[k] = t.b[k] * 20.f; C[k] = t.c[k] * 20.f; D[k] = t.d[k] * 20.f; }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float e;
        e = A[k] - D[k]; e *= e; s += e * e;
        e = B[k] - D[k]; e *= e; s += e * e;
        e = C[k] - D[k]; e *= e; s += e * e;
        e = A[k] - B[k]; e *= e; s += e * e;
        e = A[k] - C[k]; e *= e; s += e * e;
        e = B[k] - C[k]; e *= e; s += e * e;
    }
    return s;
}
__device__ __forceinline__ void edge_grad(const Tet12& t, float w, float* ga, float* gb, float* gc, float* gd) {
    float w80 = w * 80.f;   // d/dx (20 dx)^4 = 4 * 20 * (20 dx)^3
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float A = t.a[k] * 20.f, B = t.b[k] * 20.f, C = t.c[k] * 20.f, D = t.d[k] * 20.f;
        float ad = A - D, bd = B - D, cd = C - D, ab = A - B, ac = A - C, bc = B - C;
        ad = ad * ad * ad; bd = bd * bd * bd; cd = cd * cd * cd; ab = ab * ab * ab; ac = ac * ac * ac; bc = bc * bc * bc;
        ga[k] += w80 * (ad + ab + ac);
        gb[k] += w80 * (bd - ab + bc);
        gc[k] += w80 * (cd - ac - bc);
        gd[k] -= w80 * (ad + bd + cd);
    }
}

// Signed volume V = -det[A-D;B-D;C-D]/6                                            (deftet.py:239-263)
__device__ __forceinline__ float tet_volume(const Tet12& t, float* a, float* b, float* c) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { a[k] = t.a[k] - t.d[k]; b[k] = t.b[k] - t.d[k]; c[k] = t.c[k] - t.d[k]; }
    float bc[3];
    cross3(b, c, bc);
    return -dot3(a, bc) / 6.0f;
}
__device__ __forceinline__ void volume_grad(const float* a, const float* b, const float* c, float w,
                                            float* ga, float* gb, float* gc, float* gd) {
    float bc[3], ca[3], ab[3];
    cross3(b, c, bc); cross3(c, a, ca); cross3(a, b, ab);
    float w6 = -w / 6.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float x = w6 * bc[k], y = w6 * ca[k], z = w6 * ab[k];
        ga[k] += x; gb[k] += y; gc[k] += z; gd[k] -= x + y + z;
    }
}

}  // namespace dtb
