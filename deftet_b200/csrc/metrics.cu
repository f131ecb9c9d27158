// N4: evaluation metrics next to the hot path -- exact point -> triangle-mesh distance.
// Stand-in for kal.metrics.trianglemesh.point_to_mesh_distance as the reference calls it (utils/point_cloud_utils.py:48-56
// hausdorff_distance; eval.py).  Kaolin is un-vendored and un-pinned in the reference: parity unpinned, the contract is restated in
// oracle/metrics.py (true Euclidean closest point on each triangle, first strict minimum in face order).
// One thread per query point; the faces stream through shared memory in tiles that every thread of the CTA scans together
// (uniform loop, broadcast shared-memory reads): FP32-ALU bound by construction, O(P*F) like Kaolin's own kernel.
#include "common.cuh"
#include "deftet_b200.h"

namespace dtb {

#define PMD_TILE 128

// closest point on triangle (a,b,c) to p by Voronoi-region classification; returns the squared distance and the feature:
// 0 interior, 1..3 vertex a/b/c, 4..6 edge ab/bc/ca
__device__ __forceinline__ float tri_closest_sq(const float* __restrict__ t, float px, float py, float pz, int& type) {
    float abx = t[3] - t[0], aby = t[4] - t[1], abz = t[5] - t[2];
    float acx = t[6] - t[0], acy = t[7] - t[1], acz = t[8] - t[2];
    float apx = px - t[0], apy = py - t[1], apz = pz - t[2];
    float d1 = abx * apx + aby * apy + abz * apz, d2 = acx * apx + acy * apy + acz * apz;
    float qx, qy, qz;
    if (d1 <= 0.f && d2 <= 0.f) { type = 1; qx = t[0]; qy = t[1]; qz = t[2]; }
    else {
        float bpx = px - t[3], bpy = py - t[4], bpz = pz - t[5];
        float d3 = abx * bpx + aby * bpy + abz * bpz, d4 = acx * bpx + acy * bpy + acz * bpz;
        if (d3 >= 0.f && d4 <= d3) { type = 2; qx = t[3]; qy = t[4]; qz = t[5]; }
        else {
            float vc = xmul(d1, d4) - xmul(d3, d2);      // rounded products: an FMA here leaves a residual on degenerate (b == c) faces
            if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
                float v = d1 / (d1 - d3);
                type = 4; qx = t[0] + v * abx; qy = t[1] + v * aby; qz = t[2] + v * abz;
            } else {
                float cpx = px - t[6], cpy = py - t[7], cpz = pz - t[8];
                float d5 = abx * cpx + aby * cpy + abz * cpz, d6 = acx * cpx + acy * cpy + acz * cpz;
                if (d6 >= 0.f && d5 <= d6) { type = 3; qx = t[6]; qy = t[7]; qz = t[8]; }
                else {
                    float vb = xmul(d5, d2) - xmul(d1, d6);
                    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
                        float w = d2 / (d2 - d6);
                        type = 6; qx = t[0] + w * acx; qy = t[1] + w * acy; qz = t[2] + w * acz;
                    } else {
                        float va = xmul(d3, d6) - xmul(d5, d4);
                        if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
                            float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
                            type = 5; qx = t[3] + w * (t[6] - t[3]); qy = t[4] + w * (t[7] - t[4]); qz = t[5] + w * (t[8] - t[5]);
                        } else {
                            float denom = 1.f / (va + vb + vc);
                            float v = vb * denom, w = vc * denom;
                            type = 0; qx = t[0] + abx * v + acx * w; qy = t[1] + aby * v + acy * w; qz = t[2] + abz * v + acz * w;
                        }
                    }
                }
            }
        }
    }
    float dx = px - qx, dy = py - qy, dz = pz - qz;
    return dx * dx + dy * dy + dz * dz;
}

__global__ void __launch_bounds__(256) point_to_mesh_kernel(const float* __restrict__ points, const float* __restrict__ faces, int P, int F,
                                                            float* __restrict__ dist, long long* __restrict__ face_idx,
                                                            int32_t* __restrict__ dist_type) {
    __shared__ float tile[PMD_TILE * 9];
    int b = blockIdx.y;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float* fb = faces + (size_t)b * F * 9;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (i < P) { const float* p = points + ((size_t)b * P + i) * 3; px = p[0]; py = p[1]; pz = p[2]; }
    float best = 3.0e38f;
    int best_f = -1, best_t = 0;
    for (int f0 = 0; f0 < F; f0 += PMD_TILE) {
        int nf = min(PMD_TILE, F - f0);
        __syncthreads();
        for (int k = threadIdx.x; k < nf * 9; k += blockDim.x) tile[k] = fb[(size_t)f0 * 9 + k];
        __syncthreads();
        if (i < P) {
            for (int f = 0; f < nf; ++f) {
                int ty;
                float d = tri_closest_sq(tile + f * 9, px, py, pz, ty);
                if (d < best) { best = d; best_f = f0 + f; best_t = ty; }
            }
        }
    }
    if (i < P) {
        size_t o = (size_t)b * P + i;
        dist[o] = best_f >= 0 ? best : 0.f;
        if (face_idx) face_idx[o] = best_f;
        if (dist_type) dist_type[o] = best_t;
    }
}

}  // namespace dtb

using namespace dtb;

extern "C" int dtb_point_to_mesh_distance(const float* points, const float* face_vertices, int B, int P, int F, float* dist,
                                          long long* face_idx, int32_t* dist_type, void* stream) {
    DTB_REQUIRE(B >= 0 && P >= 0 && F >= 0, "point_to_mesh_distance: bad sizes");
    if ((long long)B * P == 0) return DTB_OK;
    DTB_REQUIRE(points && dist && (F == 0 || face_vertices), "point_to_mesh_distance: null argument");
    DTB_REQUIRE(B <= 65535, "point_to_mesh_distance: batch %d exceeds the grid limit", B);
    dim3 grid(cdiv(P, 256), B);
    point_to_mesh_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, face_vertices, P, F, dist, face_idx, dist_type);
    DTB_LAUNCH_CHECK("point_to_mesh");
    return DTB_OK;
}
