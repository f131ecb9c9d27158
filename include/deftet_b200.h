/* deftet_b200 -- C ABI of the Blackwell-native differentiable-tetrahedra engine.
 *
 * One shared library (deftet_b200/libdeftet_b200.so, sm_100a only, no CPU fallback) exports everything the
 * reference's per-tetrahedron hot path binds through pybind11 / ctypes.  Plain
 * pointers and sizes only; no torch types.  Two families of entry points:
 *
 *   dtb_*            device-pointer API: all array arguments are DEVICE pointers, work is enqueued on
 *                    `stream` (a cudaStream_t passed as void*) of the CURRENT device, nothing synchronises
 *                    unless stated.  Temporary memory comes from a caller-provided workspace whose size the
 *                    matching *_workspace() function reports (so the hot path never calls cudaMalloc).
 *   dtb_host_*       host-pointer API with exactly the argument lists of the reference's ctypes
 *                    `extern "C" void run(...)` builders (utils/lib/<name>/run.cpp) -- blocking, copies
 *                    in/out itself.
 *
 * Return value: 0 on success, <0 DTB_E* or >0 cudaError_t; dtb_last_error() gives the message (thread
 * local).  All functions are re-entrant and keep no global mutable state besides that message, which is
 * what nn.DataParallel's one-thread-per-GPU calling pattern needs (train_multigpu.py:136-140).
 *
 * Each declaration cites the reference interface it replaces (file:line relative to the reference root).
 */
#ifndef DEFTET_B200_H
#define DEFTET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTB_ENERGY_AMIPS 1
#define DTB_ENERGY_EDGE 2
#define DTB_ENERGY_VOLUME 4
#define DTB_ENERGY_ALL 7

/* ---- library ------------------------------------------------------------------------------------ */
/* Gradient buffers named grad_pos come with a grad_stride argument where noted: 3 = dense (..,V,3); 4 = padded (..,V,4), 16-byte
 * aligned, xyz in the first three floats -- the scatter then issues one vector reduction per vertex instead of three scalar ones. */
const char* dtb_last_error(void);
int dtb_version(void);                 /* 100 * major + minor */
int dtb_device_is_sm100(int device);   /* 1 if the device is compute capability 10.x, 0 if not, <0 on error */
/* Per-kernel CUDA-event timing for benchmarks (off by default; do not enable under CUDA-graph capture).  Tags:
 * 0 energies_fwd, 1 energies_bwd, 2 pit_tet, 3 nn_query, 4 pfd_forward, 5 bary_backward.  elapsed() synchronises. */
int dtb_profile_enable(int on);
int dtb_profile_elapsed(int tag, float* ms);

/* ---- A6/A7/A8: per-tet energies -------------------------------------------------------------------
 * Replaces DefTet.amips_energy / volume_variance / edge_length + autograd
 * (layers/DefTet/deftet.py:266-298, 239-263, 320-338) and tet_inverse_v/my_inverse (:205-233, 300-318).
 * pos (B,V,3) f32, tet (T,4) i32 (shared by the batch), inv_v (T,3,3) f32; outputs (B,) f32 each.
 * stats: (B,8) f64 scratch that forward fills and backward reads (sums, centred moments, mean volume).
 * flags: DTB_ENERGY_* mask; outputs of unselected energies are untouched and may be NULL.
 * backward ACCUMULATES into grad_pos (B,V,3) (caller zero-fills); g_* are the upstream (B,) gradients, a
 * NULL g_* means "no gradient requested for that energy".
 * The *_soup variants take the materialised tet_bxfx4x3 (B,T,4,3) tensor the reference methods receive
 * (deftet.py:66-68) and write a dense (B,T,4,3) gradient (overwritten, not accumulated). */
size_t dtb_tet_energies_workspace(int B, int V, int T);
int dtb_tet_energies_forward(const float* pos, const int32_t* tet, const float* inv_v, int B, int V, int T, int flags,
                             float* amips, float* edge, float* volvar, double* stats, void* workspace,
                             size_t workspace_bytes, void* stream);
int dtb_tet_energies_backward(const float* pos, const int32_t* tet, const float* inv_v, int B, int V, int T, int flags,
                              const double* stats, const float* g_amips, const float* g_edge, const float* g_volvar,
                              float* grad_pos, void* stream);
/* padded-gradient variant: grad_pos4 is a zero-filled (B,V,4) f32 buffer; one vector reduction per vertex update */
int dtb_tet_energies_backward_v4(const float* pos, const int32_t* tet, const float* inv_v, int B, int V, int T, int flags,
                                 const double* stats, const float* g_amips, const float* g_edge, const float* g_volvar,
                                 float* grad_pos4, void* stream);
int dtb_tet_energies_forward_soup(const float* tet_bxfx4x3, const float* inv_v, int B, int T, int flags, float* amips,
                                  float* edge, float* volvar, double* stats, void* workspace, size_t workspace_bytes,
                                  void* stream);
int dtb_tet_energies_backward_soup(const float* tet_bxfx4x3, const float* inv_v, int B, int T, int flags,
                                   const double* stats, const float* g_amips, const float* g_edge,
                                   const float* g_volvar, float* grad_soup, void* stream);
int dtb_tet_inverse_v(const float* pos0, const int32_t* tet, int V, int T, float* inv_v, void* stream);
/* Tile-local form of the same energies (csrc/energies_tiled.cu): the topology is re-encoded once per grid into tiles of 256
 * consecutive tets (distinct vertex ids, 2-byte local corner ids, per-vertex incidence lists), then forward / backward stage
 * each tile's vertices in shared memory, reduce per-sample sums without a second pass over the volumes, and scatter ONE vector
 * reduction per (vertex, tile, sample).  Same inputs, outputs, stats layout and tolerances as dtb_tet_energies_forward /
 * _backward_v4 (layers/DefTet/deftet.py:239-338 + autograd).
 *   tiles      256-byte aligned device buffer of dtb_tet_tiles_bytes(T) bytes, filled by dtb_tet_tiles_build
 *   nloc_max   device int written by the builder: the largest number of distinct vertices of any tile; the caller reads it
 *              back once (set-up time) and passes it to the forward / backward calls                                        */
size_t dtb_tet_tiles_bytes(int T);
int dtb_tet_tiles_build(const int32_t* tet, int T, int V, void* tiles, size_t tiles_bytes, int32_t* nloc_max, void* stream);
int dtb_tet_energies_forward_tiled(const float* pos, const int32_t* tet, const float* inv_v, const void* tiles, int nloc_max, int B,
                                   int V, int T, int flags, float* amips, float* edge, float* volvar, double* stats, void* stream);
int dtb_tet_energies_backward_tiled(const float* pos, const float* inv_v, const void* tiles, int nloc_max, int B, int V, int T,
                                    int flags, const double* stats, const float* g_amips, const float* g_edge,
                                    const float* g_volvar, float* grad_pos4, void* stream);

/* ---- A1: point-in-tet occupancy query + barycentric weights ------------------------------------------
 * Replaces check_condition_cuda_tet_base.forward(tet_bxfx4x3, point_pos_bxnx3, condition_bxnx1, bbox_filter_bxfx6)
 * (layers/DefTet/check_condition_tetrahedron_base/check_condition_tet.cpp:31-48, kernel
 * check_condition_tet_for.cu:124-189): cond (B,P) f32 = index of the FIRST tet (ascending id) whose four
 * same-side predicates agree, else -1.  bbox_filter_bxfx6 of the reference is dead (kernel lines :154-164
 * are commented out) and has no counterpart.  bary (B,P,4) f32, optional: weights of
 * utils/tet_utils.py:28-45 `bary_centric_tet` w.r.t. the found tet (zeros where cond == -1).
 * G: cells per axis of the uniform grid the points are binned into, <= 0 selects dtb_point_in_tet_grid_res.
 * _soup takes the materialised (B,T,4,3) tensor exactly as the reference does; the indexed form takes
 * pos (B,V,3) + tet (T,4) i32 and never materialises it.
 * dtb_tet_barycentric_backward: the reference defines no backward (utils.py:56-58 returns None, None); this
 * is autograd of bary_centric_tet: g_w (B,P,4) -> ACCUMULATES into grad_pos (B,V,3) (may be NULL) and
 * overwrites grad_points (B,P,3) (may be NULL). */
int dtb_point_in_tet_grid_res(int T, int P);
size_t dtb_point_in_tet_workspace(int B, int P, int T, int G);
int dtb_point_in_tet(const float* pos, const int32_t* tet, const float* points, int B, int V, int T, int P, int G,
                     float* cond, float* bary, void* workspace, size_t workspace_bytes, void* stream);
int dtb_point_in_tet_soup(const float* tet_bxfx4x3, const float* points, int B, int T, int P, int G, float* cond,
                          float* bary, void* workspace, size_t workspace_bytes, void* stream);
int dtb_tet_barycentric_backward(const float* pos, const int32_t* tet, const float* points, const float* cond,
                                 const float* g_w, int B, int V, int T, int P, float* grad_pos, int grad_stride,
                                 float* grad_points, void* stream);

/* Interpolation of a per-vertex field (B,V,C) at the query points through the weights of dtb_point_in_tet
 * (the differentiable form of DefTet.paste_occ, layers/DefTet/deftet.py:132-136): out (B,P,C), zeros where
 * cond == -1.  backward ACCUMULATES into g_field (may be NULL) and overwrites g_bary (B,P,4) (may be NULL). */
int dtb_tet_interpolate_forward(const float* field, const int32_t* tet, const float* cond, const float* bary, int B, int V, int C,
                                int P, float* out, void* stream);
int dtb_tet_interpolate_backward(const float* field, const int32_t* tet, const float* cond, const float* bary,
                                 const float* g_out, int B, int V, int C, int P, float* g_field, float* g_bary, void* stream);

/* Masked MSE over the located points: loss[b] = sum_i [cond_i >= 0] (x_i - t_i)^2 / max(#located, 1); x, t, cond (B,P);
 * acc (B,2) f64 scratch reused by backward, which overwrites g_x (B,P). */
int dtb_masked_mse_forward(const float* x, const float* t, const float* cond, int B, int P, double* acc, float* loss, void* stream);
int dtb_masked_mse_backward(const float* x, const float* t, const float* cond, const double* acc, const float* g_loss, int B, int P,
                            float* g_x, void* stream);

/* ---- A2: 1-nearest-neighbour index (one-sided chamfer) -----------------------------------------------
 * Replaces nearest_neighbor_cuda.forward(queries, points, result_int32, batch, nq, np, dim=3)
 * (layers/nearest_neighbor/nearest_neighbor.cpp:34-52, kernel nearest_neighbor_cuda.cu:17-55):
 * result (B,Q) i32 = argmin_j |points[b,j] - queries[b,i]|^2, strict <, lowest index wins ties. */
int dtb_nearest_neighbor_grid_res(int M);
size_t dtb_nearest_neighbor_workspace(int B, int Q, int M, int G);
int dtb_nearest_neighbor(const float* queries, const float* points, int32_t* result, int B, int Q, int M, int G,
                         void* workspace, size_t workspace_bytes, void* stream);

int dtb_nearest_neighbor_ragged(const float* queries, const int32_t* q_counts, int q_mult, const float* points,
                                int32_t* result, int B, int Qmax, int M, int G, void* workspace, size_t workspace_bytes,
                                void* stream);

/* ---- A9 / A3: predicted-surface stage in a padded-ragged batch layout -----------------------------------
 * The reference handles the ragged per-sample boundary sets with Python lists and a per-sample loop
 * (layers/DefTet/deftet.py:89-103).  Here: faces (B,Fmax,3) i32 + counts (B,) i32, no host sync.
 * dtb_boundary_faces = DefTet.get_boundary_index (deftet.py:186-195): interior faces face_fx3 (F,3) with their
 * two tets face_tet_fx2 (F,2); a face is on the predicted surface of sample b when occ[b,t0]+occ[b,t1] == 1; it
 * is emitted in table order, winding reversed when occ[b,t0] == 1.  *overflow is set to 1 if a sample has
 * more than Fmax such faces (extra faces are dropped).
 * dtb_surface_sample = mesh_utils.sample_surf_point_batch (utils/mesh_utils.py:290-299) with caller-provided
 * u = sqrt(rand), v = rand of shape (B,Fmax,S): q (B,Fmax*S,3).
 * dtb_chamfer_forward/backward = mesh_utils.point_point_distance + mean (utils/mesh_utils.py:360-366,
 * deftet.py:177,180): loss[b] = mean_i sqrt(|q_i - gt[nn_i]|^2 + 1e-10) over the counts[b]*S samples (1 when
 * the surface is empty, deftet.py:162-166); backward scatters through the sampling weights into grad_pos. */
size_t dtb_boundary_faces_workspace(int B, int F);
int dtb_boundary_faces(const int32_t* face_fx3, const int32_t* face_tet_fx2, const float* occ, int B, int T, int F, int Fmax,
                       int32_t* out_faces, int32_t* out_counts, int32_t* overflow, void* workspace, size_t workspace_bytes,
                       void* stream);
int dtb_surface_sample(const float* pos, const int32_t* faces, const int32_t* counts, const float* u, const float* v, int B,
                       int V, int Fmax, int S, float* q, void* stream);
int dtb_chamfer_forward(const float* q, const int32_t* nn, const float* gt, const int32_t* counts, int B, int Fmax, int S,
                        int M, double* acc, float* loss, void* stream);
int dtb_chamfer_backward(const float* q, const int32_t* nn, const float* gt, const int32_t* faces, const int32_t* counts,
                         const float* u, const float* v, const float* g_loss, int B, int V, int Fmax, int S, int M,
                         float* grad_pos, int grad_stride, void* stream);
int dtb_face_soup(const float* pos, const int32_t* faces, const int32_t* counts, int B, int V, int Fmax, float* soup,
                  void* stream);

/* ---- A4: point -> triangle-set squared distance ------------------------------------------------------------
 * Replaces tet_analytic_distance_batch.forward(points, faces, closest_f, closest_d, n_face_b) and
 * .backward(points, faces, closest_f, dl_dclosest_d, dldtet) (layers/DefTet/tet_analytic_distance_batch/
 * tet_analytic_distance.cpp:29-78; kernels tet_analytic_distance_for.cu:257-307, _back.cu:592-686).
 * points (B,S,3), faces (B,Fmax,3,3) f32, counts (B,) i32 (the reference's float n_face_b), closest_d /
 * closest_f (B,S) f32 (face id as float, -1 when no face is visible).  backward ACCUMULATES into dldface
 * (B,Fmax,3,3) (caller zero-fills, like utils.py:65).  _backward_indexed is the engine form: chains
 * loss_b = mean_i sqrt(d_i + 1e-10) (utils/mesh_utils.py:368-374) and scatters to grad_pos through the face
 * vertex ids.  dtb_sqrt_mean: out[b] = mean_i sqrt(d[b,i] + eps) (1 when counts[b] == 0). */
int dtb_point_face_distance_grid_res(int Fmax);
size_t dtb_point_face_distance_workspace(int B, int S, int Fmax, int G);
int dtb_point_face_distance_forward(const float* points, const float* faces, const int32_t* counts, int B, int S, int Fmax,
                                    int G, float* closest_d, float* closest_f, void* workspace, size_t workspace_bytes,
                                    void* stream);
int dtb_point_face_distance_backward(const float* points, const float* faces, const float* closest_f, const float* dl_dd,
                                     int B, int S, int Fmax, float* dldface, void* stream);
int dtb_point_face_distance_backward_indexed(const float* points, const float* soup, const int32_t* faces,
                                             const float* closest_f, const float* closest_d, const float* g_loss, int B,
                                             int S, int Fmax, int V, float* grad_pos, int grad_stride, void* stream);
int dtb_sqrt_mean(const float* d, const int32_t* counts, int B, int S, float eps, double* acc, float* out, void* stream);

/* ---- A5: boundary-face edge adjacency + normal-consistency loss -----------------------------------------------
 * Replaces tet_face_adj_m_idx.forward(face_fx3x3, adj_idx[F,30]) (layers/DefTet/tet_face_adj_m_idx/
 * tet_face_adj_m.cpp:26-34, kernel tet_face_adj_m_for.cu:72-108): per face the first 30 faces (ascending id)
 * that share an edge, -1 padded.  soup != NULL groups vertices by coordinate value (what the reference
 * does); soup == NULL groups by the vertex ids in faces (B,Fmax,3).  counts may be NULL.
 * dtb_normal_loss_* = get_surface_normal_loss (utils/mesh_utils.py:16-39): mean over directed pairs of
 * 1 - n_i.n_j, n = cross / sqrt(|cross|^2 + 1e-12) (:42-53); 0 when there is no pair, 1 when no face. */
size_t dtb_face_adjacency_workspace(int B, int Fmax, int V);
int dtb_face_adjacency(const float* soup, const int32_t* faces, const int32_t* counts, int B, int Fmax, int V, float* adj_f32,
                       int32_t* adj_i32, int32_t* deg, void* workspace, size_t workspace_bytes, void* stream);
int dtb_normal_loss_forward(const float* pos, const int32_t* faces, const int32_t* counts, const int32_t* adj, int B, int V,
                            int Fmax, float* normals_ws, double* acc, float* loss, void* stream);
int dtb_normal_loss_backward(const float* pos, const int32_t* faces, const int32_t* counts, const int32_t* adj,
                             const float* normals_ws, const double* acc, const float* g_loss, int B, int V, int Fmax,
                             float* gn_ws, float* grad_pos, int grad_stride, void* stream);

/* ---- A10-A14: topology builders ------------------------------------------------------------------------------
 * Device-pointer forms (tet (T,4) i32 on the device; outputs sized by the caller for the worst case; counts
 * are written to device int32 scalars, no host sync):
 *   dtb_tet_point_adj    unique directed vertex pairs of utils/lib/tet_point_adj/run.cpp:20-56, sorted by (a,b)
 *                        (the reference order is libstdc++ hash order); optional weight[e] = 1/deg(a), the
 *                        values of the row-normalised matrix built at tet_point_adj/interface.py:42-54.
 *   dtb_tet_to_face      utils/tet_utils.py:208-256 tet_to_face: interior faces in first-occurrence order with
 *                        the winding of the first tet (local faces (0,1,2),(1,0,3),(2,3,0),(3,2,1)), their two
 *                        tets and local face slots, and the faces seen once (cube boundary) in first-occurrence
 *                        order.  counts[0] = interior, counts[1] = boundary.  Output pointers may be NULL.
 *   dtb_tet_adj_share    utils/lib/tet_adj_share/run.cpp:40-97: rows (t0,t1,f0),(t1,t0,f1) in ascending
 *                        face-key order; *n_out = number of shared faces (= rows / 2).
 *   dtb_tet_face_adj     utils/lib/tet_face_adj/run.cpp:18-92: ordered pairs of tet-faces (4t+i) sharing an
 *                        edge, grouped by the WRAPPED int32 edge key a*n+b in signed order (the reference
 *                        overflows for n_point > 46340 and that is part of its output); *n_pairs may exceed
 *                        `capacity`, in which case pairs is truncated.
 *   dtb_collapse_vertices utils/lib/colaps_v/run.cpp:18-59: map_array[i] = id (first-occurrence order) of the
 *                        point's "%.5f-%.5f-%.5f" key (so -0.00000 != 0.00000), inverse_idx[id] = first index.
 * Host-pointer forms dtb_host_*: the exact argument lists of the reference's `extern "C" void run(...)`; they
 * allocate, copy, run and synchronise themselves, and return an error code instead of void. */
size_t dtb_tet_point_adj_workspace(int n_point, int T);
int dtb_tet_point_adj(const int32_t* tet, int n_point, int T, int32_t* edges, float* weight, int32_t* n_edge, void* workspace,
                      size_t workspace_bytes, void* stream);
size_t dtb_tet_to_face_workspace(int T);
int dtb_tet_to_face(const int32_t* tet, int n_point, int T, int32_t* face_fx3, int32_t* face_tet_fx2, int32_t* face_slot_fx2,
                    int32_t* boundary_fx3, int32_t* counts, void* workspace, size_t workspace_bytes, void* stream);
size_t dtb_tet_adj_share_workspace(int T);
int dtb_tet_adj_share(const int32_t* tet, int n_point, int T, int32_t* out, int32_t* n_out, void* workspace,
                      size_t workspace_bytes, void* stream);
size_t dtb_tet_face_adj_workspace(int T);
int dtb_tet_face_adj(const int32_t* tet, int n_point, int T, int32_t* pairs, long long capacity, int32_t* n_pairs,
                     void* workspace, size_t workspace_bytes, void* stream);
size_t dtb_collapse_vertices_workspace(int N);
int dtb_collapse_vertices(const float* points, int N, int32_t* map_array, int32_t* inverse_idx, int32_t* n_unique,
                          void* workspace, size_t workspace_bytes, void* stream);
int dtb_host_tet_point_adj(const int32_t* tet_list, int32_t* edge_p, int32_t* n_edge, int n_point, int n_tet);
int dtb_host_tet_adj_share(const int32_t* tet_list, int32_t* face_edge_p, int32_t* n_face_edge_p, int n_point, int n_tet);
int dtb_host_tet_face_adj(const int32_t* tet_list, int32_t* face_edge_p, int32_t* n_face_edge_p, int n_point, int n_tet);
int dtb_host_colaps_v(const float* point_p, int32_t* map_array_p, int32_t* inverse_idx_p, int32_t* n_colaps_v_p, int n_point);

/* ---- A15: tet-face volume rasterizer (stand-in for kal.render.mesh.deftet_sparse_render) ---------------------
 * Call site: diff_render/diftet_6_subdiv/5_rendereq/deftetrneder.py:97-100 (Kaolin itself is third-party and
 * un-pinned: parity is defined by oracle/render_oracle.c, see DESIGN.md "parity unpinned").
 * pixel_coords (B,P,2), render_ranges (B,P,2) [zmin,zmax], face_z (B,F,3), face_xy (B,F,3,2), face_feat (B,F,3,D)
 * -> out_feat (B,P,K,D) f32 and out_idx (B,P,K) i64: per pixel the first K faces (ascending id) containing it with
 * interpolated z in range, ordered by z descending; void slots are 0 / -1.  eps: kaolin's 1e-8.
 * R: cells per axis of the face-binning grid (<=0: 64); pair_capacity: capacity for (cell, face) pairs (<=0: 8 per
 * face); *overflow (device int32) is set to 1 when it was too small (results are then invalid: retry larger).
 * backward ACCUMULATES into g_xy (B,F,3,2) and g_feat (B,F,3,D); no gradient to z or the pixel. */
size_t dtb_sparse_render_workspace(int B, int P, int F, int R, long long pair_capacity);
/* exact number of (cell, face) pairs for these inputs -> *n_pairs (device u32); workspace >= (4*B + B*F)*4 + 512 bytes */
int dtb_sparse_render_pair_count(const float* pixel_coords, const float* face_xy, int B, int P, int F, int R, unsigned* n_pairs,
                                 void* workspace, size_t workspace_bytes, void* stream);
int dtb_sparse_render_forward(const float* pixel_coords, const float* render_ranges, const float* face_z, const float* face_xy,
                              const float* face_feat, int B, int P, int F, int D, int K, float eps, int R, long long pair_capacity,
                              float* out_feat, long long* out_idx, int32_t* overflow, void* workspace, size_t workspace_bytes,
                              void* stream);
int dtb_sparse_render_backward(const float* pixel_coords, const float* face_xy, const float* face_feat, const long long* idx,
                               const float* g_out, int B, int P, int F, int D, int K, float eps, float* g_xy, float* g_feat,
                               void* stream);

/* Fused render + composite fast path: rendermeshcolor's deftet_sparse_render -> peel2mask chain
 * (diff_render/diftet_6_subdiv/5_rendereq/deftetrneder.py:31-64,97-113) without the (B,P,K,D) intermediate.
 * Feature channel 0 is the opacity: alpha_k = clamp(f_k[0], 1e-10, 1-1e-10), vis_k = alpha_k prod_{i<k}(1-alpha_i) over the K
 * depth-sorted slots (void slots: f = 0); out_color (B,P,D-1) = sum_k vis_k c_k + (1 - sum_k vis_k), out_mask (B,P,1) = sum_k vis_k.
 * 2 <= D <= 8.  backward recollects the hits with the face binning the forward left in `workspace` (pass the same workspace,
 * sizes, R and pair_capacity) and ACCUMULATES into g_xy (B,F,3,2) / g_feat (B,F,3,D). */
int dtb_render_composite_forward(const float* pixel_coords, const float* render_ranges, const float* face_z, const float* face_xy,
                                 const float* face_feat, int B, int P, int F, int D, int K, float eps, int R, long long pair_capacity,
                                 float* out_color, float* out_mask, int32_t* overflow, void* workspace, size_t workspace_bytes,
                                 void* stream);
int dtb_render_composite_backward(const float* pixel_coords, const float* render_ranges, const float* face_z, const float* face_xy,
                                  const float* face_feat, const float* g_color, const float* g_mask, int B, int P, int F, int D, int K,
                                  float eps, int R, long long pair_capacity, float* g_xy, float* g_feat, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* ---- A16: inside/outside labels (stand-in for kal.ops.mesh.check_sign, layers/DefTet/deftet.py:46) ------------
 * verts (B,n,3), faces (m,3) i32 shared by the batch, points (B,p,3) -> out (B,p) u8, 1 = inside (+z ray parity,
 * half-open edge rule; oracle/render_oracle.c).  R = kaolin's hash_resolution (xy grid; <=0: 256, max 1024).
 * Synchronises the stream once (capacity check of the triangle-cell list). */
size_t dtb_check_sign_workspace(int B, int m, int R);
int dtb_check_sign(const float* verts, const int32_t* faces, const float* points, int B, int n, int m, int p, int R,
                   unsigned char* out, void* workspace, size_t workspace_bytes, void* stream);
/* dtb_check_sign + the grid resolution finally used (<= R) in *r_used: blocking like dtb_check_sign (set-up time). */
int dtb_check_sign_probe(const float* verts, const int32_t* faces, const float* points, int B, int n, int m, int p, int R,
                         unsigned char* out, int* r_used, void* workspace, size_t workspace_bytes, void* stream);
/* Non-blocking dtb_check_sign at a resolution found earlier (dtb_check_sign_probe) for a mesh of this size: no host
 * synchronisation, capturable in a CUDA graph; if the (cell, triangle) list overflows the query tests every face on the device. */
int dtb_check_sign_fixed(const float* verts, const int32_t* faces, const float* points, int B, int n, int m, int p, int R,
                         unsigned char* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- A17: Laplacian smoothness sum ||(D^-1 A) d - d||^2 (layers/DefTet/deftet.py:340-343) -------------------
 * d (B,V,3); edges (E,2) + weight (E,) from dtb_tet_point_adj (sorted by row).  resid_ws (B,V,3) and rows_ws (V+1)
 * are scratch that backward reuses; backward ACCUMULATES into grad (B,V,3). */
int dtb_laplacian_forward(const float* d, const int32_t* edges, const float* weight, int B, int V, int E, float* resid_ws,
                          unsigned* rows_ws, double* acc, float* loss, void* stream);
int dtb_laplacian_backward(const float* resid_ws, const int32_t* edges, const float* weight, const unsigned* rows_ws,
                           const float* g_loss, int B, int V, float* grad, void* stream);

/* ---- N3: topology editing + regularisers of the diff_render optimisation loop ---------
 * The reference runs these on the host in numpy / Python loops (diff_render/diftet_6_subdiv/3_model); every
 * index-valued output below is in the reference's order.
 *
 * dtb_tet_edges: unique undirected edges (lo,hi) sorted lexicographically -> edges (<=6T,2) i32, *n_edge; tet_edge (T,6)
 *   i32 = edge id of the vertex pairs (0,1),(0,2),(0,3),(1,2),(1,3),(2,3) of each tet (may be NULL).
 *   prepare_for_wz.py:186-205 generate_edge, :208-238 matchedgelist / generate_tet_edge_idx.
 * dtb_subdivide_tets: 1->8 split of the tets flagged in subdiv (u8 per tet; NULL = all): out_tet gets the unflagged tets in
 *   order, then 8 children per flagged tet; new vertex id of edge e is n_point + e.  Capacity 8T rows; *n_out = rows.
 *   prepare_for_wz.py:257-301 generate_subdivision.
 * dtb_edge_midpoints: out (E,K) = (values[e0] + values[e1]) / 2 (positions, features, pointmov alike); :241-254.
 * dtb_tet_to_face_idx: unique faces, interior and boundary, first-occurrence order, -1 partner on the boundary.
 *   prepare_for_wz.py:49-108 tet_to_face_idx(with_boundary=True) as 3_model/deftet.py:141-143 calls it.
 * dtb_tet_neighbours: (T,4) i32, column i = tet across local face i ((0,1,2),(1,0,3),(2,3,0),(3,2,1)), -1 = none.
 *   utils_tetsv.py:16-62 tet_adj_share -> tet_neighbour_idx (same neighbour sets; the reference's column order is Python
 *   dict order, its consumers only take row maxima).
 * dtb_point_adj_rows / dtb_point_adj_table: fixed-width vertex neighbour table from the sorted directed edge list of
 *   dtb_tet_point_adj: rows() gives row_start/row_end (P,) i32, degree (P,) f32 and *max_degree; table() fills (P,M) i32,
 *   ascending neighbour ids, -1 padding.  prepare_for_wz.py:112-137 generate_point_adj(_idx).
 * dtb_tet_delete: keep tet t iff max over the vertex weights of all tets reached by `levels`-step walks through the neighbour
 *   table exceeds thres (a walk that leaves the mesh contributes 0); compacts the kept tets, *n_out = count, keep (T,) u8
 *   optional.  3_model/deftet.py:290-329 pointweights2tetweights / tetweights2tetneighbourweights / deletetet and
 *   prepare_for_wz.py:171-181 delete_tet (the reference materialises a (T, 4^(levels+1)) table).
 * dtb_featlap_*: out (P,C) = (sum_{j in table[i]} x_j / weight_i - x_i)^2; backward ACCUMULATES into grad_x (P,C).
 *   3_model/deftet.py:227-250 get_featlap (+ autograd).
 * dtb_tet_volume_deviation_*: out (T,) = V_t - mean V, V = signed volume of scale*pos; acc = one f64 of scratch; backward
 *   ACCUMULATES into grad_pos (P,3).  3_model/deftet.py:252-309 get_volume_variance (+ autograd).
 * dtb_project_faces_*: camera transform, perspective divide, optional sigmoid and per-face gather in one pass:
 *   face_z (B,F,3) camera-space z, face_xy (B,F,3,2) image xy * multiplier, face_feat (B,F,3,D); pos (P,3) and feat (P,D) are
 *   shared by the B views, cam_rot (B,3,3), cam_pos (B,3), cam_proj (3,).  backward ACCUMULATES into grad_pos (P,3) and
 *   grad_feat (P,D); any of g_face_* / grad_* may be NULL.  3_model/cameraop.py:14-33 perspective, 4_render/vertex2face.py:14-28,
 *   5_rendereq/deftetrneder.py:84 (sigmoid), 3_model/deftet.py:425-470 (the repeat over views). */
size_t dtb_tet_edges_workspace(int T);
int dtb_tet_edges(const int32_t* tet, int n_point, int T, int32_t* edges, int32_t* tet_edge, int32_t* n_edge, void* workspace,
                  size_t workspace_bytes, void* stream);
size_t dtb_subdivide_tets_workspace(int T);
int dtb_subdivide_tets(const int32_t* tet, const int32_t* tet_edge, const unsigned char* subdiv, int n_point, int T,
                       int32_t* out_tet, int32_t* n_out, void* workspace, size_t workspace_bytes, void* stream);
int dtb_edge_midpoints(const float* values, int K, const int32_t* edges, int E, float* out, void* stream);
size_t dtb_tet_to_face_idx_workspace(int T);
int dtb_tet_to_face_idx(const int32_t* tet, int n_point, int T, int32_t* face_fx3, int32_t* face_tet_fx2, int32_t* face_slot_fx2,
                        int32_t* n_face, void* workspace, size_t workspace_bytes, void* stream);
size_t dtb_tet_neighbours_workspace(int T);
int dtb_tet_neighbours(const int32_t* tet, int n_point, int T, int32_t* neighbour_tx4, void* workspace, size_t workspace_bytes,
                       void* stream);
int dtb_point_adj_rows(const int32_t* edges, int E, int n_point, int32_t* row_start, int32_t* row_end, float* degree,
                       int32_t* max_degree, void* stream);
int dtb_point_adj_table(const int32_t* edges, int E, int n_point, const int32_t* row_start, int M, int32_t* table, void* stream);
size_t dtb_tet_delete_workspace(int T);
int dtb_tet_delete(const int32_t* tet, const float* point_weight, const int32_t* neighbour, int T, int levels, float thres,
                   int32_t* out_tet, int32_t* n_out, unsigned char* keep, void* workspace, size_t workspace_bytes, void* stream);
int dtb_featlap_forward(const float* x, const int32_t* table, const float* weight, int P, int M, int C, float* out, void* stream);
int dtb_featlap_backward(const float* x, const int32_t* table, const float* weight, const float* g_out, int P, int M, int C,
                         float* grad_x, void* stream);
int dtb_tet_volume_deviation_forward(const float* pos, const int32_t* tet, int T, float scale, float* out, double* acc, void* stream);
int dtb_tet_volume_deviation_backward(const float* pos, const int32_t* tet, int T, float scale, const float* g_out, double* acc,
                                      float* grad_pos, void* stream);
int dtb_project_faces_forward(const float* pos, const float* feat, const int32_t* faces, const float* cam_rot, const float* cam_pos,
                              const float* cam_proj, int B, int F, int D, float multiplier, int sigmoid, float* face_z,
                              float* face_xy, float* face_feat, void* stream);
int dtb_project_faces_backward(const float* pos, const float* feat, const int32_t* faces, const float* cam_rot, const float* cam_pos,
                               const float* cam_proj, int B, int F, int D, float multiplier, int sigmoid, const float* g_face_xy,
                               const float* g_face_feat, const float* g_face_z, float* grad_pos, float* grad_feat, void* stream);

/* ---- N4: evaluation metric next to the hot path ---------------------------------------
 * Exact squared distance from each point to the closest triangle of a mesh; stand-in for
 * kal.metrics.trianglemesh.point_to_mesh_distance at utils/point_cloud_utils.py:48-56 (hausdorff_distance) -- Kaolin is
 * un-vendored and un-pinned: parity unpinned, contract in oracle/metrics.py.  points (B,P,3), face_vertices (B,F,3,3);
 * dist (B,P) f32, face_idx (B,P) i64 (first strict minimum in face order; may be NULL), dist_type (B,P) i32 (0 interior,
 * 1-3 vertex, 4-6 edge ab/bc/ca; may be NULL).  Forward only (the reference uses it under no_grad in evaluation). */
int dtb_point_to_mesh_distance(const float* points, const float* face_vertices, int B, int P, int F, float* dist,
                               long long* face_idx, int32_t* dist_type, void* stream);

/* ---- N4 (second half): voxel-feature sampling at grid vertices / tet centroids ---------
 * Replaces layers/pv_module/functional/devoxelization.py:47-53 trilinear_devoxelize -- the live definition: normalise
 * (coords*2+1)/r-1, flip, torch F.grid_sample(bilinear, padding_mode='border', align_corners=False) -- as
 * layers/pc_model.py:182-194 sample_f calls it per encoder level, and its autograd.
 * feat (B,C,R,R,R) f32 contiguous.  coords: voxel coordinates, element (b,k,n) at coords[b*cs_b + k*cs_k + n*cs_n]
 * ((B,3,N) contiguous = strides 3N, N, 1; the reference passes a permuted view = 3N, 1, 3).  With
 * DTB_DEVOX_FROM_POSITIONS the same array holds positions p and the kernel applies sample_f's own prelude
 * c = clamp((p + 0.5) * R, 0, R - 1) (pc_model.py:186-191), including its gradient.
 * out (B,C,N): element (b,c,n) at out[b*out_batch_stride + c*N + n] (out_batch_stride > C*N writes a channel slice of the
 * concatenated sample_f tensor in place).
 * backward: grad_out laid out like out.  grad_feat (B,C,R^3) is OVERWRITTEN (NULL to skip); grad_coords has the strides of
 * coords and is ACCUMULATED into (the caller zeroes it once for all levels; NULL to skip; needs feat).
 * DTB_DEVOX_GLOBAL_GATHER forces the no-staging kernels that serve R > 36; DTB_DEVOX_SIMPLE selects the one-point-per-thread
 * kernels instead of the four-points-per-thread ones; DTB_DEVOX_NO_OWNER / NO_SORT / FORCE_SORT choose the volume-gradient kernel
 * (all for self-tests and A/B timing; same results up to summation order). */
#define DTB_DEVOX_FROM_POSITIONS 1
#define DTB_DEVOX_GLOBAL_GATHER 2
#define DTB_DEVOX_SIMPLE 4
#define DTB_DEVOX_NO_OWNER 8      /* volume gradient: not the channel-owner kernel (R^3 <= 512) but the shared-atomic one */
#define DTB_DEVOX_NO_SORT 16      /* volume gradient: ignore the workspace (no sorted reduction) */
#define DTB_DEVOX_FORCE_SORT 32    /* volume gradient: sorted reduction whenever a workspace is given, also below 8 points per voxel */
int dtb_trilinear_devoxelize_forward(const float* feat, const float* coords, long long cs_b, long long cs_k, long long cs_n, int B,
                                     int C, int N, int R, int flags, float* out, long long out_batch_stride, void* stream);
int dtb_trilinear_devoxelize_backward(const float* feat, const float* coords, long long cs_b, long long cs_k, long long cs_n,
                                      const float* grad_out, long long grad_out_batch_stride, int B, int C, int N, int R, int flags,
                                      float* grad_feat, float* grad_coords, void* stream);
/* The same with temporary memory: the volume gradient is then computed from the points radix-sorted by voxel (no shared-memory
 * atomics), when a voxel collects 8 points or more on average.  *_workspace returns the bytes to provide (0: not applicable). */
size_t dtb_trilinear_devoxelize_backward_workspace(int B, int C, int N, int R, int flags);
int dtb_trilinear_devoxelize_backward_ws(const float* feat, const float* coords, long long cs_b, long long cs_k, long long cs_n,
                                         const float* grad_out, long long grad_out_batch_stride, int B, int C, int N, int R, int flags,
                                         float* grad_feat, float* grad_coords, void* workspace, size_t workspace_bytes, void* stream);

/* ---- N2: graph-convolution neighbourhood product on the A10 adjacency ---------------------
 * Replaces utils/matrix_utils.py:22-33 sparse_batch_matmul (torch.sparse.mm on a transposed/reshaped copy of the dense operand)
 * as layers/gcn_decoder.py:44-56 GraphConv.forward calls it.
 * dtb_coo_to_csr: COO triplets (int64 row/col as torch sparse tensors hold them, f32 values, any order) -> CSR of the matrix
 *   (transpose = 0) or of its transpose (transpose = 1, used for the backward pass): row_ptr (n_major+1) i32, col (nnz) i32
 *   ascending within a row, val (nnz) f32.  Set-up time (once per adjacency).
 * dtb_spmm_csr: out (B,n_rows,p) = A @ x (B,n_cols,p) for every sample; HBM bound, 2*4*B*n*p algorithmic bytes. */
size_t dtb_coo_to_csr_workspace(long long nnz);
int dtb_coo_to_csr(const long long* rows, const long long* cols, const float* vals, long long nnz, int n_rows, int n_cols,
                   int transpose, int32_t* row_ptr, int32_t* col, float* val, void* workspace, size_t workspace_bytes, void* stream);
int dtb_spmm_csr(const int32_t* row_ptr, const int32_t* col, const float* val, const float* x, int B, int n_rows, int n_cols, int p,
                 float* out, void* stream);

/* ---- device-wide primitives (exported for the self-tests; also usable by integrators) ----------------- */
size_t dtb_prim_scan_workspace(size_t n);
int dtb_prim_exclusive_scan_u32(const unsigned* in, unsigned* out, size_t n, unsigned* total, void* ws, size_t ws_bytes,
                                void* stream);
size_t dtb_prim_sort_workspace(size_t n);
int dtb_prim_radix_sort_pairs_u64(unsigned long long* keys_in, unsigned* vals_in, unsigned long long* keys_out,
                                  unsigned* vals_out, size_t n, int key_bits, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEFTET_B200_H */
