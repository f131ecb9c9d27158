/* `run.so` shims: the reference's ctypes wrappers load os.getcwd()/utils/lib/<name>/run.so and call
 * `void run(...)` (utils/lib/<name>/interface.py:14-18).  Each shim exports that exact symbol and forwards to the
 * matching dtb_host_* entry point of libdeftet_b200.so.  Built once per builder with -DSHIM_<NAME>. */
#include <stdint.h>
#include <stdio.h>

#include "deftet_b200.h"

static void report(int rc, const char* what) {
    if (rc) fprintf(stderr, "deftet_b200 %s failed (%d): %s\n", what, rc, dtb_last_error());
}

#if defined(SHIM_TET_POINT_ADJ)
void run(int* tet_list, int* edge_p, int* n_edge, int n_point, int n_tet) {
    report(dtb_host_tet_point_adj(tet_list, edge_p, n_edge, n_point, n_tet), "tet_point_adj");
}
#elif defined(SHIM_TET_ADJ_SHARE)
void run(int* tet_list, int* face_edge_p, int* n_face_edge_p, int n_point, int n_tet) {
    report(dtb_host_tet_adj_share(tet_list, face_edge_p, n_face_edge_p, n_point, n_tet), "tet_adj_share");
}
#elif defined(SHIM_TET_FACE_ADJ)
void run(int* tet_list, int* face_edge_p, int* n_face_edge_p, int n_point, int n_tet) {
    report(dtb_host_tet_face_adj(tet_list, face_edge_p, n_face_edge_p, n_point, n_tet), "tet_face_adj");
}
#elif defined(SHIM_COLAPS_V)
void run(float* point_p, int* map_array_p, int* inverse_idx_p, int* n_colaps_v_p, int n_point) {
    report(dtb_host_colaps_v(point_p, map_array_p, inverse_idx_p, n_colaps_v_p, n_point), "colaps_v");
}
#else
#error "define one of SHIM_TET_POINT_ADJ / SHIM_TET_ADJ_SHARE / SHIM_TET_FACE_ADJ / SHIM_COLAPS_V"
#endif
