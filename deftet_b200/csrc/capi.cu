// Library-level C ABI: error channel, version, device probe.
#include "common.cuh"
#include "deftet_b200.h"
#include <stdarg.h>
#include <string.h>

namespace dtb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

struct ProfSlot { cudaEvent_t a = nullptr, b = nullptr; bool recorded = false; };
// process-wide on purpose: autograd runs backward kernels from its own thread
static ProfSlot g_prof[PROF_NTAGS];
static volatile int g_prof_on = 0;
void prof_begin(int tag, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfSlot& p = g_prof[tag];
    if (!p.a) { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
    cudaEventRecord(p.a, st);
}
void prof_end(int tag, cudaStream_t st) {
    if (!g_prof_on) return;
    ProfSlot& p = g_prof[tag];
    if (!p.a) return;
    cudaEventRecord(p.b, st);
    p.recorded = true;
}
}  // namespace dtb

// Per-kernel timing hooks: dtb_profile_enable(1) makes the library bracket its dominant kernels with CUDA events on the
// launching stream; dtb_profile_elapsed(tag, &ms) synchronises on the stop event and returns the last duration.
// Tags: 0 energies_fwd, 1 energies_bwd, 2 pit_tet, 3 nn_query, 4 pfd_forward, 5 bary_backward.
extern "C" int dtb_profile_enable(int on) { dtb::g_prof_on = on; return 0; }
extern "C" int dtb_profile_elapsed(int tag, float* ms) {
    if (tag < 0 || tag >= dtb::PROF_NTAGS || !ms) { dtb::set_error("profile_elapsed: bad tag"); return dtb::DTB_EINVAL; }
    dtb::ProfSlot& p = dtb::g_prof[tag];
    if (!p.recorded) { *ms = -1.f; return 0; }
    DTB_CUDA(cudaEventSynchronize(p.b));
    DTB_CUDA(cudaEventElapsedTime(ms, p.a, p.b));
    return 0;
}

extern "C" const char* dtb_last_error(void) { return dtb::g_err; }
extern "C" int dtb_version(void) { return 100; }
extern "C" int dtb_device_is_sm100(int device) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) { dtb::check_cuda(e, "cudaGetDeviceProperties"); return -(int)e; }
    return p.major == 10 ? 1 : 0;
}

// ---- primitive self-test hooks (exercised by tests/, not used by the product path) ----------------
#include "prims.cuh"
extern "C" size_t dtb_prim_scan_workspace(size_t n) { return dtb::scan_workspace_bytes(n); }
extern "C" int dtb_prim_exclusive_scan_u32(const unsigned* in, unsigned* out, size_t n, unsigned* total, void* ws, size_t ws_bytes,
                                           void* stream) {
    return dtb::exclusive_scan_u32(in, out, n, total, ws, ws_bytes, (cudaStream_t)stream);
}
extern "C" size_t dtb_prim_sort_workspace(size_t n) { return dtb::sort_workspace_bytes(n); }
extern "C" int dtb_prim_radix_sort_pairs_u64(unsigned long long* keys_in, unsigned* vals_in, unsigned long long* keys_out,
                                             unsigned* vals_out, size_t n, int key_bits, void* ws, size_t ws_bytes, void* stream) {
    return dtb::radix_sort_pairs_u64(keys_in, vals_in, keys_out, vals_out, n, key_bits, ws, ws_bytes, (cudaStream_t)stream);
}
