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
}  // namespace dtb

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
