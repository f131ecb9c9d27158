// deftet_b200 -- shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define DTB_SM_COUNT 148

namespace dtb {

// ---- error channel -------------------------------------------------------------------------------
// Every C-ABI entry point returns 0 on success or a negative DTB_E* / positive cudaError_t code and
// records a message retrievable with dtb_last_error() (thread-local, so DataParallel's per-GPU
// threads do not race: SURVEY.md section 8b threading row).
enum { DTB_OK = 0, DTB_EINVAL = -1, DTB_EWORKSPACE = -2, DTB_EOVERFLOW = -3 };
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define DTB_CUDA(call)                                                  \
    do {                                                                \
        int _e = dtb::check_cuda((call), #call);                        \
        if (_e) return _e;                                              \
    } while (0)
#define DTB_LAUNCH_CHECK(name) DTB_CUDA((cudaError_t)cudaGetLastError())
#define DTB_REQUIRE(cond, ...)                                          \
    do {                                                                \
        if (!(cond)) { dtb::set_error(__VA_ARGS__); return dtb::DTB_EINVAL; } \
    } while (0)

// arrays read with 16-byte vector loads / TMA bulk copies: a misaligned pointer must come back as an error, not as a device fault
#define DTB_ALIGNED16(p) ((((size_t)(p)) & 15) == 0)
#define DTB_REQUIRE_ALIGNED16(p, what) DTB_REQUIRE(DTB_ALIGNED16(p), "%s must be 16-byte aligned (got %p)", what, (const void*)(p))

// ---- optional per-kernel CUDA-event timing (bench.py roofline leg; off by default, never on under graph capture) ----
enum ProfTag { PROF_ENERGIES_FWD = 0, PROF_ENERGIES_BWD, PROF_PIT_TET, PROF_NN_QUERY, PROF_PFD_FORWARD, PROF_BARY_BWD, PROF_NTAGS };
void prof_begin(int tag, cudaStream_t st);
void prof_end(int tag, cudaStream_t st);

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Bump allocator over a caller-provided workspace (no cudaMalloc on the hot path).
struct Workspace {
    char* base; size_t size; size_t off; bool ok;
    Workspace(void* p, size_t n) : base((char*)p), size(n), off(0), ok(true) {}
    template <typename T> T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        if (base == nullptr) { off += bytes; return nullptr; }      // sizing pass
        if (off + bytes > size) { ok = false; off += bytes; return nullptr; }
        T* r = (T*)(base + off); off += bytes; return r;
    }
};

// ---- exact (non-contracted) fp32 arithmetic ------------------------------------------------------
// Predicates and distances that decide an *index* are evaluated with the rounding sequence of the
// reference source as written (one IEEE rounding per operator, no FMA contraction) so that the result
// is bit-identical to the CPU oracle (oracle/ is built with -ffp-contract=off) on every input.
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }

// ---- vertex-gradient scatter ---------------------------------------------------------------------------------
// grad is (.., V, stride) with stride 3 (dense xyz) or 4 (padded, 16-byte aligned): with stride 4 the three components go
// out as ONE vector reduction (REDG.E.ADD.F32x4, sm_90+), which matters because the scatter kernels are bound by the
// reduction issue rate, not by bandwidth.
__device__ __forceinline__ void grad_add3(float* grad, size_t vid, int stride, float x, float y, float z) {
    if (stride == 4) {
        atomicAdd(reinterpret_cast<float4*>(grad) + vid, make_float4(x, y, z, 0.f));
    } else {
        atomicAdd(grad + vid * 3, x); atomicAdd(grad + vid * 3 + 1, y); atomicAdd(grad + vid * 3 + 2, z);
    }
}

// ---- warp / block reductions ---------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- TMA (1-D bulk async copy) + mbarrier --------------------------------------------------------
// cp.async.bulk global->shared with mbarrier completion (SASS: UBLKCP).  Sizes and both addresses must be
// multiples of 16 bytes.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

}  // namespace dtb
