// common.cuh -- shared host/device helpers of libclica_sm100.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/clica.h"

namespace clica {

// ---- error reporting (thread-local; never throws across the C ABI) ------------------------------
char* last_error_buffer();   // defined in abi.cu
inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CLICA_CUDA_OK(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return clica::fail((int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                               __FILE__, __LINE__);                                               \
    } while (0)

#define CLICA_REQUIRE(cond, code, ...)                                                            \
    do {                                                                                          \
        if (!(cond)) return clica::fail((code), __VA_ARGS__);                                     \
    } while (0)

struct DeviceInfo {
    int sm_count = 0;
    int cc_major = 0;
    int cc_minor = 0;
    int ok = 0;
};
int get_device_info(DeviceInfo* out);   // cached per device; defined in abi.cu
int tc_sm_reserve();                    // SMs the persistent tcgen05 GEMMs leave free (clica_tc_set_sm_reserve); abi.cu

// ---- launch accounting / optional per-family CUDA-event timing (bench.py's roofline numbers) ------
enum KernelFamily : int {
    kFamLossFwd = 0,    // lpnce_fwd_kernel
    kFamLossBwd = 1,    // lpnce_bwd_kernel
    kFamLossAux = 2,    // finalize / prep / reduce
    kFamGemmTc = 3,     // tcgen05 GEMM
    kFamGemmSimt = 4,   // CUDA-core GEMM
    kFamAdam = 5,
    kFamMisc = 6,       // column sums, operand packing
    kNumFamilies = 7
};
// RAII: counts `launches` kernels for `family`; when profiling is enabled also brackets them with a pair of
// CUDA events on `st` (read back by clica_prof_collect).  Defined in abi.cu.
struct LaunchScope {
    LaunchScope(cudaStream_t st, int family, int launches = 1);
    ~LaunchScope();
    cudaStream_t st_;
    int slot_;
};

// true the first time it is called for (`flags`, current device): per-function attributes such as
// cudaFuncAttributeMaxDynamicSharedMemorySize are per device, so "set once per process" is not enough.
// Callers set the attributes when it returns true (idempotent, so a rare double set under a race is harmless).
struct PerDeviceOnce { bool done[64] = {}; };
inline bool first_on_this_device(PerDeviceOnce& o) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (o.done[dev]) return false;
    o.done[dev] = true;
    return true;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
// integer environment switch, read at every call (cheap; lets tests toggle experimental paths in-process)
int env_flag(const char* name, int dflt);   // defined in abi.cu
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- kernel launch with programmatic dependent launch (PDL) ------------------------------------------
// Every kernel of this library starts with pdl_wait() (griddepcontrol.wait: all memory operations of the preceding
// kernel in the stream are complete and visible) followed by pdl_trigger(), and is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may be scheduled -- and run their prologue (barrier
// init, TMEM allocation, tensor-map fetch) -- while the tail of the preceding kernel is still draining, instead of
// after a full kernel boundary.  Stream capture records these as programmatic edges, so the CUDA-graph step keeps
// them.  CLICA_PDL=0 launches with plain stream order.
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = env_flag("CLICA_PDL", 1) != 0 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ---- device helpers ------------------------------------------------------------------------------
#ifdef __CUDACC__

// programmatic dependent launch (see launch_k): wait for the preceding kernel's memory, then let the next one be scheduled
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 4-byte and 16-byte cp.async with zero-fill of the bytes beyond src_bytes (LDGSTS).
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// nearest-even rounding of an fp32 value to tf32 precision (10 explicit mantissa bits); the result is exactly
// representable in tf32, so the tensor core's own fp32 -> tf32 conversion cannot change it.
__device__ __forceinline__ float round_to_tf32(float v) {
    uint32_t u = __float_as_uint(v);
    u += 0xFFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    return __uint_as_float(u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__

}  // namespace clica
