// Micro-probe: back-to-back tcgen05.mma cost (cycles per instruction) for kind::tf32 vs kind::f16 (bf16) at N = 128 / 256,
// operands in (uninitialised) shared memory, no TMA traffic.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 ...
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)lt << 61;
    return d;
}
template <int KIND>   // 0 tf32, 1 bf16
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND, int N, int MNMAJOR>
__global__ void __launch_bounds__(128, 1) probe(long long* cycles, int n_mma, int nslices) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    // zero the operand region so that the math is finite
    for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t fmt = (KIND == 0) ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)MNMAJOR << 15) | ((uint32_t)MNMAJOR << 16) |
                               ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a_base = base, b_base = base + 64 * 1024;
        long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const int ks = i & (nslices - 1);   // walk over different K slices / stages like a real main loop
            uint64_t da, db;
            if (MNMAJOR) { da = make_desc(a_base + ks * 1024, 4096, 512, 1); db = make_desc(b_base + ks * 1024, 4096, 512, 1); }
            else { da = make_desc(a_base + (ks / 4) * 16384 + (ks % 4) * 32, 16, 1024, 2); db = make_desc(b_base + (ks / 4) * 32768 + (ks % 4) * 32, 16, 1024, 2); }
            mma<KIND>(tmem, da, db, idesc, i > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

template <int KIND, int N, int MN>
void run(const char* name, int blocks) {
    long long* d; cudaMalloc(&d, sizeof(long long) * blocks);
    const int n_mma = 4096;
    cudaFuncSetAttribute(probe<KIND, N, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe<KIND, N, MN><<<blocks, 128, 180 * 1024>>>(d, 64, 8);
    probe<KIND, N, MN><<<blocks, 128, 180 * 1024>>>(d, n_mma, 8);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[1024]; cudaMemcpy(h, d, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += h[i]; mean /= blocks;
    const double kper = (KIND == 0) ? 8 : 16;
    printf("%-34s blocks=%3d  %s  %8.1f cyc/MMA  -> %7.1f flop/clk/SM\n", name, blocks, cudaGetErrorString(e), mean / n_mma,
           2.0 * 128 * N * kper * n_mma / mean);
    cudaFree(d);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int blocks : {1, sms}) {
        run<0, 128, 0>("tf32 K-major N=128", blocks);
        run<0, 256, 0>("tf32 K-major N=256", blocks);
        run<0, 128, 1>("tf32 MN-major N=128", blocks);
        run<1, 128, 0>("bf16 K-major N=128", blocks);
        run<1, 256, 0>("bf16 K-major N=256", blocks);
    }
    return 0;
}
