// Micro-benchmark: FP32 CUDA-core pipe throughput on sm_100a for scalar FFMA, packed FFMA2 and mixes.
// Calibrates the "FP32 pipe roof" the fused loss kernel is measured against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipe_probe tools/fp32_pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: FFMA x8 chains, 1: FFMA2 x8 chains (16 lanes-ops), 2: 4 FFMA2 + 4 FFMA... see below
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b) {
    float s[16];
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x * 0.002f + i, threadIdx.x * 0.003f - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {          // 16 scalar FFMA = 16 lane-ops per thread
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = fmaf(s[i], a, b);
        } else if (MODE == 1) {   // 8 packed FFMA2 = 16 lane-ops
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __ffma2_rn(v[i], a2, b2);
        } else if (MODE == 2) {   // 4 FFMA2 + 8 FFMA = 16 lane-ops (half the lanes packed)
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = __ffma2_rn(v[i], a2, b2);
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = fmaf(s[i], a, b);
        } else if (MODE == 3) {   // 6 FFMA2 + 4 FFMA = 16 lane-ops
#pragma unroll
            for (int i = 0; i < 6; ++i) v[i] = __ffma2_rn(v[i], a2, b2);
#pragma unroll
            for (int i = 0; i < 4; ++i) s[i] = fmaf(s[i], a, b);
        } else if (MODE == 4) {   // 8 FADD2 = 16 lane-ops
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __fadd2_rn(v[i], a2);
        } else if (MODE == 5) {   // 16 FADD
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = s[i] + a;
        } else if (MODE == 6) {   // 16 FFMA + 4 MUFU.EX2 (co-issue check)
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = fmaf(s[i], a, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(s[i])); s[i] = y; }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += s[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int sms, float* out) {
    const int iters = 20000, blocks = sms * 8;
    probe<MODE><<<blocks, 256>>>(out, 100, 0.999f, 0.001f);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double laneops = (double)blocks * 256 * iters * 16;
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %7.2f T lane-op/s  = %6.1f lane-ops/clk/SM at %.0f MHz (nominal max clock)\n", name, ms,
           laneops / (ms * 1e-3) / 1e12, laneops / (ms * 1e-3) / ((double)clk_khz * 1e3) / sms, clk_khz / 1e3);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    run<0>("16 FFMA", sms, out);
    run<1>("8 FFMA2", sms, out);
    run<2>("4 FFMA2 + 8 FFMA", sms, out);
    run<3>("6 FFMA2 + 4 FFMA", sms, out);
    run<4>("8 FADD2", sms, out);
    run<5>("16 FADD", sms, out);
    run<6>("16 FFMA + 4 EX2", sms, out);
    return 0;
}
