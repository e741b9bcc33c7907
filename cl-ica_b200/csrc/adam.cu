// adam.cu -- fused multi-tensor Adam step (one launch per <= 32 tensors instead of torch's foreach chain).
// Replaces torch.optim.Adam.step as used by /root/reference/main_mlp.py:283,312 (defaults: betas
// (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).  Update rule of torch's default path:
//   m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2
//   p -= (lr / (1 - b1^t)) * m / ( sqrt(v) / sqrt(1 - b2^t) + eps )
#include "common.cuh"

#include <math.h>

namespace clica {
namespace {

constexpr int kMaxTensors = 32;
constexpr int kChunk = 1024;   // elements per block

struct AdamArgs {
    float* p[kMaxTensors];
    const float* g[kMaxTensors];
    float* m[kMaxTensors];
    float* v[kMaxTensors];
    long long numel[kMaxTensors];
    int chunk_start[kMaxTensors + 1];   // prefix sum of ceil(numel / kChunk)
    int n;
    float step_size;       // lr / (1 - b1^t)
    float inv_sqrt_bc2;    // 1 / sqrt(1 - b2^t)
    float beta1, beta2, eps, grad_scale;
    const float* dev_scalars;   // non-null: (step_size, inv_sqrt_bc2) live on the device (CUDA-graph replays)
};

// Device-resident step state of the capturable variant: 16 bytes = { int64 step; float step_size; float inv_sqrt_bc2 }.
// One thread advances the step count and refreshes the two bias-correction scalars (same double-precision
// formulas as the host path), so that a captured graph can be replayed without any host-side state.
__global__ void adam_tick_kernel(long long* state, float lr, float beta1, float beta2) {
    pdl_enter();
    const long long t = state[0] + 1;
    state[0] = t;
    float* sc = reinterpret_cast<float*>(state + 1);
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    sc[0] = (float)((double)lr / bc1);
    sc[1] = (float)(1.0 / sqrt(bc2));
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
    pdl_enter();
    int t = 0;
    while (t + 1 < a.n && a.chunk_start[t + 1] <= (int)blockIdx.x) ++t;
    const long long base = (long long)((int)blockIdx.x - a.chunk_start[t]) * kChunk;
    float* __restrict__ p = a.p[t];
    const float* __restrict__ g = a.g[t];
    float* __restrict__ m = a.m[t];
    float* __restrict__ v = a.v[t];
    const long long numel = a.numel[t];
    const float step_size = a.dev_scalars ? __ldg(a.dev_scalars) : a.step_size;
    const float inv_sqrt_bc2 = a.dev_scalars ? __ldg(a.dev_scalars + 1) : a.inv_sqrt_bc2;
#pragma unroll
    for (int i = 0; i < kChunk / 256; ++i) {
        const long long e = base + threadIdx.x + 256 * i;
        if (e < numel) {
            const float gr = g[e] * a.grad_scale;
            const float mn = a.beta1 * m[e] + (1.f - a.beta1) * gr;
            const float vn = a.beta2 * v[e] + (1.f - a.beta2) * gr * gr;
            m[e] = mn;
            v[e] = vn;
            const float denom = sqrtf(vn) * inv_sqrt_bc2 + a.eps;
            p[e] -= step_size * (mn / denom);
        }
    }
}

}  // namespace
}  // namespace clica

using namespace clica;

namespace {
int adam_launch(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                float* const* exp_avg_sq, const int64_t* numel, float step_size, float inv_sqrt_bc2,
                const float* dev_scalars, float beta1, float beta2, float eps, float grad_scale, cudaStream_t st) {
    for (int first = 0; first < n; first += kMaxTensors) {
        AdamArgs a;
        a.n = (n - first < kMaxTensors) ? (n - first) : kMaxTensors;
        int chunks = 0;
        for (int t = 0; t < a.n; ++t) {
            const int k = first + t;
            CLICA_REQUIRE(params[k] && grads[k] && exp_avg[k] && exp_avg_sq[k] && numel[k] >= 0, CLICA_E_BADARG,
                          "adam_step: tensor %d has a null pointer or negative size", k);
            a.p[t] = params[k]; a.g[t] = grads[k]; a.m[t] = exp_avg[k]; a.v[t] = exp_avg_sq[k];
            a.numel[t] = numel[k];
            a.chunk_start[t] = chunks;
            chunks += (int)((numel[k] + kChunk - 1) / kChunk);
        }
        a.chunk_start[a.n] = chunks;
        a.step_size = step_size;
        a.inv_sqrt_bc2 = inv_sqrt_bc2;
        a.dev_scalars = dev_scalars;
        a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.grad_scale = grad_scale;
        if (chunks == 0) continue;
        { LaunchScope ls(st, kFamAdam); launch_k(adam_kernel, chunks, 256, 0, st, a); }
        CLICA_CUDA_OK(cudaGetLastError());
    }
    return 0;
}
}  // namespace

extern "C" int clica_adam_step(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                               float* const* exp_avg_sq, const int64_t* numel, float lr, float beta1,
                               float beta2, float eps, int64_t step, float grad_scale, void* stream) {
    CLICA_REQUIRE(n >= 0 && (n == 0 || (params && grads && exp_avg && exp_avg_sq && numel)), CLICA_E_BADARG,
                  "adam_step: null pointer");
    CLICA_REQUIRE(step >= 1, CLICA_E_BADARG, "adam_step: step must be >= 1 (got %lld)", (long long)step);
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    return adam_launch(n, params, grads, exp_avg, exp_avg_sq, numel, (float)((double)lr / bc1),
                       (float)(1.0 / sqrt(bc2)), nullptr, beta1, beta2, eps, grad_scale, (cudaStream_t)stream);
}

extern "C" int clica_adam_step_capturable(int n, float* const* params, const float* const* grads,
                                          float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel,
                                          float lr, float beta1, float beta2, float eps, void* step_state,
                                          float grad_scale, void* stream) {
    CLICA_REQUIRE(n >= 0 && (n == 0 || (params && grads && exp_avg && exp_avg_sq && numel)), CLICA_E_BADARG,
                  "adam_step_capturable: null pointer");
    CLICA_REQUIRE(step_state && (((uintptr_t)step_state) & 7u) == 0, CLICA_E_ALIGN,
                  "adam_step_capturable: step_state must be an 8-byte aligned 16-byte device buffer");
    cudaStream_t st = (cudaStream_t)stream;
    { LaunchScope ls(st, kFamAdam); launch_k(adam_tick_kernel, 1, 1, 0, st, (long long*)step_state, lr, beta1, beta2); }
    CLICA_CUDA_OK(cudaGetLastError());
    return adam_launch(n, params, grads, exp_avg, exp_avg_sq, numel, 0.f, 0.f,
                       reinterpret_cast<const float*>((long long*)step_state + 1), beta1, beta2, eps, grad_scale, st);
}
