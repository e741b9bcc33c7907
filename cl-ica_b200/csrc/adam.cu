// adam.cu -- fused multi-tensor Adam step (one launch per <= 32 tensors instead of torch's foreach chain).
// Replaces torch.optim.Adam.step as used by /root/reference/main_mlp.py:283,312 (defaults: betas
// (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).  Update rule of torch's default path:
//   m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2
//   p -= (lr / (1 - b1^t)) * m / ( sqrt(v) / sqrt(1 - b2^t) + eps )
#include "common.cuh"

#include <math.h>

namespace clica {
namespace {

constexpr int kMaxTensors = 24;
constexpr int kChunk = 2048;   // elements per block (two float4 per thread)

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
    // optional per tensor: the updated parameter is ALSO written in the tensor-core operand format of the encoder GEMMs
    // ((hi, lo) planes with row pitch pack_ld, see gemm_tc.cuh) -- the weight re-pack that otherwise follows every
    // optimizer step as its own pass over the parameters (read 4 B + write 8 B per weight)
    float* pack_hi[kMaxTensors];
    float* pack_lo[kMaxTensors];      // null with pack_hi set: one plane holding the fp32 value (TF32 mode)
    int pack_cols[kMaxTensors];
    int pack_ld[kMaxTensors];
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

__device__ __forceinline__ float adam_update(float p, float gr, float& m, float& v, float b1, float b2, float eps,
                                             float step_size, float inv_sqrt_bc2) {
    m = b1 * m + (1.f - b1) * gr;
    v = b2 * v + (1.f - b2) * gr * gr;
    const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
    return p - step_size * (m / denom);
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
    float* __restrict__ hi = a.pack_hi[t];
    float* __restrict__ lo = a.pack_lo[t];
    const int cols = a.pack_cols[t], ld = a.pack_ld[t];
    const long long numel = a.numel[t];
    const float step_size = a.dev_scalars ? __ldg(a.dev_scalars) : a.step_size;
    const float inv_sqrt_bc2 = a.dev_scalars ? __ldg(a.dev_scalars + 1) : a.inv_sqrt_bc2;
    // 128-bit path: all four arrays 16-byte aligned (and whole rows of 4 for the packed copy)
    const bool vec = (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                        reinterpret_cast<uintptr_t>(v)) & 15u) == 0) && (hi == nullptr || (cols & 3) == 0);
#pragma unroll
    for (int i = 0; i < kChunk / 1024; ++i) {
        const long long e = base + 4 * (threadIdx.x + 256 * i);
        if (e >= numel) break;
        if (vec && e + 3 < numel) {
            const float4 gv = *reinterpret_cast<const float4*>(g + e);
            float4 mv = *reinterpret_cast<const float4*>(m + e), vv = *reinterpret_cast<const float4*>(v + e);
            float4 pv = *reinterpret_cast<const float4*>(p + e);
            pv.x = adam_update(pv.x, gv.x * a.grad_scale, mv.x, vv.x, a.beta1, a.beta2, a.eps, step_size, inv_sqrt_bc2);
            pv.y = adam_update(pv.y, gv.y * a.grad_scale, mv.y, vv.y, a.beta1, a.beta2, a.eps, step_size, inv_sqrt_bc2);
            pv.z = adam_update(pv.z, gv.z * a.grad_scale, mv.z, vv.z, a.beta1, a.beta2, a.eps, step_size, inv_sqrt_bc2);
            pv.w = adam_update(pv.w, gv.w * a.grad_scale, mv.w, vv.w, a.beta1, a.beta2, a.eps, step_size, inv_sqrt_bc2);
            *reinterpret_cast<float4*>(m + e) = mv;
            *reinterpret_cast<float4*>(v + e) = vv;
            *reinterpret_cast<float4*>(p + e) = pv;
            if (hi != nullptr) {
                const long long r = e / cols;
                const long long idx = r * ld + (e - r * cols);
                if (lo != nullptr) {
                    float4 h, l;
                    h.x = round_to_tf32(pv.x); h.y = round_to_tf32(pv.y); h.z = round_to_tf32(pv.z); h.w = round_to_tf32(pv.w);
                    l.x = round_to_tf32(pv.x - h.x); l.y = round_to_tf32(pv.y - h.y);
                    l.z = round_to_tf32(pv.z - h.z); l.w = round_to_tf32(pv.w - h.w);
                    *reinterpret_cast<float4*>(hi + idx) = h;
                    *reinterpret_cast<float4*>(lo + idx) = l;
                } else {
                    *reinterpret_cast<float4*>(hi + idx) = pv;
                }
            }
        } else {
            for (int j = 0; j < 4 && e + j < numel; ++j) {
                float mn = m[e + j], vn = v[e + j];
                const float pn = adam_update(p[e + j], g[e + j] * a.grad_scale, mn, vn, a.beta1, a.beta2, a.eps, step_size, inv_sqrt_bc2);
                m[e + j] = mn; v[e + j] = vn; p[e + j] = pn;
                if (hi != nullptr) {
                    const long long r = (e + j) / cols;
                    const long long idx = r * ld + ((e + j) - r * cols);
                    if (lo != nullptr) { const float h = round_to_tf32(pn); hi[idx] = h; lo[idx] = round_to_tf32(pn - h); }
                    else hi[idx] = pn;
                }
            }
        }
    }
}

}  // namespace
}  // namespace clica

using namespace clica;

namespace {
int adam_launch(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                float* const* exp_avg_sq, const int64_t* numel, float step_size, float inv_sqrt_bc2,
                const float* dev_scalars, float beta1, float beta2, float eps, float grad_scale, cudaStream_t st,
                float* const* pack_hi = nullptr, float* const* pack_lo = nullptr, const int* pack_cols = nullptr,
                const int* pack_ld = nullptr) {
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
            a.pack_hi[t] = pack_hi ? pack_hi[k] : nullptr;
            a.pack_lo[t] = (pack_hi && pack_lo) ? pack_lo[k] : nullptr;
            a.pack_cols[t] = (pack_hi && pack_hi[k]) ? pack_cols[k] : 1;
            a.pack_ld[t] = (pack_hi && pack_hi[k]) ? pack_ld[k] : 1;
            if (a.pack_hi[t]) {
                CLICA_REQUIRE(a.pack_cols[t] >= 1 && a.pack_ld[t] >= a.pack_cols[t] && numel[k] % a.pack_cols[t] == 0, CLICA_E_BADARG,
                              "adam_step: tensor %d: packed copy with %d columns, pitch %d does not tile %lld elements", k,
                              a.pack_cols[t], a.pack_ld[t], (long long)numel[k]);
            }
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

// Same as clica_adam_step_capturable; additionally tensor k with pack_hi[k] != NULL (a [numel/pack_cols[k]] x pack_cols[k]
// weight matrix) is re-written in the tensor-core operand format at pack_hi[k] / pack_lo[k] (row pitch pack_ld[k] floats;
// the padding columns are not touched): the encoder's next forward needs no separate clica_mlp_pack_weights pass.
extern "C" int clica_adam_step_capturable_packed(int n, float* const* params, const float* const* grads,
                                                 float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel,
                                                 float lr, float beta1, float beta2, float eps, void* step_state,
                                                 float grad_scale, float* const* pack_hi, float* const* pack_lo,
                                                 const int* pack_cols, const int* pack_ld, void* stream) {
    CLICA_REQUIRE(n >= 0 && (n == 0 || (params && grads && exp_avg && exp_avg_sq && numel)), CLICA_E_BADARG,
                  "adam_step_capturable_packed: null pointer");
    CLICA_REQUIRE(pack_hi == nullptr || (pack_cols && pack_ld), CLICA_E_BADARG, "adam_step_capturable_packed: pack_cols / pack_ld missing");
    CLICA_REQUIRE(step_state && (((uintptr_t)step_state) & 7u) == 0, CLICA_E_ALIGN,
                  "adam_step_capturable_packed: step_state must be an 8-byte aligned 16-byte device buffer");
    cudaStream_t st = (cudaStream_t)stream;
    { LaunchScope ls(st, kFamAdam); launch_k(adam_tick_kernel, 1, 1, 0, st, (long long*)step_state, lr, beta1, beta2); }
    CLICA_CUDA_OK(cudaGetLastError());
    return adam_launch(n, params, grads, exp_avg, exp_avg_sq, numel, 0.f, 0.f,
                       reinterpret_cast<const float*>((long long*)step_state + 1), beta1, beta2, eps, grad_scale, st,
                       pack_hi, pack_lo, pack_cols, pack_ld);
}
