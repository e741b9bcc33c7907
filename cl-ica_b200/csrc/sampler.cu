// sampler.cu -- device-side latent samplers: what the reference's spaces.py / spaces_utils.py draw on the
// host (CPU Gamma sampling, boolean-mask rejection loops with one .item() sync per round, allclose asserts) drawn by
// one kernel launch on the GPU, with no host synchronisation.
//
// Replaces, for CUDA devices:
//   /root/reference/spaces.py:47-119   NRealSpace.normal / laplace / generalized_normal
//   /root/reference/spaces.py:134-231  NSphereSpace.uniform / normal / laplace / generalized_normal (draw in R^n, project)
//   /root/reference/spaces.py:273-351  NBoxSpace.uniform / normal / laplace / generalized_normal (per-element rejection)
//   /root/reference/spaces_utils.py:82-103   sample_generalized_normal (sign * Gamma(1/p, 1)^(1/p))
//   /root/reference/spaces_utils.py:106-142  truncated_rejection_resampling
// Same distributions, different random stream: Philox4x32-10, counter = (element index, attempt, call offset), key =
// seed -- every element owns its counters, so the result is deterministic whatever the launch geometry.
// One thread per row (the sphere projection needs the row norm; n <= 1024).
#include "common.cuh"

#include <math.h>

namespace clica {
namespace {

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)

struct SampleParams {
    float* out; int ld; int rows; int n;
    int space;                 // 0 R^n, 1 sphere (project to unit norm), 2 box (per-element rejection into [lo, hi])
    int dist;                  // 0 uniform, 1 normal, 2 laplace, 3 generalized normal
    const float* mean; int ld_mean; int mean_rows;     // mean_rows: 0 none, 1 one row for all, else per row
    float scale; float p; float lo; float hi;
    uint32_t seed_lo, seed_hi, off_lo, off_hi;
};

// one draw of the unit-scale, zero-mean variate of `dist` for element e, attempt t
__device__ __forceinline__ float draw(const Philox& ph, const SampleParams& q, uint32_t e, uint32_t t) {
    const uint4 r = ph(e, t, q.off_lo, q.off_hi);
    if (q.dist == 1) {                                   // N(0, 1): Box-Muller
        return sqrtf(-2.f * __logf(u01(r.x))) * __cosf(6.283185307179586f * u01(r.y));
    }
    if (q.dist == 2) {                                   // Laplace(0, 1): inverse CDF
        const float u = u01(r.x) - 0.5f;
        return -copysignf(__logf(1.f - 2.f * fabsf(u)), u);
    }
    if (q.dist == 3) {                                   // sign * Gamma(1/p, 1)^(1/p)
        // Marsaglia-Tsang for shape a + 1 >= 1, then the boost Gamma(a) = Gamma(a + 1) * U^(1/a)
        const float a = 1.f / q.p;
        const float dd = a + 1.f - (1.f / 3.f), cc = rsqrtf(9.f * dd);
        float g = dd;
        for (uint32_t k = 0; k < 64u; ++k) {
            const uint4 s = ph(e, t, q.off_lo ^ 0x9E3779B9u, q.off_hi + 1u + k);
            const float x = sqrtf(-2.f * __logf(u01(s.x))) * __cosf(6.283185307179586f * u01(s.y));
            float v = 1.f + cc * x;
            if (v <= 0.f) continue;
            v = v * v * v;
            if (__logf(u01(s.z)) < 0.5f * x * x + dd - dd * v + dd * __logf(v)) { g = dd * v; break; }
        }
        g *= __powf(u01(r.x), q.p);                      // U^(1/a), 1/a = p
        const float mag = __powf(g, a);                  // Gamma^(1/p)
        return (r.y & 1u) ? mag : -mag;
    }
    return u01(r.x);                                     // uniform on (0, 1)
}

__global__ void __launch_bounds__(128) sampler_kernel(const SampleParams q) {
    pdl_enter();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= q.rows) return;
    const Philox ph{q.seed_lo, q.seed_hi};
    float* o = q.out + (size_t)row * q.ld;
    const float* m = q.mean_rows == 0 ? nullptr : q.mean + (size_t)(q.mean_rows == 1 ? 0 : row) * q.ld_mean;
    float nrm = 0.f;
    for (int c = 0; c < q.n; ++c) {
        const uint32_t e = (uint32_t)row * (uint32_t)q.n + (uint32_t)c;
        const float mu = m ? __ldg(m + c) : 0.f;
        float v;
        if (q.dist == 0) {
            if (q.space == 2) v = q.lo + (q.hi - q.lo) * draw(ph, q, e, 0u);           // U(lo, hi)
            else {                                                                      // sphere uniform: N(0, I), projected
                const uint4 r0 = ph(e, 0u, q.off_lo, q.off_hi);
                v = sqrtf(-2.f * __logf(u01(r0.x))) * __cosf(6.283185307179586f * u01(r0.y));
            }
        } else {
            v = mu + q.scale * draw(ph, q, e, 0u);
            if (q.space == 2) {                           // redraw THIS element until it lies in the box
                for (uint32_t t = 1; (v < q.lo || v > q.hi) && t < 4096u; ++t) v = mu + q.scale * draw(ph, q, e, t);
                v = fminf(fmaxf(v, q.lo), q.hi);          // (only after 4096 rejections: mean far outside the box)
            }
        }
        nrm = fmaf(v, v, nrm);
        o[c] = v;
    }
    if (q.space == 1) {
        const float inv = 1.f / sqrtf(nrm);               // IEEE sqrt + divide: the reference asserts | |z| - 1 | <= 1e-5 on these rows
        for (int c = 0; c < q.n; ++c) o[c] *= inv;
    }
}

}  // namespace
}  // namespace clica

using namespace clica;

extern "C" int clica_sample_latents(float* out, int ld, int rows, int n, int space, int dist, const float* mean,
                                    int ld_mean, int mean_rows, float scale, float p, float box_lo, float box_hi,
                                    uint64_t seed, uint64_t offset, void* stream) {
    CLICA_REQUIRE(out && rows >= 0 && n >= 1 && n <= 1024 && ld >= n, CLICA_E_BADARG, "sample_latents: bad output shape");
    CLICA_REQUIRE(space >= 0 && space <= 2 && dist >= 0 && dist <= 3, CLICA_E_BADARG, "sample_latents: unknown space / distribution");
    CLICA_REQUIRE(!(space == 0 && dist == 0), CLICA_E_UNSUPPORTED, "sample_latents: uniform is not defined on R^n (spaces.py:45-46)");
    CLICA_REQUIRE(mean_rows == 0 || (mean && ld_mean >= n && (mean_rows == 1 || mean_rows == rows)), CLICA_E_BADARG,
                  "sample_latents: mean must be one row or one row per sample");
    CLICA_REQUIRE(dist != 3 || p > 0.f, CLICA_E_BADARG, "sample_latents: generalized normal needs p > 0");
    CLICA_REQUIRE(space != 2 || box_hi > box_lo, CLICA_E_BADARG, "sample_latents: empty box");
    if (rows == 0) return 0;
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    SampleParams q;
    q.out = out; q.ld = ld; q.rows = rows; q.n = n; q.space = space; q.dist = dist;
    q.mean = mean; q.ld_mean = ld_mean; q.mean_rows = mean_rows; q.scale = scale; q.p = p; q.lo = box_lo; q.hi = box_hi;
    q.seed_lo = (uint32_t)seed; q.seed_hi = (uint32_t)(seed >> 32); q.off_lo = (uint32_t)offset; q.off_hi = (uint32_t)(offset >> 32) * 65537u;
    cudaStream_t st = (cudaStream_t)stream;
    { LaunchScope ls(st, kFamMisc); launch_k(sampler_kernel, ceil_div(rows, 128), 128, 0, st, q); }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}
