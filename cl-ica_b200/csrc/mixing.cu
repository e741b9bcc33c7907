// mixing.cu -- the frozen mixing network g of h = f o g (main_mlp.py:313), forward only.
//
// Replaces, for CUDA fp32 inputs, the nn.Sequential built by /root/reference/invertible_network_utils.py:87-123
// (construct_invertible_mlp: L bias-free n x n Linear layers with LeakyReLU(slope) in between, weights frozen).
// In torch that is L cuBLAS SGEMMs + (L-1) elementwise kernels (~20 us of launches at n = 10, M = 12288 for
// 3 n^2 MAC per row); here one thread carries one row through all layers in registers, the L weight matrices sit in
// shared memory and are read as warp broadcasts.  Exact fp32 (FFMA), same summation order as a row-major dot product.
#include "common.cuh"

namespace clica {
namespace {

constexpr int kMixMaxLayers = 8;

struct MixParams {
    const float* x; int ldx;
    float* y; int ldy;
    const float* W[kMixMaxLayers];     // W[l] is [n, n] row-major (nn.Linear weight: out x in), contiguous
    int L, n, M;
    float slope;
};

template <int NMAX>
__global__ void __launch_bounds__(128) mixing_fwd_kernel(const MixParams q) {
    pdl_enter();
    extern __shared__ __align__(16) float Ws[];      // [L][n][n]
    const int nn = q.n * q.n;
    for (int l = 0; l < q.L; ++l)
        for (int i = threadIdx.x; i < nn; i += blockDim.x) Ws[l * nn + i] = __ldg(q.W[l] + i);
    __syncthreads();
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= q.M) return;
    float v[NMAX];
#pragma unroll
    for (int k = 0; k < NMAX; ++k) v[k] = (k < q.n) ? __ldg(q.x + (size_t)row * q.ldx + k) : 0.f;
    for (int l = 0; l < q.L; ++l) {
        const float* Wl = Ws + l * nn;
        float o[NMAX];
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            float a = 0.f;
            if (j < q.n) {
#pragma unroll
                for (int k = 0; k < NMAX; ++k)
                    if (k < q.n) a = fmaf(Wl[j * q.n + k], v[k], a);
            }
            o[j] = a;
        }
        const bool act = (l != q.L - 1);
#pragma unroll
        for (int j = 0; j < NMAX; ++j) v[j] = (act && o[j] < 0.f) ? o[j] * q.slope : o[j];
    }
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
        if (j < q.n) q.y[(size_t)row * q.ldy + j] = v[j];
}

}  // namespace
}  // namespace clica

using namespace clica;

extern "C" int clica_mixing_fwd(const float* x, int ldx, const float* const* W, int L, int n, int M, float slope,
                                float* y, int ldy, void* stream) {
    CLICA_REQUIRE(x && W && y, CLICA_E_BADARG, "mixing_fwd: null pointer");
    CLICA_REQUIRE(L >= 1 && L <= kMixMaxLayers, CLICA_E_UNSUPPORTED, "mixing_fwd: %d layers (supported: 1..%d)", L, kMixMaxLayers);
    CLICA_REQUIRE(n >= 1 && n <= 48, CLICA_E_UNSUPPORTED, "mixing_fwd: width %d (supported: 1..48)", n);
    CLICA_REQUIRE(M >= 1 && ldx >= n && ldy >= n, CLICA_E_BADARG, "mixing_fwd: bad shape M=%d n=%d ldx=%d ldy=%d", M, n, ldx, ldy);
    const size_t smem = (size_t)L * n * n * sizeof(float);
    CLICA_REQUIRE(smem <= 48 * 1024, CLICA_E_UNSUPPORTED, "mixing_fwd: %d layers of %d x %d weights exceed 48 KB of shared memory", L, n, n);
    DeviceInfo di;
    int rc = get_device_info(&di);
    if (rc) return rc;
    MixParams q = {};
    q.x = x; q.ldx = ldx; q.y = y; q.ldy = ldy; q.L = L; q.n = n; q.M = M; q.slope = slope;
    for (int l = 0; l < L; ++l) {
        CLICA_REQUIRE(W[l] != nullptr, CLICA_E_BADARG, "mixing_fwd: weight %d is null", l);
        q.W[l] = W[l];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ceil_div(M, 128);
    {
        LaunchScope ls(st, kFamGemmSimt);
        if (n <= 16) launch_k(mixing_fwd_kernel<16>, grid, 128, smem, st, q);
        else launch_k(mixing_fwd_kernel<48>, grid, 128, smem, st, q);
    }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}
