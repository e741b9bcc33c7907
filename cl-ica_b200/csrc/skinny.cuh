// skinny.cuh -- the encoder's first / last layers (n -> 10n and 10n -> n, main_mlp.py:297-309).
//
// One of their GEMM dimensions is n (5, 10, 40): a 128x128-tiled GEMM is pure latency on them (measured 30-40 us
// each at n = 10, M = 6144, i.e. 15% of the step) although they move < 5 MB and do < 0.1% of the flops.  Three
// small exact-fp32 kernels replace it; all read / write the tensor-core plane format ((hi, lo), value = hi + lo):
//   skinny_kin   reduction dim small (<= 48):  out[m, n] = epi( sum_k X[m, k] B(k, n) )   forward of layer 0,
//                backward-data of the last layer (mask + fused bias-gradient column sums)
//   skinny_nout  output dim small  (<= 48):   out[m, n] = act( sum_k X[m, k] W[n, k] + b[n] )   forward of the last layer
//   skinny_dw    dW[n, k] = sum_m dY[m, n] X[m, k],  db[n] = sum_m dY[m, n]   with N*K small (either layer)
#pragma once
#include "common.cuh"

namespace clica {

struct SkinnyKinParams {
    const float* x_hi; const float* x_lo; int ldx;     // [M, K]
    const float* B; long long b_sk, b_sn;              // B(k, n) = B[k*b_sk + n*b_sn]
    const float* bias;                                 // [N] or null (epi 0)
    const float* aux; int ldaux; float slope;          // mask source (epi 1) / LeakyReLU slope
    float* o_hi; float* o_lo; int ldo;                 // o_lo non-null: hi = tf32-round(v), lo = v - hi
    float* colsum;                                     // optional [N], pre-zeroed
    int M, N, K, epi;                                  // epi 0: bias + act, 1: mask
};

struct SkinnyNoutParams {
    const float* x_hi; const float* x_lo; int ldx;     // [M, K]
    const float* W; int ldw; const float* bias;        // [N, K]
    float* out; int ldo; float slope;
    int M, N, K;
};

struct SkinnyDwParams {
    const float* dy_hi; const float* dy_lo; int lddy;  // [M, N]
    const float* x_hi; const float* x_lo; int ldx;     // [M, K]
    float* dW; int lddw; float* db;                    // pre-zeroed; db nullable
    int M, N, K;
};

#ifdef __CUDACC__

constexpr int kSkRows = 32;

template <int KMAX>
__global__ void __launch_bounds__(256) skinny_kin_kernel(const SkinnyKinParams q) {
    __shared__ float xs[kSkRows][KMAX];
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * kSkRows;
    for (int idx = tid; idx < kSkRows * KMAX; idx += 256) {       // columns K..KMAX-1 are zero-filled
        const int r = idx / KMAX, k = idx - r * KMAX;
        const int m = row0 + r;
        float v = 0.f;
        if (m < q.M && k < q.K) {
            v = __ldg(q.x_hi + (size_t)m * q.ldx + k);
            if (q.x_lo) v += __ldg(q.x_lo + (size_t)m * q.ldx + k);
        }
        xs[r][k] = v;
    }
    __syncthreads();
    const int nrows = min(kSkRows, q.M - row0);
    for (int n = tid; n < q.N; n += 256) {
        float w[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) w[k] = (k < q.K) ? __ldg(q.B + k * q.b_sk + n * q.b_sn) : 0.f;
        const float bias_v = (q.epi == 0 && q.bias) ? __ldg(q.bias + n) : 0.f;
        float csum = 0.f;
        for (int r = 0; r < nrows; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) acc = fmaf(xs[r][k], w[k], acc);
            const size_t m = (size_t)(row0 + r);
            float v = acc;
            if (q.epi == 0) {
                v += bias_v;
                v = v > 0.f ? v : v * q.slope;
            } else if (q.aux) {
                v *= (__ldg(q.aux + m * q.ldaux + n) > 0.f) ? 1.f : q.slope;
            }
            csum += v;
            if (q.o_lo) {
                const float h = round_to_tf32(v);
                q.o_hi[m * q.ldo + n] = h;
                q.o_lo[m * q.ldo + n] = round_to_tf32(v - h);
            } else {
                q.o_hi[m * q.ldo + n] = v;
            }
        }
        if (q.colsum) atomicAdd(q.colsum + n, csum);
    }
}

// one warp per row: lanes split K (coalesced), N accumulators per lane, butterfly reduction at the end
template <int NMAX>
__global__ void __launch_bounds__(256) skinny_nout_kernel(const SkinnyNoutParams q) {
    extern __shared__ __align__(16) float Ws[];      // [N][K]
    for (int idx = threadIdx.x; idx < q.N * q.K; idx += 256) {
        const int n = idx / q.K, k = idx - n * q.K;
        Ws[idx] = __ldg(q.W + (size_t)n * q.ldw + k);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warps_total = gridDim.x * 8;
    for (int m = blockIdx.x * 8 + warp; m < q.M; m += warps_total) {
        float acc[NMAX];
#pragma unroll
        for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
        const float* xh = q.x_hi + (size_t)m * q.ldx;
        const float* xl = q.x_lo ? q.x_lo + (size_t)m * q.ldx : nullptr;
        for (int k = lane; k < q.K; k += 32) {
            float xv = __ldg(xh + k);
            if (xl) xv += __ldg(xl + k);
#pragma unroll
            for (int n = 0; n < NMAX; ++n)
                if (n < q.N) acc[n] = fmaf(xv, Ws[n * q.K + k], acc[n]);
        }
        float mine = 0.f;
#pragma unroll
        for (int n = 0; n < NMAX; ++n) {
            if (n < q.N) {                                   // q.N is warp-uniform
                const float t = warp_sum(acc[n]);
                if (lane == n) mine = t;
            }
        }
        if (lane < q.N) {
            float v = mine + (q.bias ? __ldg(q.bias + lane) : 0.f);
            v = v > 0.f ? v : v * q.slope;
            q.out[(size_t)m * q.ldo + lane] = v;
        }
        if (NMAX > 32) {                                     // columns 32.. of wide outputs
            float mine2 = 0.f;
#pragma unroll
            for (int n = 32; n < NMAX; ++n) {
                if (n < q.N) {
                    const float t = warp_sum(acc[n]);
                    if (lane == n - 32) mine2 = t;
                }
            }
            if (lane + 32 < q.N) {
                float v = mine2 + (q.bias ? __ldg(q.bias + lane + 32) : 0.f);
                v = v > 0.f ? v : v * q.slope;
                q.out[(size_t)m * q.ldo + lane + 32] = v;
            }
        }
    }
}

// each CTA reduces a block of rows held in shared memory; one thread per (n, k) output (k == K is the bias gradient)
static __global__ void __launch_bounds__(256) skinny_dw_kernel(const SkinnyDwParams q) {
    extern __shared__ __align__(16) float sm[];
    float* dys = sm;                                  // [kSkRows][N]
    float* xs = sm + kSkRows * q.N;                   // [kSkRows][K + 1]
    const int KK = q.K + 1;
    const int row0 = blockIdx.x * kSkRows;
    for (int idx = threadIdx.x; idx < kSkRows * q.N; idx += 256) {
        const int r = idx / q.N, n = idx - r * q.N;
        const int m = row0 + r;
        float v = 0.f;
        if (m < q.M) {
            v = __ldg(q.dy_hi + (size_t)m * q.lddy + n);
            if (q.dy_lo) v += __ldg(q.dy_lo + (size_t)m * q.lddy + n);
        }
        dys[idx] = v;
    }
    for (int idx = threadIdx.x; idx < kSkRows * KK; idx += 256) {
        const int r = idx / KK, k = idx - r * KK;
        const int m = row0 + r;
        float v = 0.f;
        if (m < q.M) {
            if (k < q.K) {
                v = __ldg(q.x_hi + (size_t)m * q.ldx + k);
                if (q.x_lo) v += __ldg(q.x_lo + (size_t)m * q.ldx + k);
            } else {
                v = 1.f;
            }
        }
        xs[idx] = v;
    }
    __syncthreads();
    const int total = q.N * KK;
    for (int o = threadIdx.x; o < total; o += 256) {
        const int n = o / KK, k = o - n * KK;
        if (k == q.K && q.db == nullptr) continue;
        float acc = 0.f;
#pragma unroll 8
        for (int r = 0; r < kSkRows; ++r) acc = fmaf(dys[r * q.N + n], xs[r * KK + k], acc);
        if (k < q.K) atomicAdd(q.dW + (size_t)n * q.lddw + k, acc);
        else atomicAdd(q.db + n, acc);
    }
}

#endif  // __CUDACC__
}  // namespace clica
