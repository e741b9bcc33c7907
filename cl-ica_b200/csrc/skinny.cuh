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

// ---- skinny_kin: small reduction dim ------------------------------------------------------------------------
// CTA = 64 rows x 128 output columns; thread = (column group of 4, row slot of 8): float4 stores of whole
// 512-byte row segments per warp, the B(k, 4 columns) slice of a thread lives in registers (KMAX <= 16) or is read
// as one LDS.128 per k, x values are warp-broadcast shared-memory reads.
constexpr int kKinRows = 64;
constexpr int kKinCols = 128;

template <int KMAX>
__global__ void __launch_bounds__(256) skinny_kin_kernel(const SkinnyKinParams q) {
    pdl_enter();
    __shared__ __align__(16) float Bs[KMAX][kKinCols];
    __shared__ __align__(16) float xs[kKinRows][KMAX];
    __shared__ float cs[kKinCols];
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * kKinRows;
    const int n0 = blockIdx.y * kKinCols;
    // staging: every thread first issues ALL of its global loads (independent, in flight together), then stores --
    // a load -> store loop would pay one full memory latency per iteration
    {
        constexpr int NB = KMAX * kKinCols / 256;            // 8 (KMAX 16) or 24 (KMAX 48) elements per thread
        float v[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int idx = tid + 256 * u;
            const int k = idx / kKinCols, c = idx - k * kKinCols;
            const int n = n0 + c;
            v[u] = (k < q.K && n < q.N) ? __ldg(q.B + (long long)k * q.b_sk + (long long)n * q.b_sn) : 0.f;
        }
        constexpr int NX = kKinRows * KMAX / 256;            // 4 or 12
        float xh[NX], xl[NX];
#pragma unroll
        for (int u = 0; u < NX; ++u) {                       // columns K..KMAX-1 are zero-filled
            const int idx = tid + 256 * u;
            const int r = idx / KMAX, k = idx - r * KMAX;
            const int m = row0 + r;
            const bool ok = (m < q.M && k < q.K);
            xh[u] = ok ? __ldg(q.x_hi + (size_t)m * q.ldx + k) : 0.f;
            xl[u] = (ok && q.x_lo) ? __ldg(q.x_lo + (size_t)m * q.ldx + k) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int idx = tid + 256 * u;
            Bs[idx / kKinCols][idx % kKinCols] = v[u];
        }
#pragma unroll
        for (int u = 0; u < NX; ++u) {
            const int idx = tid + 256 * u;
            xs[idx / KMAX][idx % KMAX] = xh[u] + xl[u];
        }
    }
    if (tid < kKinCols) cs[tid] = 0.f;
    __syncthreads();
    const int cg = tid & 31, rs = tid >> 5;
    const int nb = n0 + 4 * cg;
    const int nrows = min(kKinRows, q.M - row0);
    if (nb < q.N) {
        const int nv = min(4, q.N - nb);                       // valid columns of this group
        const bool vec_o = (nv == 4) && ((q.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(q.o_hi) & 15u) == 0) &&
                           (q.o_lo == nullptr || (reinterpret_cast<uintptr_t>(q.o_lo) & 15u) == 0);
        const bool vec_a = (nv == 4) && q.aux && ((q.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(q.aux) & 15u) == 0);
        float bias_v[4] = {0.f, 0.f, 0.f, 0.f};
        if (q.epi == 0 && q.bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < nv) bias_v[j] = __ldg(q.bias + nb + j);
        }
        float4 breg[KMAX <= 16 ? KMAX : 1];
        if constexpr (KMAX <= 16) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) breg[k] = *reinterpret_cast<const float4*>(&Bs[k][4 * cg]);
        }
        float csum[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = rs; r < nrows; r += 8) {
            float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k4 = 0; k4 < KMAX; k4 += 4) {
                const float4 xv = *reinterpret_cast<const float4*>(&xs[r][k4]);
                const float xk[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float4 b;
                    if constexpr (KMAX <= 16) b = breg[k4 + kk];
                    else b = *reinterpret_cast<const float4*>(&Bs[k4 + kk][4 * cg]);
                    a[0] = fmaf(xk[kk], b.x, a[0]); a[1] = fmaf(xk[kk], b.y, a[1]);
                    a[2] = fmaf(xk[kk], b.z, a[2]); a[3] = fmaf(xk[kk], b.w, a[3]);
                }
            }
            const size_t m = (size_t)(row0 + r);
            if (q.epi == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { a[j] += bias_v[j]; a[j] = a[j] > 0.f ? a[j] : a[j] * q.slope; }
            } else if (q.aux) {
                float mk[4] = {1.f, 1.f, 1.f, 1.f};
                const float* ap = q.aux + m * q.ldaux + nb;
                if (vec_a) {
                    const float4 av = __ldg(reinterpret_cast<const float4*>(ap));
                    mk[0] = av.x; mk[1] = av.y; mk[2] = av.z; mk[3] = av.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < nv) mk[j] = __ldg(ap + j);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) a[j] *= (mk[j] > 0.f) ? 1.f : q.slope;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) csum[j] += a[j];
            float* oh = q.o_hi + m * q.ldo + nb;
            if (q.o_lo) {
                float* ol = q.o_lo + m * q.ldo + nb;
                float h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { h[j] = round_to_tf32(a[j]); l[j] = round_to_tf32(a[j] - h[j]); }
                if (vec_o) {
                    *reinterpret_cast<float4*>(oh) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4*>(ol) = make_float4(l[0], l[1], l[2], l[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (j < nv) { oh[j] = h[j]; ol[j] = l[j]; }
                }
            } else if (vec_o) {
                *reinterpret_cast<float4*>(oh) = make_float4(a[0], a[1], a[2], a[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j < nv) oh[j] = a[j];
            }
        }
        if (q.colsum) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < nv) atomicAdd(&cs[4 * cg + j], csum[j]);
        }
    }
    if (q.colsum) {
        __syncthreads();
        if (tid < kKinCols && n0 + tid < q.N) atomicAdd(q.colsum + n0 + tid, cs[tid]);
    }
}

// ---- skinny_nout_small: output dim <= 16 ---------------------------------------------------------------------
// CTA = 64 rows; X (hi + lo summed) and W are staged in shared memory in K-chunks of 128 with coalesced loads,
// thread = one (row, n) output per item (n fastest: coalesced stores), up to 4 items per thread.
constexpr int kNsRows = 64;
constexpr int kNsKC = 128;
constexpr int kNsMaxN = 16;
constexpr int kNsItems = kNsRows * kNsMaxN / 256;
inline size_t nout_small_smem_bytes(int N) { return (size_t)(kNsRows + N) * (kNsKC + 1) * sizeof(float); }

static __global__ void __launch_bounds__(256) skinny_nout_small_kernel(const SkinnyNoutParams q) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];
    constexpr int LD = kNsKC + 1;
    float* xs = sm;                       // [kNsRows][LD]
    float* ws = sm + kNsRows * LD;        // [N][LD]
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * kNsRows;
    const int nrows = min(kNsRows, q.M - row0);
    const int items = nrows * q.N;
    float acc[kNsItems];
#pragma unroll
    for (int i = 0; i < kNsItems; ++i) acc[i] = 0.f;
    for (int k0 = 0; k0 < q.K; k0 += kNsKC) {
        const int kc = min(kNsKC, q.K - k0);
        if (k0 > 0) __syncthreads();
        {   // all global loads of the chunk first (independent), then the shared-memory stores
            constexpr int NV = kNsRows * kNsKC / 4 / 256;    // 8 float4 per thread
            const bool vec = ((q.ldx & 3) == 0) && ((k0 & 3) == 0) &&
                             ((reinterpret_cast<uintptr_t>(q.x_hi) & 15u) == 0) &&
                             (q.x_lo == nullptr || (reinterpret_cast<uintptr_t>(q.x_lo) & 15u) == 0);
            float4 vh[NV], vl[NV];
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int idx = tid + 256 * u;
                const int r = idx / (kNsKC / 4), k = 4 * (idx - r * (kNsKC / 4));
                vh[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                vl[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nrows && k < kc) {
                    const size_t off = (size_t)(row0 + r) * q.ldx + k0 + k;
                    if (vec && k + 3 < kc) {
                        vh[u] = __ldg(reinterpret_cast<const float4*>(q.x_hi + off));
                        if (q.x_lo) vl[u] = __ldg(reinterpret_cast<const float4*>(q.x_lo + off));
                    } else {
                        vh[u].x = __ldg(q.x_hi + off);
                        if (k + 1 < kc) vh[u].y = __ldg(q.x_hi + off + 1);
                        if (k + 2 < kc) vh[u].z = __ldg(q.x_hi + off + 2);
                        if (k + 3 < kc) vh[u].w = __ldg(q.x_hi + off + 3);
                        if (q.x_lo) {
                            vl[u].x = __ldg(q.x_lo + off);
                            if (k + 1 < kc) vl[u].y = __ldg(q.x_lo + off + 1);
                            if (k + 2 < kc) vl[u].z = __ldg(q.x_lo + off + 2);
                            if (k + 3 < kc) vl[u].w = __ldg(q.x_lo + off + 3);
                        }
                    }
                }
            }
            constexpr int NW = (kNsMaxN * kNsKC + 255) / 256;  // 8 weights per thread
            float wv[NW];
#pragma unroll
            for (int u = 0; u < NW; ++u) {
                const int idx = tid + 256 * u;
                const int n = idx / kNsKC, k = idx - n * kNsKC;
                wv[u] = (n < q.N && k < kc) ? __ldg(q.W + (size_t)n * q.ldw + k0 + k) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < NV; ++u) {
                const int idx = tid + 256 * u;
                const int r = idx / (kNsKC / 4), k = 4 * (idx - r * (kNsKC / 4));
                float* d = xs + r * LD + k;
                d[0] = vh[u].x + vl[u].x; d[1] = vh[u].y + vl[u].y; d[2] = vh[u].z + vl[u].z; d[3] = vh[u].w + vl[u].w;
            }
#pragma unroll
            for (int u = 0; u < NW; ++u) {
                const int idx = tid + 256 * u;
                const int n = idx / kNsKC, k = idx - n * kNsKC;
                if (n < q.N) ws[n * LD + k] = wv[u];
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kNsItems; ++i) {
            const int item = tid + 256 * i;
            if (item < items) {
                const int r = item / q.N, n = item - r * q.N;
                const float* xr = xs + r * LD;
                const float* wr = ws + n * LD;
                float a0 = acc[i], a1 = 0.f;
                int k = 0;
                for (; k + 1 < kc; k += 2) { a0 = fmaf(xr[k], wr[k], a0); a1 = fmaf(xr[k + 1], wr[k + 1], a1); }
                if (k < kc) a0 = fmaf(xr[k], wr[k], a0);
                acc[i] = a0 + a1;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kNsItems; ++i) {
        const int item = tid + 256 * i;
        if (item < items) {
            const int r = item / q.N, n = item - r * q.N;
            float v = acc[i] + (q.bias ? __ldg(q.bias + n) : 0.f);
            v = v > 0.f ? v : v * q.slope;
            q.out[(size_t)(row0 + r) * q.ldo + n] = v;
        }
    }
}

// ---- skinny_dw_small: N * (K + 1) <= 2048 outputs ------------------------------------------------------------
// Persistent over 32-row tiles: every thread keeps up to 8 outputs (n, k) in registers while its CTA walks its
// tiles, then issues ONE atomic per output per CTA (the tile-at-a-time kernel below issues one per output per tile).
constexpr int kDwsMaxOut = 2048;
constexpr int kDwsItems = kDwsMaxOut / 256;

static __global__ void __launch_bounds__(256) skinny_dw_small_kernel(const SkinnyDwParams q) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];
    float* dys = sm;                                  // [kSkRows][N]
    float* xs = sm + kSkRows * q.N;                   // [kSkRows][K + 1]
    const int KK = q.K + 1;
    const int total = q.N * KK;
    float acc[kDwsItems];
#pragma unroll
    for (int i = 0; i < kDwsItems; ++i) acc[i] = 0.f;
    const int ntiles = (q.M + kSkRows - 1) / kSkRows;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int row0 = t * kSkRows;
        if (t != (int)blockIdx.x) __syncthreads();
        // batches of 8 independent loads per thread (hi and lo), then the shared-memory stores
        for (int base = threadIdx.x; base < kSkRows * q.N; base += 256 * 8) {
            float vh[8], vl[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + 256 * u;
                const int r = idx / q.N, n = idx - r * q.N;
                const int m = row0 + r;
                const bool ok = (idx < kSkRows * q.N) && (m < q.M);
                vh[u] = ok ? __ldg(q.dy_hi + (size_t)m * q.lddy + n) : 0.f;
                vl[u] = (ok && q.dy_lo) ? __ldg(q.dy_lo + (size_t)m * q.lddy + n) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + 256 * u;
                if (idx < kSkRows * q.N) dys[idx] = vh[u] + vl[u];
            }
        }
        for (int base = threadIdx.x; base < kSkRows * KK; base += 256 * 8) {
            float vh[8], vl[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + 256 * u;
                const int r = idx / KK, k = idx - r * KK;
                const int m = row0 + r;
                const bool ok = (idx < kSkRows * KK) && (m < q.M);
                vh[u] = !ok ? 0.f : (k < q.K ? __ldg(q.x_hi + (size_t)m * q.ldx + k) : 1.f);
                vl[u] = (ok && k < q.K && q.x_lo) ? __ldg(q.x_lo + (size_t)m * q.ldx + k) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + 256 * u;
                if (idx < kSkRows * KK) xs[idx] = vh[u] + vl[u];
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kDwsItems; ++i) {
            const int o = threadIdx.x + 256 * i;
            if (o < total) {
                const int n = o / KK, k = o - n * KK;
                float a = acc[i];
#pragma unroll 8
                for (int r = 0; r < kSkRows; ++r) a = fmaf(dys[r * q.N + n], xs[r * KK + k], a);
                acc[i] = a;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kDwsItems; ++i) {
        const int o = threadIdx.x + 256 * i;
        if (o < total) {
            const int n = o / KK, k = o - n * KK;
            if (k < q.K) atomicAdd(q.dW + (size_t)n * q.lddw + k, acc[i]);
            else if (q.db) atomicAdd(q.db + n, acc[i]);
        }
    }
}

// one warp per row: lanes split K (coalesced), N accumulators per lane, butterfly reduction at the end
template <int NMAX>
__global__ void __launch_bounds__(256) skinny_nout_kernel(const SkinnyNoutParams q) {
    pdl_enter();
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
    pdl_enter();
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
