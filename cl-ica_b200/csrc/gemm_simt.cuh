// gemm_simt.cuh -- exact-fp32 CUDA-core GEMM used for the encoder's skinny layers (K or N = n, which TMA
// cannot address: a row of n = 10 floats is 40 bytes) and as the CLICA_GEMM_FP32 mode.
//
//   C[M x N] = epilogue( sum_k A(m,k) * B(k,n) )       A(m,k) = A[m*a_sm + k*a_sk],  B(k,n) = B[k*b_sk + n*b_sn]
//
// 128x128x8 tiles, 256 threads, 8x8 outputs per thread (two 4-wide groups 64 apart so that the 128-bit
// shared-memory reads are conflict-free), register-staged double buffering.  A_KC / B_KC say whether the
// operand is contiguous along the reduction dimension (selects the coalesced global-load mapping).
// Operands / results may be (hi, lo) plane pairs of the tensor-core path (value = hi + lo): the layers next
// to a tcgen05 layer read and write that format directly.
#pragma once
#include "common.cuh"

namespace clica {

enum SimtEpilogue : int {
    kEpiBiasAct = 0,    // C = act(acc + bias[n])            (forward)
    kEpiMask = 1,       // C = acc * (aux(m,n) > 0 ? 1 : slope)  (backward data; aux nullable -> no mask)
    kEpiAtomic = 2,     // atomicAdd(C, acc)                 (split-K backward weight; C pre-zeroed)
};

struct SimtGemmParams {
    const float* A; const float* A_lo; long long a_sm, a_sk;     // A_lo nullable (same strides)
    const float* B; const float* B_lo; long long b_sk, b_sn;     // B_lo nullable
    float* C; float* C_lo; int ldc;                              // C_lo non-null: C = tf32-round(v), C_lo = v - C
    int M, N, K;
    int k_chunk;             // reduction elements per blockIdx.z (split-K); == K when gridDim.z == 1
    const float* bias;       // [N] or null
    const float* aux; int ldaux;
    float slope;
    int epilogue;
    float* colsum;           // optional [N]: += column sums of the stored values (kEpiMask only; caller zeroes it)
};

#ifdef __CUDACC__

constexpr int kSBM = 128, kSBN = 128, kSBK = 8, kSPad = 4;

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256, 2) gemm_simt_kernel(const SimtGemmParams q) {
    pdl_enter();
    __shared__ __align__(16) float As[2][kSBK][kSBM + kSPad];
    __shared__ __align__(16) float Bs[2][kSBK][kSBN + kSPad];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.y * kSBM, n0 = blockIdx.x * kSBN;
    const int kbeg = blockIdx.z * q.k_chunk;
    const int kend = min(q.K, kbeg + q.k_chunk);

    // global -> register staging: 4 elements of A and 4 of B per thread per k-tile
    int a_row[4], a_k[4], b_col[4], b_k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + 256 * i;
        if (A_KC) { a_row[i] = idx >> 3; a_k[i] = idx & 7; } else { a_k[i] = idx >> 7; a_row[i] = idx & 127; }
        if (B_KC) { b_col[i] = idx >> 3; b_k[i] = idx & 7; } else { b_k[i] = idx >> 7; b_col[i] = idx & 127; }
    }
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + a_row[i], k = k0 + a_k[i];
            float va = 0.f;
            if (m < q.M && k < kend) {
                va = __ldg(q.A + m * q.a_sm + k * q.a_sk);
                if (q.A_lo) va += __ldg(q.A_lo + m * q.a_sm + k * q.a_sk);
            }
            ra[i] = va;
            const int n = n0 + b_col[i], kb = k0 + b_k[i];
            float vb = 0.f;
            if (n < q.N && kb < kend) {
                vb = __ldg(q.B + kb * q.b_sk + n * q.b_sn);
                if (q.B_lo) vb += __ldg(q.B_lo + kb * q.b_sk + n * q.b_sn);
            }
            rb[i] = vb;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            As[buf][a_k[i]][a_row[i]] = ra[i];
            Bs[buf][b_k[i]][b_col[i]] = rb[i];
        }
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    fetch(kbeg);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += kSBK) {
        const bool more = (k0 + kSBK < kend);
        if (more) fetch(k0 + kSBK);
#pragma unroll
        for (int k = 0; k < kSBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= q.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= q.N) continue;
            float v = acc[i][j];
            const size_t o = (size_t)m * q.ldc + n;
            if (q.epilogue == kEpiAtomic) {
                atomicAdd(q.C + o, v);
                continue;
            }
            if (q.epilogue == kEpiBiasAct) {
                if (q.bias) v += __ldg(q.bias + n);
                v = v > 0.f ? v : v * q.slope;
            } else if (q.aux) {
                v *= (__ldg(q.aux + (size_t)m * q.ldaux + n) > 0.f) ? 1.f : q.slope;
            }
            csum[j] += v;
            if (q.C_lo) {
                const float h = round_to_tf32(v);
                q.C[o] = h;
                q.C_lo[o] = round_to_tf32(v - h);
            } else {
                q.C[o] = v;
            }
        }
    }
    if (q.colsum != nullptr && q.epilogue == kEpiMask) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n < q.N) atomicAdd(q.colsum + n, csum[j]);
        }
    }
}

// db[n] += sum_m (hi[m, n] + lo[m, n]) over this block's row range; db pre-zeroed
static __global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ hi, const float* __restrict__ lo, int ld,
                                                      int M, int N, int rows_per_block, float* __restrict__ db) {
    pdl_enter();
    __shared__ float red[8][33];
    const int n = blockIdx.x * 32 + threadIdx.x;
    const int mbeg = blockIdx.y * rows_per_block, mend = min(M, mbeg + rows_per_block);
    float s = 0.f;
    if (n < N)
        for (int m = mbeg + threadIdx.y; m < mend; m += 8) {
            s += __ldg(hi + (size_t)m * ld + n);
            if (lo) s += __ldg(lo + (size_t)m * ld + n);
        }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
        atomicAdd(db + n, t);
    }
}

#endif  // __CUDACC__

// host launchers (defined in linear_api.cu); operands given as (hi, lo) pairs with lo nullable
int simt_gemm(const SimtGemmParams& q, bool a_kc, bool b_kc, int splits, cudaStream_t st);
int simt_colsum(const float* hi, const float* lo, int ld, int M, int N, float* db, cudaStream_t st, bool prezeroed = false);

}  // namespace clica
