// lpnce_kernels.cuh -- fused pairwise-Lp-distance + InfoNCE (soft-max cross-entropy) kernels.
//
// Replaces /root/reference/losses.py:443-477 (+ :506-510) and the autograd graph behind it.
// The reference materialises B x M x d and B x M temporaries; here neither exists:
//
//   "row-owner" scheme -- every thread OWNS R rows of one operand (their features live in registers,
//   pre-negated), all threads of a CTA walk the rows of the other operand, which are STREAMED through
//   shared memory in cp.async double-buffered tiles and read with warp-broadcast 128-bit loads.  A pair
//   (owner row, streamed row) is handled by one thread (or by F adjacent lanes that each own a d/F feature
//   slice when d > 40; only the distance then needs log2(F) shuffles):
//     forward : D = sum_c |s_c - o_c|^p  ->  online log-sum-exp (lazy re-scaling, logits <= 0)
//     backward: w = exp2((-D*coef - m2) - ls) ->  g_owner += w * d|t|^p/dt      (distance recomputed)
//   so there are no atomics and no cross-thread reductions of gradients in the inner loops.  The general
//   backward runs the kernel in two roles (anchors own / negatives own); when the negatives are a row
//   permutation of the anchors (torch.roll in main_mlp.py:272, and always in the row-sharded multi-GPU form)
//   the distance is symmetric and ONE merged pass accumulates both roles.
//   Feature PAIRS are processed with the sm_100 packed-fp32 instructions (add.f32x2 / fma.rn.f32x2 via
//   __fadd2_rn / __ffma2_rn): half the issue slots per pair-element.
//   Column splits (grid.y) fill all SMs; split partials are merged by a small finalize / reduce kernel
//   in a fixed order, so results are deterministic.
#pragma once
#include "common.cuh"

#include <math.h>

namespace clica {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTN = 128;          // streamed rows per shared-memory tile (divided by the feature-split factor F)
constexpr int kCW = 4;            // warps of a CTA that split the streamed rows of a tile
constexpr float kRescale = 24.f;  // lazy soft-max re-scaling threshold (log2 units)

// rows owned per thread: 2 while the register budget allows it
// (the EXPERIMENTAL fast forward owns 4 rows per thread for d <= 10: its inner loop has no per-pair soft-max
// bookkeeping left to hide the broadcast LDS and the loop control behind)
constexpr int fwd_rows_per_thread(int DP, bool fast = false) { return (fast && DP <= 5) ? 4 : (DP <= 12 ? 2 : 1); }
constexpr int bwd_rows_per_thread(int DP) { return DP <= 5 ? 2 : 1; }
constexpr int rows_per_cta(int R, int F = 1) { return (32 / F) * R * (kWarps / kCW); }
// shared-memory layout of a streamed row: F feature slices of 2*DP floats (+4 floats of padding when F > 1 so
// that the F lanes of a pair hit different bank groups)
constexpr int slice_floats(int DP, int F) { return 2 * DP + (F > 1 ? 4 : 0); }
constexpr int row_floats(int DP, int F) { return F * slice_floats(DP, F); }
constexpr int tile_rows(int F) { return kTN / F; }

inline size_t fwd_smem_bytes(int DP, int R, int F = 1) {
    size_t tiles = 2ull * tile_rows(F) * row_floats(DP, F) * sizeof(float);
    size_t merge = 2ull * kWarps * R * 32 * sizeof(float);
    return tiles > merge ? tiles : merge;
}
inline size_t bwd_smem_bytes(int DP, int R, int F = 1) {
    size_t tiles = 2ull * (tile_rows(F) * row_floats(DP, F) + 4 * tile_rows(F)) * sizeof(float);
    size_t merge = (size_t)(kCW - 1) * (kWarps / kCW) * R * DP * 32 * sizeof(float2);
    return tiles > merge ? tiles : merge;
}

struct FwdParams {
    const float* O; int ldO; int BO;     // owner rows (anchors z1)
    const float* S; int ldS; int MS;     // streamed rows (negatives z3)
    int d; float coef; float pg;         // coef = log2(e)/tau, pg = p (generic-exponent kernels)
    int tiles_per_split; int flat16;
    float* part_m; float* part_s; int part_stride;   // [nsplit][part_stride]
    int* counter;                        // zeroed here for the finalize kernel's last-block reduction
    int fast;                            // EXPERIMENTAL: fixed reference point 0 (see lpnce_fwd_kernel, FAST)
};

// Row statistics of the forward, per anchor: (m2, ls) = (reference maximum of the log2-domain logits,
// log2 of the sum of exp2(logit - m2)); the soft-max weight of a pair is exp2((-D*coef - m2) - ls), i.e. the
// SAME arithmetic the forward used, so the weights of a row sum to 1 to fp32 accuracy however large |lse| is.
// w(owner i, streamed j) = ES[j] * exp2( (-D*coef - m2) - ls ), (m2, ls) taken from the owner (anchor role)
// or from the streamed row (column role);  gacc_i += w * G'(s_j - o_i)/p
struct BwdRole {
    const float* O; int ldO; int BO;
    const float* S; int ldS; int MS;
    const float2* LO;                   // owner row statistics (nullable -> (0, 0))
    const float2* LS; const float* ES;  // streamed row statistics and coefficient (both or neither)
    const float* EO;                    // MERGED mode (owner set == streamed set, e.g. z3 = roll(z1)): owner coefficient;
                                        // distance is symmetric, so one pass accumulates both roles:
                                        // w = EO[i] exp2(.. - stat_i) + ES[j] exp2(.. - stat_j)
    int tiles_per_split; int nsplit; int flat16;
    float* part; int part_rows;        // [nsplit][part_rows][F*2*DP]
    int row_tiles;
};
struct BwdParams {
    BwdRole role[2];
    int nroles;
    int d; float coef; float pg;
};

#ifdef __CUDACC__

// ---- per-exponent arithmetic ----------------------------------------------------------------------
// P = 1,2,3,4: multiply-only; P = 0: generic real exponent through lg2/ex2 (pg = p).
template <int P>
struct Lp {
    // acc += |t|^p for the two features packed in t
    static __device__ __forceinline__ float2 accum(float2 t, float2 acc, float pg) {
        if constexpr (P == 2) {
            return __ffma2_rn(t, t, acc);
        } else if constexpr (P == 1) {
            acc.x += fabsf(t.x);
            acc.y += fabsf(t.y);
            return acc;
        } else if constexpr (P == 3) {
            float2 u = __fmul2_rn(t, t);
            acc.x = fmaf(fabsf(t.x), u.x, acc.x);
            acc.y = fmaf(fabsf(t.y), u.y, acc.y);
            return acc;
        } else if constexpr (P == 4) {
            float2 u = __fmul2_rn(t, t);
            return __ffma2_rn(u, u, acc);
        } else {
            acc.x += ex2_approx(pg * lg2_approx(fabsf(t.x)));   // |0|^p: lg2 -> -inf -> ex2 -> 0
            acc.y += ex2_approx(pg * lg2_approx(fabsf(t.y)));
            return acc;
        }
    }
    // g += w * sign(t)|t|^(p-1)     (the factor p is applied once, by the reduce kernel)
    static __device__ __forceinline__ float2 grad(float w, float2 t, float2 g, float pg) {
        if constexpr (P == 2) {
            return __ffma2_rn(make_float2(w, w), t, g);
        } else if constexpr (P == 1) {
            // w * sign(t), sign(0) = 0: a zero difference contributes exactly nothing (torch masks it too).  w keeps
            // its own sign (negative upstream gradients, alpha > 1): copysignf(w, t) would drop it.
            g.x += (t.x > 0.f) ? w : ((t.x < 0.f) ? -w : 0.f);
            g.y += (t.y > 0.f) ? w : ((t.y < 0.f) ? -w : 0.f);
            return g;
        } else if constexpr (P == 3) {
            g.x = fmaf(w, t.x * fabsf(t.x), g.x);
            g.y = fmaf(w, t.y * fabsf(t.y), g.y);
            return g;
        } else if constexpr (P == 4) {
            float2 u = __fmul2_rn(t, t);
            return __ffma2_rn(make_float2(w, w), __fmul2_rn(u, t), g);
        } else {
            float ax = ex2_approx((pg - 1.f) * lg2_approx(fabsf(t.x)));
            float ay = ex2_approx((pg - 1.f) * lg2_approx(fabsf(t.y)));
            g.x = fmaf(w, copysignf(ax, t.x), g.x);   // p > 1 here: |0|^(p-1) = 0
            g.y = fmaf(w, copysignf(ay, t.y), g.y);
            return g;
        }
    }
};

// sum over the F adjacent lanes that share a pair (F is a power of two <= 8)
template <int F>
__device__ __forceinline__ float slice_sum(float v) {
#pragma unroll
    for (int o = 1; o < F; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- streamed-tile loader -------------------------------------------------------------------------
// dst: [kTN / F][F][2*DP (+4)] floats.  Rows >= MS and features >= d are zero-filled.
template <int DP, int F>
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ S, int ldS, int MS,
                                          int d, int k0, int flat16, int tid) {
    constexpr int W = 2 * DP * F;            // real features per row (padded to the slice grid)
    constexpr int SLW = slice_floats(DP, F);
    constexpr int TW = row_floats(DP, F);
    constexpr int TNF = tile_rows(F);
    if (F == 1 && flat16) {   // d == W, ldS == d, 16B-aligned base: the tile is one contiguous chunk
        constexpr int G = TNF * W / 4;
        const long long total = (long long)MS * d;
        for (int g = tid; g < G; g += kThreads) {
            long long e = (long long)k0 * d + 4ll * g;
            long long rem = total - e;
            int bytes = rem >= 4 ? 16 : (rem > 0 ? (int)rem * 4 : 0);
            cp_async_16(dst + 4 * g, bytes > 0 ? (const void*)(S + e) : (const void*)S, bytes);
        }
    } else {
        for (int e = tid; e < TNF * W; e += kThreads) {
            const int k = e / W, c = e - k * W;
            const int f = c / (2 * DP), cc = c - f * (2 * DP);
            const bool ok = (k0 + k < MS) && (c < d);
            cp_async_4(dst + k * TW + f * SLW + cc, ok ? (const void*)(S + (size_t)(k0 + k) * ldS + c) : (const void*)S, ok ? 4 : 0);
        }
    }
}

// own rows (this lane's feature slice), negated and zero-padded, into registers
template <int DP, int R, int F>
__device__ __forceinline__ void load_owner_rows(float2 (&na)[R][DP], const float* __restrict__ O, int ldO,
                                                int BO, int d, int row_base, int lane) {
    const int fs = lane % F, lr = lane / F;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row_base + r * (32 / F) + lr;
#pragma unroll
        for (int c = 0; c < DP; ++c) {
            const int f0 = fs * 2 * DP + 2 * c;
            float v0 = 0.f, v1 = 0.f;
            if (row < BO) {
                if (f0 < d) v0 = __ldg(O + (size_t)row * ldO + f0);
                if (f0 + 1 < d) v1 = __ldg(O + (size_t)row * ldO + f0 + 1);
            }
            na[r][c] = make_float2(-v0, -v1);
        }
    }
}

// feature slices of streamed rows kk and kk+1 of a tile -> bb[0..DP) and bb[DP..2DP)
template <int DP, int F>
__device__ __forceinline__ void load_pair_rows(float2 (&bb)[2 * DP], const float* tile, int kk, int fs) {
    constexpr int SLW = slice_floats(DP, F);
    constexpr int TW = row_floats(DP, F);
    if constexpr (F == 1) {
        // rows kk, kk+1 are contiguous: 2*TW floats = DP float4 (kk even -> 16B aligned)
        const float4* bq = reinterpret_cast<const float4*>(tile + kk * TW);
#pragma unroll
        for (int c = 0; c < DP; ++c) {
            const float4 v = bq[c];
            bb[2 * c] = make_float2(v.x, v.y);
            bb[2 * c + 1] = make_float2(v.z, v.w);
        }
    } else {
        static_assert(DP % 2 == 0, "feature-split kernels need an even number of feature pairs");
        const float4* b0 = reinterpret_cast<const float4*>(tile + kk * TW + fs * SLW);
        const float4* b1 = reinterpret_cast<const float4*>(tile + (kk + 1) * TW + fs * SLW);
#pragma unroll
        for (int c = 0; c < DP / 2; ++c) {
            const float4 u = b0[c], v = b1[c];
            bb[2 * c] = make_float2(u.x, u.y); bb[2 * c + 1] = make_float2(u.z, u.w);
            bb[DP + 2 * c] = make_float2(v.x, v.y); bb[DP + 2 * c + 1] = make_float2(v.z, v.w);
        }
    }
}

// ================================ forward ==========================================================
// FAST (EXPERIMENTAL, off unless CLICA_LPNCE_FAST=1; written after round 1's GPU budget was spent): every logit is
// -D*coef <= 0, so exp2 can never overflow and the running reference point can simply be 0: no per-pair maximum, vote
// or re-scaling (about a quarter of the instructions of the inner loop at d = 10).  What is lost is protection
// against UNDERflow of a whole row (all logits < -126); the finalize kernel detects such rows (sum < 2^-80, where
// flushed terms could matter) and recomputes them with a robust two-pass loop.
template <int P, int DP, int R, int CW, int F, bool FAST>
__global__ void __launch_bounds__(kThreads) lpnce_fwd_kernel(const FwdParams q) {
    constexpr int RW = kWarps / CW;
    constexpr int RPW = 32 / F;                // rows per warp per r
    constexpr int ROWS = RPW * R * RW;
    constexpr int TW = row_floats(DP, F);      // floats per streamed row in shared memory
    constexpr int TNF = tile_rows(F);
    constexpr int CPW = TNF / CW;              // streamed rows per warp per tile
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rw = warp / CW, cw = warp % CW;
    const int fs = lane % F, lr = lane / F;
    const int row_base = blockIdx.x * ROWS + rw * (RPW * R);
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *q.counter = 0;

    float2 na[R][DP];
    load_owner_rows<DP, R, F>(na, q.O, q.ldO, q.BO, q.d, row_base, lane);

    const int ntiles = (q.MS + TNF - 1) / TNF;
    const int t0 = blockIdx.y * q.tiles_per_split;
    const int t1 = min(ntiles, t0 + q.tiles_per_split);
    load_tile<DP, F>(smem, q.S, q.ldS, q.MS, q.d, t0 * TNF, q.flat16, tid);
    cp_async_commit();

    float m[R], s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { m[r] = 0.f; s[r] = 0.f; }

    for (int t = t0; t < t1; ++t) {
        const int stage = (t - t0) & 1;
        if (t + 1 < t1) {
            load_tile<DP, F>(smem + (stage ^ 1) * TNF * TW, q.S, q.ldS, q.MS, q.d, (t + 1) * TNF, q.flat16, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* tile = smem + stage * TNF * TW;
        const int nvalid = min(TNF, q.MS - t * TNF);
        if (!FAST && t == t0) {
            // reference point of the lazy soft-max: the logit of the first streamed row of this split
            const float2* b = reinterpret_cast<const float2*>(tile + fs * slice_floats(DP, F));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < DP; ++c) acc = Lp<P>::accum(__fadd2_rn(b[c], na[r][c]), acc, q.pg);
                m[r] = -slice_sum<F>(acc.x + acc.y) * q.coef;
            }
        }
        const int c_end = min(cw * CPW + CPW, nvalid);
        for (int kk = cw * CPW; kk < c_end; kk += 2) {
            float2 bb[2 * DP];
            load_pair_rows<DP, F>(bb, tile, kk, fs);
            const bool has1 = (kk + 1 < c_end);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < DP; ++c) {
                    a0 = Lp<P>::accum(__fadd2_rn(bb[c], na[r][c]), a0, q.pg);
                    a1 = Lp<P>::accum(__fadd2_rn(bb[DP + c], na[r][c]), a1, q.pg);
                }
                const float D0 = slice_sum<F>(a0.x + a0.y);
                const float D1s = slice_sum<F>(a1.x + a1.y);
                if constexpr (FAST) {
                    const float e0 = ex2_approx(D0 * -q.coef);
                    const float e1 = ex2_approx(has1 ? D1s * -q.coef : -INFINITY);   // ex2(-inf) = +0, branch-free
                    s[r] += e0 + e1;
                    continue;
                }
                const float D1 = has1 ? D1s : INFINITY;
                float x0 = fmaf(D0, -q.coef, -m[r]);
                float x1 = fmaf(D1, -q.coef, -m[r]);
                const float hi = fmaxf(x0, x1);
                if (__any_sync(0xffffffffu, hi > kRescale)) {   // warp-uniform branch; rare: a much closer negative
                    if (hi > kRescale) {                        // than the reference point appeared
                        const float mn = m[r] + hi;
                        s[r] *= ex2_approx(m[r] - mn);
                        m[r] = mn;
                        x0 = fmaf(D0, -q.coef, -mn);
                        x1 = fmaf(D1, -q.coef, -mn);
                    }
                }
                s[r] += ex2_approx(x0) + ex2_approx(x1);
            }
        }
        __syncthreads();
    }

    // merge the CW column-warps that share a row group, then publish this split's partial (m, s)
    float* mm = smem;                       // [kWarps][R][32]
    float* ms = smem + kWarps * R * 32;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        mm[(warp * R + r) * 32 + lane] = m[r];
        ms[(warp * R + r) * 32 + lane] = s[r];
    }
    __syncthreads();
    if (cw == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float M = m[r];
#pragma unroll
            for (int w2 = 1; w2 < CW; ++w2) M = fmaxf(M, mm[((warp + w2) * R + r) * 32 + lane]);
            float Ssum = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < CW; ++w2) {
                const int idx = ((warp + w2) * R + r) * 32 + lane;
                Ssum += ms[idx] * exp2f(mm[idx] - M);
            }
            const int row = row_base + r * RPW + lr;
            if (row < q.BO && fs == 0) {
                q.part_m[(size_t)blockIdx.y * q.part_stride + row] = M;
                q.part_s[(size_t)blockIdx.y * q.part_stride + row] = Ssum;
            }
        }
    }
}

// ================================ backward =========================================================
template <int P, int DP, int R, int CW, int F>
__global__ void __launch_bounds__(kThreads) lpnce_bwd_kernel(const BwdParams q) {
    constexpr int RW = kWarps / CW;
    constexpr int RPW = 32 / F;
    constexpr int ROWS = RPW * R * RW;
    constexpr int TW = row_floats(DP, F);
    constexpr int TNF = tile_rows(F);
    constexpr int CPW = TNF / CW;
    constexpr int PTW = F * 2 * DP;                    // floats per row of the partial-gradient buffer
    constexpr int TILE_FLOATS = TNF * TW + 4 * TNF;    // features + (m2, ls, ES, pad) per streamed row
    extern __shared__ __align__(16) float smem[];
    const BwdRole& ro = q.role[blockIdx.z];
    if ((int)blockIdx.x >= ro.row_tiles || (int)blockIdx.y >= ro.nsplit) return;
    const bool has_ss = (ro.LS != nullptr);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rw = warp / CW, cw = warp % CW;
    const int fs = lane % F, lr = lane / F;
    const int row_base = blockIdx.x * ROWS + rw * (RPW * R);

    float2 na[R][DP];
    float2 gacc[R][DP];
    float lo_m[R], lo_s[R], eo[R];
    const bool merged = (ro.EO != nullptr);
    load_owner_rows<DP, R, F>(na, ro.O, ro.ldO, ro.BO, q.d, row_base, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row_base + r * RPW + lr;
#pragma unroll
        for (int c = 0; c < DP; ++c) gacc[r][c] = make_float2(0.f, 0.f);
        const float2 st = (ro.LO != nullptr && row < ro.BO) ? __ldg(ro.LO + row) : make_float2(0.f, 0.f);
        lo_m[r] = st.x; lo_s[r] = st.y;
        eo[r] = (merged && row < ro.BO) ? __ldg(ro.EO + row) : 0.f;
    }

    const int ntiles = (ro.MS + TNF - 1) / TNF;
    const int t0 = blockIdx.y * ro.tiles_per_split;
    const int t1 = min(ntiles, t0 + ro.tiles_per_split);

    auto issue_tile = [&](int stage, int t) {
        float* dst = smem + stage * TILE_FLOATS;
        load_tile<DP, F>(dst, ro.S, ro.ldS, ro.MS, q.d, t * TNF, ro.flat16, tid);
        if (has_ss && tid < TNF) {
            const int j = t * TNF + tid;
            const bool ok = j < ro.MS;
            float* sdst = dst + TNF * TW + 4 * tid;
            const float* stat = reinterpret_cast<const float*>(ro.LS);
            cp_async_4(sdst, ok ? (const void*)(stat + 2 * j) : (const void*)stat, ok ? 4 : 0);
            cp_async_4(sdst + 1, ok ? (const void*)(stat + 2 * j + 1) : (const void*)stat, ok ? 4 : 0);
            cp_async_4(sdst + 2, ok ? (const void*)(ro.ES + j) : (const void*)ro.ES, ok ? 4 : 0);   // 0 for tail rows
        }
        cp_async_commit();
    };
    issue_tile(0, t0);

    for (int t = t0; t < t1; ++t) {
        const int stage = (t - t0) & 1;
        if (t + 1 < t1) { issue_tile(stage ^ 1, t + 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float* tile = smem + stage * TILE_FLOATS;
        const float4* ss = reinterpret_cast<const float4*>(tile + TNF * TW);
        const int nvalid = min(TNF, ro.MS - t * TNF);
        const int c_end = min(cw * CPW + CPW, nvalid);
        for (int kk = cw * CPW; kk < c_end; kk += 2) {
            float2 bb[2 * DP];
            load_pair_rows<DP, F>(bb, tile, kk, fs);
            const bool has1 = (kk + 1 < c_end);
            float sm0 = 0.f, sm1 = 0.f, sl0 = 0.f, sl1 = 0.f, es0 = 1.f, es1 = has1 ? 1.f : 0.f;
            if (has_ss) {
                const float4 s0 = ss[kk], s1 = ss[kk + 1];   // (m2, ls, ES, -)
                sm0 = s0.x; sl0 = s0.y; es0 = s0.z;
                sm1 = s1.x; sl1 = s1.y; es1 = has1 ? s1.z : 0.f;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < DP; ++c) {
                    a0 = Lp<P>::accum(__fadd2_rn(bb[c], na[r][c]), a0, q.pg);
                    a1 = Lp<P>::accum(__fadd2_rn(bb[DP + c], na[r][c]), a1, q.pg);
                }
                float w0, w1;
                const float D0 = slice_sum<F>(a0.x + a0.y), D1 = slice_sum<F>(a1.x + a1.y);
                if (merged) {   // both roles at once (warp-uniform branch); m2 is subtracted inside the fma
                    w0 = eo[r] * ex2_approx(fmaf(D0, -q.coef, -lo_m[r]) - lo_s[r]) + es0 * ex2_approx(fmaf(D0, -q.coef, -sm0) - sl0);
                    w1 = eo[r] * ex2_approx(fmaf(D1, -q.coef, -lo_m[r]) - lo_s[r]) + es1 * ex2_approx(fmaf(D1, -q.coef, -sm1) - sl1);
                    if (!has1) w1 = 0.f;
                } else {        // exactly one of (owner, streamed) statistics is non-zero
                    w0 = es0 * ex2_approx(fmaf(D0, -q.coef, -(lo_m[r] + sm0)) - (lo_s[r] + sl0));
                    w1 = es1 * ex2_approx(fmaf(D1, -q.coef, -(lo_m[r] + sm1)) - (lo_s[r] + sl1));
                }
#pragma unroll
                for (int c = 0; c < DP; ++c) {
                    gacc[r][c] = Lp<P>::grad(w0, __fadd2_rn(bb[c], na[r][c]), gacc[r][c], q.pg);
                    gacc[r][c] = Lp<P>::grad(w1, __fadd2_rn(bb[DP + c], na[r][c]), gacc[r][c], q.pg);
                }
            }
        }
        __syncthreads();
    }

    // sum the CW column-warps of each row group (fixed order), write this split's partial
    float2* buf = reinterpret_cast<float2*>(smem);   // [(CW-1) * RW][R][DP][32]
    if (cw > 0) {
        const int slot = (cw - 1) * RW + rw;
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int c = 0; c < DP; ++c) buf[((slot * R + r) * DP + c) * 32 + lane] = gacc[r][c];
    }
    __syncthreads();
    if (cw == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int row = row_base + r * RPW + lr;
#pragma unroll
            for (int c = 0; c < DP; ++c) {
                float2 g = gacc[r][c];
#pragma unroll
                for (int w2 = 1; w2 < CW; ++w2) {
                    const float2 o = buf[((((w2 - 1) * RW + rw) * R + r) * DP + c) * 32 + lane];
                    g.x += o.x; g.y += o.y;
                }
                if (row < ro.BO)
                    *reinterpret_cast<float2*>(ro.part + ((size_t)blockIdx.y * ro.part_rows + row) * PTW + fs * 2 * DP + 2 * c) = g;
            }
        }
    }
}

// ---- launch helpers (instantiated once per exponent, see lpnce_inst.cuh) ---------------------------
template <int P, int DP, int F>
int launch_fwd_pd(const FwdParams& q, dim3 grid, cudaStream_t st) {
    constexpr int R = fwd_rows_per_thread(DP), RF = fwd_rows_per_thread(DP, true);
    auto kern = q.fast ? lpnce_fwd_kernel<P, DP, RF, kCW, F, true> : lpnce_fwd_kernel<P, DP, R, kCW, F, false>;
    const size_t smem = q.fast ? fwd_smem_bytes(DP, RF, F) : fwd_smem_bytes(DP, R, F);
    if (smem > 48 * 1024)
        CLICA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(q);
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}
template <int P, int DP, int F>
int launch_bwd_pd(const BwdParams& q, dim3 grid, cudaStream_t st) {
    constexpr int R = bwd_rows_per_thread(DP);
    auto kern = lpnce_bwd_kernel<P, DP, R, kCW, F>;
    const size_t smem = bwd_smem_bytes(DP, R, F);
    if (smem > 48 * 1024)
        CLICA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, st>>>(q);
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

// resident CTAs per SM of one instantiation (sizes the grid: splits are chosen so that the CTAs fill whole waves)
template <int P, int DP, int F>
int occ_fwd_pd(int fast) {
    static int cached[2] = {0, 0};
    const int i = fast ? 1 : 0;
    if (cached[i] == 0) {
        constexpr int R = fwd_rows_per_thread(DP), RF = fwd_rows_per_thread(DP, true);
        int n = 0;
        cudaError_t e = fast
            ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lpnce_fwd_kernel<P, DP, RF, kCW, F, true>, kThreads, fwd_smem_bytes(DP, RF, F))
            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lpnce_fwd_kernel<P, DP, R, kCW, F, false>, kThreads, fwd_smem_bytes(DP, R, F));
        if (e != cudaSuccess) n = 1;
        cached[i] = n < 1 ? 1 : n;
    }
    return cached[i];
}
template <int P, int DP, int F>
int occ_bwd_pd() {
    static int cached = 0;
    if (cached == 0) {
        constexpr int R = bwd_rows_per_thread(DP);
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lpnce_bwd_kernel<P, DP, R, kCW, F>, kThreads, bwd_smem_bytes(DP, R, F)) != cudaSuccess) n = 1;
        cached = n < 1 ? 1 : n;
    }
    return cached;
}

// (DP, F) combinations that are instantiated: F = 1 for d <= 40; F in {2, 4, 8} x DP in {16, 20} for d <= 320
#define CLICA_DISPATCH_DPF(P_, DPV, FV, FN, ...)                                                        \
    switch ((FV) * 100 + (DPV)) {                                                                       \
        case 102: return FN<P_, 2, 1>(__VA_ARGS__);                                                     \
        case 103: return FN<P_, 3, 1>(__VA_ARGS__);                                                     \
        case 104: return FN<P_, 4, 1>(__VA_ARGS__);                                                     \
        case 105: return FN<P_, 5, 1>(__VA_ARGS__);                                                     \
        case 108: return FN<P_, 8, 1>(__VA_ARGS__);                                                     \
        case 112: return FN<P_, 12, 1>(__VA_ARGS__);                                                    \
        case 116: return FN<P_, 16, 1>(__VA_ARGS__);                                                    \
        case 120: return FN<P_, 20, 1>(__VA_ARGS__);                                                    \
        case 216: return FN<P_, 16, 2>(__VA_ARGS__);                                                    \
        case 220: return FN<P_, 20, 2>(__VA_ARGS__);                                                    \
        case 416: return FN<P_, 16, 4>(__VA_ARGS__);                                                    \
        case 420: return FN<P_, 20, 4>(__VA_ARGS__);                                                    \
        case 816: return FN<P_, 16, 8>(__VA_ARGS__);                                                    \
        case 820: return FN<P_, 20, 8>(__VA_ARGS__);                                                    \
        default: return clica::fail(CLICA_E_UNSUPPORTED, "no loss kernel for DP=%d F=%d", DPV, FV);      \
    }

#endif  // __CUDACC__

// one translation unit per exponent keeps the build parallel: lpnce_p{0,1,2,3,4}.cu define these
int launch_fwd_p0(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s);
int launch_fwd_p1(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s);
int launch_fwd_p2(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s);
int launch_fwd_p3(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s);
int launch_fwd_p4(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s);
int launch_bwd_p0(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int launch_bwd_p1(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int launch_bwd_p2(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int launch_bwd_p3(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int launch_bwd_p4(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int occ_fwd_p0(int DP, int F, int fast); int occ_fwd_p1(int DP, int F, int fast); int occ_fwd_p2(int DP, int F, int fast); int occ_fwd_p3(int DP, int F, int fast); int occ_fwd_p4(int DP, int F, int fast);
int occ_bwd_p0(int DP, int F); int occ_bwd_p1(int DP, int F); int occ_bwd_p2(int DP, int F); int occ_bwd_p3(int DP, int F); int occ_bwd_p4(int DP, int F);

}  // namespace clica
