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
//   Column splits (grid.y) fill all SMs; the LAST split CTA of a row tile to finish (arrival counter in the
//   workspace) merges the split partials of its rows in a fixed order -- per-item loss, row statistics and the
//   three means in the forward; the summed gradient plus the positive-pair term in the backward -- so the forward
//   is ONE launch and the backward is ONE launch, and results are deterministic.
//   p = 2 additionally has a "dot form" of the distance (see below): half the instructions, guarded by a norm bound.
#pragma once
#include "common.cuh"

#include <math.h>

namespace clica {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTN = 128;          // streamed rows per shared-memory tile (divided by the feature-split factor F)
constexpr int kCW = 4;            // warps of a CTA that split the streamed rows of a tile
constexpr float kRescale = 24.f;  // lazy soft-max re-scaling threshold (log2 units)
constexpr float kDotFloor = -40.f;   // dot form: the running reference point must stay above this (no re-scaling there)
// Arrival counters (last-CTA-done reductions) live in the first kCounterBytes of every loss workspace.  The caller
// provides them ZERO (once, when the workspace is allocated); every kernel restores the zeros before it ends.
constexpr size_t kCounterBytes = CLICA_LPNCE_COUNTER_BYTES;
constexpr int kMaxRowTilesFwd = (int)(kCounterBytes / 4) - 4;
constexpr int kMaxRowTilesBwd = (int)(kCounterBytes / 8);      // per role

// P == 2 with one lane per pair: the distance may be evaluated in "dot form" (see half_norm_of_row)
constexpr bool dot_capable(int P, int F) { return P == 2 && F == 1; }

// rows owned per thread: 2 while the register budget allows it; the dot form (p = 2, d <= 10) has so little work per
// pair that 4 owner rows per streamed-row load pay (R4 variant, selected at run time)
constexpr int fwd_rows_per_thread(int DP, bool r4 = false) { return (r4 && DP <= 5) ? 4 : (DP <= 12 ? 2 : 1); }
constexpr int bwd_rows_per_thread(int DP) { return DP <= 5 ? 2 : 1; }
constexpr int rows_per_cta(int R, int F = 1) { return (32 / F) * R * (kWarps / kCW); }
// shared-memory layout of a streamed row: F feature slices of 2*DP floats (+4 floats of padding when F > 1 so
// that the F lanes of a pair hit different bank groups)
constexpr int slice_floats(int DP, int F) { return 2 * DP + (F > 1 ? 4 : 0); }
constexpr int row_floats(int DP, int F) { return F * slice_floats(DP, F); }
constexpr int tile_rows(int F) { return kTN / F; }
// per-stage floats
constexpr int fwd_stage_floats(int DP, int F) { return tile_rows(F) * row_floats(DP, F); }
constexpr int bwd_stage_floats(int DP, int F) { return tile_rows(F) * row_floats(DP, F) + 4 * tile_rows(F); }   // + (m2, ls, E, -) per row

inline size_t fwd_smem_bytes(int DP, int R, int F = 1) {
    size_t tiles = 2ull * fwd_stage_floats(DP, F) * sizeof(float);
    size_t merge = 2ull * kWarps * R * 32 * sizeof(float);
    size_t fin = 3ull * kThreads * sizeof(double);        // fused finalize: per-thread partial means
    size_t m = tiles > merge ? tiles : merge;
    return m > fin ? m : fin;
}
inline size_t bwd_smem_bytes(int DP, int R, int F = 1) {
    size_t tiles = 2ull * bwd_stage_floats(DP, F) * sizeof(float);
    size_t merge = (size_t)(kCW - 1) * (kWarps / kCW) * R * DP * 32 * sizeof(float2);
    size_t fin = 2ull * kThreads * sizeof(float);        // final phase: per-row coefficients
    size_t m = tiles > merge ? tiles : merge;
    return m > fin ? m : fin;
}

struct FwdParams {
    const float* O; int ldO; int BO;     // owner rows (anchors z1)
    const float* S; int ldS; int MS;     // streamed rows (negatives z3)
    const float* Z2; int ld2;            // positives (fused finalize)
    int d; float coef; float pg;         // coef = log2(e)/tau, pg = p
    float tau; float alpha; int include_pos;
    int tiles_per_split; int nsplit; int flat16;
    float dot_limit;                     // dot form allowed while coef*(|a-c|^2 + max_j |b_j-c|^2) <= dot_limit (<= 0: off)
    float* part_m; float* part_s; int part_stride;   // [nsplit][part_stride]
    int* done_counter; int* tile_counter;            // zero on entry, zero on exit
    double* block_sums;                              // [row_tiles][3]
    float* loss_i; float* lse; float* pos; float2* rowstat; float* scalars;
};

// Row statistics of the forward, per anchor: (m2, ls) = (reference maximum of the log2-domain logits,
// log2 of the sum of exp2(logit - m2)); the soft-max weight of a pair is exp2((-D*coef - m2) - ls), i.e. the
// SAME arithmetic the forward used, so the weights of a row sum to 1 to fp32 accuracy however large |lse| is.
// w(owner i, streamed j) = ES[j] * exp2( (-D*coef - m2) - ls ), (m2, ls) taken from the owner (anchor role)
// or from the streamed row (column role);  gacc_i += w * G'(s_j - o_i)/p
// Upstream gradient: gl_i = g_mean * inv_count + g_loss_i[i]; E_i = 2 gl_i (1-alpha)/tau.  With g_loss_i == NULL
// (the training step) E is one scalar the kernel derives from *g_mean itself and the final phase (sum of the split
// partials, positive-pair term) is fused into the kernel (FUSED mode); per-item upstream gradients take the
// prep -> kernel -> reduce route with E / CP arrays.
struct BwdRole {
    const float* O; int ldO; int BO;
    const float* S; int ldS; int MS;
    const float2* LO;                   // owner row statistics (nullable -> (0, 0))
    const float2* LS;                   // streamed row statistics (nullable -> (0, 0))
    const float* ES;                    // non-fused: streamed coefficient array (with LS); fused: NULL
    const float* EO;                    // non-fused MERGED mode: owner coefficient array
    int merged;                         // owner set == streamed set (e.g. z3 = roll(z1), and always when sharded): the
                                        // distance is symmetric, so one pass accumulates both roles:
                                        // w = E_i exp2(.. - stat_i) + E_j exp2(.. - stat_j)
    int stream_weighted;                // fused: multiply by E (column role / merged); anchor role: 0
    int tiles_per_split; int nsplit; int flat16;
    float* part; int part_rows;         // [nsplit][part_rows][F*2*DP]
    int row_tiles;
    // fused final phase (last CTA of a row tile): g_out = p * (scale * sum_s part) [+ CP * G'(o - z2), g_z2 = -that]
    int* tile_counter;                  // zero on entry, zero on exit
    float* g_out; int ldg;              // nullable
    int scale_by_E;                     // anchor role: partials still need E_row
    const float* Z2; int ld2; const float2* LP; const float* POS; float* g_z2; int ldg2; int with_pos;
};
struct BwdParams {
    BwdRole role[2];
    int nroles;
    int d; float coef; float pg;
    float tau; float alpha; int include_pos;
    int fused;                          // see above
    const float* g_mean; float default_g; float inv_count;
    float dot_limit;
};

#ifdef __CUDACC__

// ---- per-exponent arithmetic ----------------------------------------------------------------------
// P = 1,2,3,4: multiply-only; P = 0: generic real exponent through lg2/ex2 (pg = p); P = kSim: dot-product similarity.
constexpr int kSim = 5;
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
    // pair forms: b = two features of the streamed row, na = the NEGATED features of the owner row.
    // P = kSim (dot-product similarity, the reference's SimCLRLoss, losses.py:186): the "distance" is D = -a.b, so that
    // the logit -D/tau = a.b/tau and everything downstream (soft-max, weights, -dD/d owner = b) is shared with the Lp family
    static __device__ __forceinline__ float2 pair_acc(float2 b, float2 na, float2 acc, float pg) {
        if constexpr (P == kSim) return __ffma2_rn(b, na, acc);
        else return accum(__fadd2_rn(b, na), acc, pg);
    }
    static __device__ __forceinline__ float2 pair_grad(float w, float2 b, float2 na, float2 g, float pg) {
        if constexpr (P == kSim) return __ffma2_rn(make_float2(w, w), b, g);
        else return grad(w, __fadd2_rn(b, na), g, pg);
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

// scalar |t|^p and d|t|^p/dt for the O(B*d) positive-pair work (accurate libm paths for generic p)
__device__ __forceinline__ float abs_pow(float t, float p) {
    float a = fabsf(t);
    if (p == 1.f) return a;
    if (p == 2.f) return a * a;
    if (p == 3.f) return a * a * a;
    if (p == 4.f) { float u = a * a; return u * u; }
    return a == 0.f ? 0.f : exp2f(p * log2f(a));
}
__device__ __forceinline__ float dabs_pow(float t, float p) {   // 0 at t == 0 (torch masks it too)
    if (t == 0.f) return 0.f;
    float a = fabsf(t);
    float m;
    if (p == 1.f) m = 1.f;
    else if (p == 2.f) m = 2.f * a;
    else if (p == 3.f) m = 3.f * a * a;
    else if (p == 4.f) m = 4.f * a * a * a;
    else m = p * exp2f((p - 1.f) * log2f(a));
    return copysignf(m, t);
}

// ---- dot form (p = 2, one lane per pair) -------------------------------------------------------------
// |a - b|^2 / 2 = |a|^2/2 + |b|^2/2 - a.b : five packed FMAs per pair at d = 10 instead of five packed subtractions +
// five packed FMAs, and the half norms ride along (owner: a register; streamed: one float per row, computed by each
// warp for the 32 rows of the tile it is about to walk).  The cancellation error of the form is ~1e-7 * (|a|^2 + |b|^2)
// in D, so it is used -- warp by warp, tile by tile -- only while  coef * (|a|^2 + max_j |b_j|^2) <= dot_limit (8):
// measured by fp32 emulation the logit's absolute error is then <= ~2.5e-7 * that bound (rms 6x smaller; unit-sphere
// outputs at tau = 1 sit at 2.9), and every logit lies in [-2 * dot_limit, 0] so no re-scaling can be needed.  Every
// other (warp, tile) takes the subtract-then-square loop, which is exact for coincident rows and has no cancellation.
// Returns |row|^2 / 2 of this lane's streamed row of the warp's slice (0 for rows past the valid count).
template <int DP>
__device__ __forceinline__ float half_norm_of_row(const float* tile_row, bool valid) {
    const float2* row = reinterpret_cast<const float2*>(tile_row);
    float h0 = 0.f, h1 = 0.f;
#pragma unroll
    for (int c = 0; c < DP; ++c) {
        const float2 v = row[c];
        h0 = fmaf(v.x, v.x, h0); h1 = fmaf(v.y, v.y, h1);
    }
    return valid ? 0.5f * (h0 + h1) : 0.f;
}

// ================================ forward ==========================================================
// One launch: pair walk -> split partials -> (last CTA of each row tile) merge of the splits, positive pair, per-item
// loss, row statistics -> (last row tile) the three means.  All merges run in a fixed order: deterministic.
template <int P, int DP, int R, int CW, int F>
__global__ void __launch_bounds__(kThreads) lpnce_fwd_kernel(const FwdParams q) {
    pdl_enter();
    constexpr bool DOT = dot_capable(P, F);
    constexpr int RW = kWarps / CW;
    constexpr int RPW = 32 / F;                // rows per warp per r
    constexpr int ROWS = RPW * R * RW;
    constexpr int TW = row_floats(DP, F);      // floats per streamed row in shared memory
    constexpr int TNF = tile_rows(F);
    constexpr int CPW = TNF / CW;              // streamed rows per warp per tile
    constexpr int STAGE = fwd_stage_floats(DP, F);
    static_assert(ROWS <= kThreads, "the fused finalize gives one thread per owner row");
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_flag;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rw = warp / CW, cw = warp % CW;
    const int fs = lane % F, lr = lane / F;
    const int row_base = blockIdx.x * ROWS + rw * (RPW * R);

    float2 na[R][DP];
    load_owner_rows<DP, R, F>(na, q.O, q.ldO, q.BO, q.d, row_base, lane);
    float ha[R];
    const bool dot_on = DOT && q.dot_limit > 0.f;
    __shared__ __align__(16) float2 hbw[DOT ? kWarps : 1][32];   // per warp: (|b|^2/2, 0) of its 32 streamed rows of the tile
    if constexpr (DOT) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float h0 = 0.f, h1 = 0.f;
#pragma unroll
            for (int c = 0; c < DP; ++c) { h0 = fmaf(na[r][c].x, na[r][c].x, h0); h1 = fmaf(na[r][c].y, na[r][c].y, h1); }
            ha[r] = 0.5f * (h0 + h1);
        }
    }

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
            load_tile<DP, F>(smem + (stage ^ 1) * STAGE, q.S, q.ldS, q.MS, q.d, (t + 1) * TNF, q.flat16, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* tile = smem + stage * STAGE;
        const int nvalid = min(TNF, q.MS - t * TNF);
        if (t == t0) {
            // reference point of the lazy soft-max: the logit of the first streamed row of this split
            const float2* b = reinterpret_cast<const float2*>(tile + fs * slice_floats(DP, F));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < DP; ++c) acc = Lp<P>::pair_acc(b[c], na[r][c], acc, q.pg);
                m[r] = -slice_sum<F>(acc.x + acc.y) * q.coef;
            }
        }
        const int c_end = min(cw * CPW + CPW, nvalid);
        bool use_dot = false;
        if constexpr (DOT) {
            if (dot_on) {
                // warp-uniform decision for this warp's CPW (= 32) streamed rows of the tile
                static_assert(!DOT || CPW == 32, "one streamed row per lane");
                const int jrow = cw * CPW + lane;
                const float h = half_norm_of_row<DP>(tile + jrow * TW, jrow < nvalid);
                __syncwarp();                                   // the previous tile's readers of hbw are done
                hbw[warp][lane] = make_float2(h, 0.f);
                float hb = h;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) hb = fmaxf(hb, __shfl_xor_sync(0xffffffffu, hb, o));
                bool ok = true;
#pragma unroll
                for (int r = 0; r < R; ++r) ok = ok && (2.f * q.coef * (ha[r] + hb) <= q.dot_limit) && (m[r] >= kDotFloor);
                __syncwarp();
                use_dot = __all_sync(0xffffffffu, ok);
            }
        }
        if (use_dot) {
            if constexpr (DOT) {
                const float K = -2.f * q.coef;
                float ci[R];
#pragma unroll
                for (int r = 0; r < R; ++r) ci[r] = fmaf(K, ha[r], -m[r]);
                const float4* hbv = reinterpret_cast<const float4*>(&hbw[warp][0]);
                for (int kk = cw * CPW; kk < c_end; kk += 2) {
                    float2 bb[2 * DP];
                    load_pair_rows<DP, F>(bb, tile, kk, fs);
                    const float4 h2 = hbv[(kk - cw * CPW) >> 1];  // (|b_kk|^2/2, 0, |b_kk+1|^2/2, 0)
                    const bool has1 = (kk + 1 < c_end);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float2 a0 = make_float2(h2.x, h2.y), a1 = make_float2(h2.z, h2.w);
#pragma unroll
                        for (int c = 0; c < DP; ++c) {
                            a0 = __ffma2_rn(bb[c], na[r][c], a0);
                            a1 = __ffma2_rn(bb[DP + c], na[r][c], a1);
                        }
                        const float x0 = fmaf(a0.x + a0.y, K, ci[r]);
                        const float x1 = has1 ? fmaf(a1.x + a1.y, K, ci[r]) : -INFINITY;   // ex2(-inf) = +0
                        s[r] += ex2_approx(x0) + ex2_approx(x1);
                    }
                }
            }
        } else {
            for (int kk = cw * CPW; kk < c_end; kk += 2) {
                float2 bb[2 * DP];
                load_pair_rows<DP, F>(bb, tile, kk, fs);
                const bool has1 = (kk + 1 < c_end);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int c = 0; c < DP; ++c) {
                        a0 = Lp<P>::pair_acc(bb[c], na[r][c], a0, q.pg);
                        a1 = Lp<P>::pair_acc(bb[DP + c], na[r][c], a1, q.pg);
                    }
                    const float D0 = slice_sum<F>(a0.x + a0.y);
                    const float D1s = slice_sum<F>(a1.x + a1.y);
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

    // ---- fused finalize: the last split CTA of this row tile merges all splits of its rows ----------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) s_flag = (atomicAdd(q.tile_counter + blockIdx.x, 1) == q.nsplit - 1);
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    float v_loss = 0.f, v_pos = 0.f, v_lse = 0.f;
    const int i = blockIdx.x * ROWS + tid;
    if (tid < ROWS && i < q.BO) {
        // positive pair; loads are issued in independent batches of 8 (a load -> use loop pays one latency per trip)
        float ps = 0.f;
        const float* a = q.O + (size_t)i * q.ldO;
        const float* b = q.Z2 + (size_t)i * q.ld2;
        for (int c0 = 0; c0 < q.d; c0 += 8) {
            float av[8], bv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = c0 + u < q.d;
                av[u] = ok ? __ldg(a + c0 + u) : 0.f;
                bv[u] = ok ? __ldg(b + c0 + u) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) ps += (P == kSim) ? -av[u] * bv[u] : abs_pow(av[u] - bv[u], q.pg);   // 0 for the padding
        }
        // the positive's logit: its rounded value only competes for the row maximum; its soft-max term is evaluated with
        // the SAME fused expression the backward uses for w+ (fma(pos, -coef, -m2)), so that term / sum is reproduced
        // exactly there even when the positive IS the maximum and |logit| ~ 1e3 (a product rounded on one side and fused
        // on the other differs by half an ulp of the logit: 1e-4 on w+ ~ 1)
        const float xp = __fmul_rn(-ps, q.coef);
        float M = q.include_pos ? xp : -INFINITY;
        // pass 1: row maximum over the split partials; pass 2 (partials now in L2): rescaled sum, fixed order.
        // __ldcg: the partials were written by other CTAs -- never through this SM's L1
        for (int s0 = 0; s0 < q.nsplit; s0 += 8) {
            float pm[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                pm[u] = (s0 + u < q.nsplit) ? __ldcg(q.part_m + (size_t)(s0 + u) * q.part_stride + i) : -INFINITY;
#pragma unroll
            for (int u = 0; u < 8; ++u) M = fmaxf(M, pm[u]);
        }
        float S = q.include_pos ? exp2f(fmaf(ps, -q.coef, -M)) : 0.f;
        for (int s0 = 0; s0 < q.nsplit; s0 += 8) {
            float pm[8], psum[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = s0 + u < q.nsplit;
                pm[u] = ok ? __ldcg(q.part_m + (size_t)(s0 + u) * q.part_stride + i) : 0.f;
                psum[u] = ok ? __ldcg(q.part_s + (size_t)(s0 + u) * q.part_stride + i) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (s0 + u < q.nsplit) S += psum[u] * exp2f(pm[u] - M);
        }
        const float ls = log2f(S);
        q.rowstat[i] = make_float2(M, ls);
        float l = (M + ls) * 0.6931471805599453f;
        if (!q.include_pos) l -= logf((float)q.MS);
        const float li = 2.f * (q.alpha * ps / q.tau + (1.f - q.alpha) * l);
        q.loss_i[i] = li;
        q.lse[i] = l;
        q.pos[i] = ps;
        v_loss = li; v_pos = ps / q.tau; v_lse = l;
    }
    // deterministic means: per-row-tile sums, then the last row tile to finish adds them in tile order
    float* red = smem;                       // [3][8] (the pair walk is over: shared memory is free)
    __syncthreads();
    v_loss = warp_sum(v_loss); v_pos = warp_sum(v_pos); v_lse = warp_sum(v_lse);
    if (lane == 0) { red[warp] = v_loss; red[8 + warp] = v_pos; red[16 + warp] = v_lse; }
    __syncthreads();
    if (tid == 0) {
        double s0 = 0, s1 = 0, s2 = 0;
        for (int w = 0; w < 8; ++w) { s0 += red[w]; s1 += red[8 + w]; s2 += red[16 + w]; }
        q.block_sums[3 * blockIdx.x + 0] = s0;
        q.block_sums[3 * blockIdx.x + 1] = s1;
        q.block_sums[3 * blockIdx.x + 2] = s2;
        q.tile_counter[blockIdx.x] = 0;                          // restore the zero for the next launch
        __threadfence();
        s_flag = (atomicAdd(q.done_counter, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    double* fin = reinterpret_cast<double*>(smem);               // [3][256]: 6 KB <= every kernel's tile buffers
    {
        const volatile double* bs = q.block_sums;
        double s0 = 0, s1 = 0, s2 = 0;
        for (unsigned b = tid; b < gridDim.x; b += kThreads) { s0 += bs[3 * b]; s1 += bs[3 * b + 1]; s2 += bs[3 * b + 2]; }
        fin[tid] = s0; fin[kThreads + tid] = s1; fin[2 * kThreads + tid] = s2;
    }
    __syncthreads();
    if (tid < 3) {
        double t = 0;
        const int nb = gridDim.x < (unsigned)kThreads ? (int)gridDim.x : kThreads;
        for (int k = 0; k < nb; ++k) t += fin[tid * kThreads + k];
        q.scalars[tid] = (float)(t / q.BO);
    }
    if (tid == 0) *q.done_counter = 0;
}

// ================================ backward =========================================================
template <int P, int DP, int R, int CW, int F>
__global__ void __launch_bounds__(kThreads) lpnce_bwd_kernel(const BwdParams q) {
    pdl_enter();
    constexpr bool DOT = dot_capable(P, F);
    constexpr int RW = kWarps / CW;
    constexpr int RPW = 32 / F;
    constexpr int ROWS = RPW * R * RW;
    constexpr int TW = row_floats(DP, F);
    constexpr int TNF = tile_rows(F);
    constexpr int CPW = TNF / CW;
    constexpr int PTW = F * 2 * DP;                    // floats per row of the partial-gradient buffer
    constexpr int STAGE = bwd_stage_floats(DP, F);     // features + (m2, ls, E, |b|^2/2) per streamed row
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_flag;
    const BwdRole& ro = q.role[blockIdx.z];
    if ((int)blockIdx.x >= ro.row_tiles || (int)blockIdx.y >= ro.nsplit) return;
    const bool has_ss = (ro.LS != nullptr);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rw = warp / CW, cw = warp % CW;
    const int fs = lane % F, lr = lane / F;
    const int row_base = blockIdx.x * ROWS + rw * (RPW * R);
    const bool merged = ro.merged != 0;
    // fused mode: one upstream scalar for every row
    const float gl_u = q.fused ? (q.g_mean ? __ldg(q.g_mean) : q.default_g) * q.inv_count : 0.f;
    const float E_u = 2.f * gl_u * (1.f - q.alpha) / q.tau;

    float2 na[R][DP];
    float2 gacc[R][DP];
    float lo_m[R], lo_s[R], eo[R];
    load_owner_rows<DP, R, F>(na, ro.O, ro.ldO, ro.BO, q.d, row_base, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row_base + r * RPW + lr;
#pragma unroll
        for (int c = 0; c < DP; ++c) gacc[r][c] = make_float2(0.f, 0.f);
        const float2 st = (ro.LO != nullptr && row < ro.BO) ? __ldg(ro.LO + row) : make_float2(0.f, 0.f);
        lo_m[r] = st.x; lo_s[r] = st.y;
        eo[r] = 0.f;
        if (merged && row < ro.BO) eo[r] = q.fused ? E_u : __ldg(ro.EO + row);
    }
    float ha[R], wsum[R];
    const bool dot_on = DOT && q.dot_limit > 0.f;
    __shared__ __align__(16) float2 hbw[DOT ? kWarps : 1][32];   // per warp: (|b|^2/2, 0) of its 32 streamed rows of the tile
    if constexpr (DOT) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float h0 = 0.f, h1 = 0.f;
#pragma unroll
            for (int c = 0; c < DP; ++c) { h0 = fmaf(na[r][c].x, na[r][c].x, h0); h1 = fmaf(na[r][c].y, na[r][c].y, h1); }
            ha[r] = 0.5f * (h0 + h1);
            wsum[r] = 0.f;
        }
    }
    // streamed-side coefficient when it is uniform: E (column role / merged) or 1 (anchor role)
    const bool es_array = (ro.ES != nullptr);
    const float es_u = q.fused ? (ro.stream_weighted ? E_u : 1.f) : 1.f;

    const int ntiles = (ro.MS + TNF - 1) / TNF;
    const int t0 = blockIdx.y * ro.tiles_per_split;
    const int t1 = min(ntiles, t0 + ro.tiles_per_split);

    auto issue_tile = [&](int stage, int t) {
        float* dst = smem + stage * STAGE;
        load_tile<DP, F>(dst, ro.S, ro.ldS, ro.MS, q.d, t * TNF, ro.flat16, tid);
        if (has_ss && tid < TNF) {
            const int j = t * TNF + tid;
            const bool ok = j < ro.MS;
            float* sdst = dst + TNF * TW + 4 * tid;
            const float* stat = reinterpret_cast<const float*>(ro.LS);
            cp_async_4(sdst, ok ? (const void*)(stat + 2 * j) : (const void*)stat, ok ? 4 : 0);
            cp_async_4(sdst + 1, ok ? (const void*)(stat + 2 * j + 1) : (const void*)stat, ok ? 4 : 0);
            if (es_array) cp_async_4(sdst + 2, ok ? (const void*)(ro.ES + j) : (const void*)ro.ES, ok ? 4 : 0);   // 0 for tail rows
        }
        cp_async_commit();
    };
    issue_tile(0, t0);

    for (int t = t0; t < t1; ++t) {
        const int stage = (t - t0) & 1;
        if (t + 1 < t1) { issue_tile(stage ^ 1, t + 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float* tile = smem + stage * STAGE;
        const float4* ss = reinterpret_cast<const float4*>(tile + TNF * TW);
        const int nvalid = min(TNF, ro.MS - t * TNF);
        const int c_end = min(cw * CPW + CPW, nvalid);
        bool use_dot = false;
        if constexpr (DOT) {
            if (dot_on) {
                static_assert(!DOT || CPW == 32, "one streamed row per lane");
                const int jrow = cw * CPW + lane;
                const float h = half_norm_of_row<DP>(tile + jrow * TW, jrow < nvalid);
                __syncwarp();
                hbw[warp][lane] = make_float2(h, 0.f);
                float hb = h;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) hb = fmaxf(hb, __shfl_xor_sync(0xffffffffu, hb, o));
                bool ok = true;
#pragma unroll
                for (int r = 0; r < R; ++r) ok = ok && (2.f * q.coef * (ha[r] + hb) <= q.dot_limit);
                __syncwarp();
                use_dot = __all_sync(0xffffffffu, ok);
            }
        }
        for (int kk = cw * CPW; kk < c_end; kk += 2) {
            float2 bb[2 * DP];
            load_pair_rows<DP, F>(bb, tile, kk, fs);
            const bool has1 = (kk + 1 < c_end);
            float sm0 = 0.f, sm1 = 0.f, sl0 = 0.f, sl1 = 0.f, es0 = es_u, es1 = has1 ? es_u : 0.f;
            if (has_ss) {
                const float4 s0 = ss[kk], s1 = ss[kk + 1];   // (m2, ls, ES, -)
                sm0 = s0.x; sl0 = s0.y; sm1 = s1.x; sl1 = s1.y;
                if (es_array) { es0 = s0.z; es1 = has1 ? s1.z : 0.f; }
            }
            if (use_dot) {
                if constexpr (DOT) {
                    const float K = -2.f * q.coef;
                    const float4 h2 = reinterpret_cast<const float4*>(&hbw[warp][0])[(kk - cw * CPW) >> 1];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        float2 a0 = make_float2(h2.x, h2.y), a1 = make_float2(h2.z, h2.w);
#pragma unroll
                        for (int c = 0; c < DP; ++c) {
                            a0 = __ffma2_rn(bb[c], na[r][c], a0);
                            a1 = __ffma2_rn(bb[DP + c], na[r][c], a1);
                        }
                        const float kha = K * ha[r];
                        const float x0 = fmaf(a0.x + a0.y, K, kha);      // -D * coef
                        const float x1 = fmaf(a1.x + a1.y, K, kha);
                        float w0, w1;
                        if (merged) {
                            w0 = eo[r] * ex2_approx((x0 - lo_m[r]) - lo_s[r]) + es0 * ex2_approx((x0 - sm0) - sl0);
                            w1 = eo[r] * ex2_approx((x1 - lo_m[r]) - lo_s[r]) + es1 * ex2_approx((x1 - sm1) - sl1);
                            if (!has1) w1 = 0.f;
                        } else {
                            w0 = es0 * ex2_approx((x0 - (lo_m[r] + sm0)) - (lo_s[r] + sl0));
                            w1 = es1 * ex2_approx((x1 - (lo_m[r] + sm1)) - (lo_s[r] + sl1));
                        }
                        const float2 w0v = make_float2(w0, w0), w1v = make_float2(w1, w1);
#pragma unroll
                        for (int c = 0; c < DP; ++c) {
                            gacc[r][c] = __ffma2_rn(w0v, bb[c], gacc[r][c]);
                            gacc[r][c] = __ffma2_rn(w1v, bb[DP + c], gacc[r][c]);
                        }
                        wsum[r] += w0 + w1;
                    }
                }
                continue;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < DP; ++c) {
                    a0 = Lp<P>::pair_acc(bb[c], na[r][c], a0, q.pg);
                    a1 = Lp<P>::pair_acc(bb[DP + c], na[r][c], a1, q.pg);
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
                    gacc[r][c] = Lp<P>::pair_grad(w0, bb[c], na[r][c], gacc[r][c], q.pg);
                    gacc[r][c] = Lp<P>::pair_grad(w1, bb[DP + c], na[r][c], gacc[r][c], q.pg);
                }
            }
        }
        __syncthreads();
    }
    if constexpr (DOT) {
        // dot-form tiles accumulated sum_j w_j b_j: complete them to sum_j w_j (b_j - a) with na = -a
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float2 wv = make_float2(wsum[r], wsum[r]);
#pragma unroll
            for (int c = 0; c < DP; ++c) gacc[r][c] = __ffma2_rn(wv, na[r][c], gacc[r][c]);
        }
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
    if (!q.fused) return;

    // ---- fused final phase: the last split CTA of this row tile sums the splits (fixed order) and adds the
    // positive-pair term:  g_out = p * scale * sum_s part[s]  +  CP * G'(o - z2),   g_z2 = -CP * G'(o - z2)
    __threadfence();
    __syncthreads();
    if (tid == 0) s_flag = (atomicAdd(ro.tile_counter + blockIdx.x, 1) == ro.nsplit - 1);
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    const int row0 = blockIdx.x * ROWS;
    float* cp_s = smem;                              // [ROWS] positive-pair coefficient of each owner row
    __syncthreads();
    if (tid < ROWS) {
        float cpv = 0.f;
        const int row = row0 + tid;
        if (ro.with_pos && row < ro.BO) {
            const float2 st = __ldg(ro.LP + row);
            const float wpos = q.include_pos ? exp2f(fmaf(__ldg(ro.POS + row), -q.coef, -st.x) - st.y) : 0.f;
            cpv = 2.f * gl_u * (q.alpha - (1.f - q.alpha) * wpos) / q.tau;
        }
        cp_s[tid] = cpv;
    }
    __syncthreads();
    const float scale = (P == kSim ? 1.f : q.pg) * (ro.scale_by_E ? E_u : 1.f);
    const int nel = ROWS * q.d;
    for (int idx = tid; idx < nel; idx += kThreads) {
        const int r = idx / q.d, c = idx - r * q.d;
        const int row = row0 + r;
        if (row >= ro.BO) break;
        float acc = 0.f;
        for (int s0 = 0; s0 < ro.nsplit; s0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                v[u] = (s0 + u < ro.nsplit) ? __ldcg(ro.part + ((size_t)(s0 + u) * ro.part_rows + row) * PTW + c) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u];
        }
        float g = scale * acc;
        if (ro.with_pos) {
            const float ov = __ldg(ro.O + (size_t)row * ro.ldO + c), zv = __ldg(ro.Z2 + (size_t)row * ro.ld2 + c);
            if constexpr (P == kSim) {     // pos-"distance" -o.z2: d/do = -z2, d/dz2 = -o
                g -= cp_s[r] * zv;
                if (ro.g_z2) ro.g_z2[(size_t)row * ro.ldg2 + c] = -cp_s[r] * ov;
            } else {
                const float gp = cp_s[r] * dabs_pow(ov - zv, q.pg);
                g += gp;
                if (ro.g_z2) ro.g_z2[(size_t)row * ro.ldg2 + c] = -gp;
            }
        }
        if (ro.g_out) ro.g_out[(size_t)row * ro.ldg + c] = g;
    }
    if (tid == 0) ro.tile_counter[blockIdx.x] = 0;   // restore the zero for the next launch
}

// ---- launch helpers (instantiated once per exponent, see lpnce_inst.cuh) ---------------------------
// r4: the 4-rows-per-thread forward (only instantiated where fwd_rows_per_thread(DP, true) differs and the dot form exists)
template <int P, int DP, int F>
constexpr bool has_r4() { return dot_capable(P, F) && fwd_rows_per_thread(DP, true) != fwd_rows_per_thread(DP, false); }

template <int P, int DP, int F>
int launch_fwd_pd(const FwdParams& q, dim3 grid, cudaStream_t st, int r4) {
    constexpr int R = fwd_rows_per_thread(DP), R4 = fwd_rows_per_thread(DP, true);
    if constexpr (has_r4<P, DP, F>()) {
        if (r4) {
            auto kern = lpnce_fwd_kernel<P, DP, R4, kCW, F>;
            const size_t smem = fwd_smem_bytes(DP, R4, F);
            if (smem > 48 * 1024)
                CLICA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            launch_k(kern, grid, kThreads, smem, st, q);
            CLICA_CUDA_OK(cudaGetLastError());
            return 0;
        }
    }
    auto kern = lpnce_fwd_kernel<P, DP, R, kCW, F>;
    const size_t smem = fwd_smem_bytes(DP, R, F);
    if (smem > 48 * 1024)
        CLICA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_k(kern, grid, kThreads, smem, st, q);
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
    launch_k(kern, grid, kThreads, smem, st, q);
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

// resident CTAs per SM of one instantiation (sizes the grid: splits are chosen so that the CTAs fill whole waves)
template <int P, int DP, int F>
int occ_fwd_pd(int r4) {
    static int cached[2] = {0, 0};
    const int i = (r4 && has_r4<P, DP, F>()) ? 1 : 0;
    if (cached[i] == 0) {
        constexpr int R = fwd_rows_per_thread(DP), R4 = fwd_rows_per_thread(DP, true);
        int n = 0;
        cudaError_t e;
        if constexpr (has_r4<P, DP, F>()) {
            e = i ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lpnce_fwd_kernel<P, DP, R4, kCW, F>, kThreads, fwd_smem_bytes(DP, R4, F))
                  : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lpnce_fwd_kernel<P, DP, R, kCW, F>, kThreads, fwd_smem_bytes(DP, R, F));
        } else {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, lpnce_fwd_kernel<P, DP, R, kCW, F>, kThreads, fwd_smem_bytes(DP, R, F));
        }
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

// one translation unit per exponent keeps the build parallel: lpnce_p{0,1,2,3,4,5}.cu define these (5 = kSim)
int launch_fwd_p0(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s, int r4);
int launch_fwd_p1(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s, int r4);
int launch_fwd_p2(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s, int r4);
int launch_fwd_p3(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s, int r4);
int launch_fwd_p4(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s, int r4);
int launch_fwd_p5(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s, int r4);
int launch_bwd_p0(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int launch_bwd_p1(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int launch_bwd_p2(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int launch_bwd_p3(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int launch_bwd_p4(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int launch_bwd_p5(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s);
int occ_fwd_p0(int DP, int F, int r4); int occ_fwd_p1(int DP, int F, int r4); int occ_fwd_p2(int DP, int F, int r4); int occ_fwd_p3(int DP, int F, int r4); int occ_fwd_p4(int DP, int F, int r4); int occ_fwd_p5(int DP, int F, int r4);
int occ_bwd_p0(int DP, int F); int occ_bwd_p1(int DP, int F); int occ_bwd_p2(int DP, int F); int occ_bwd_p3(int DP, int F); int occ_bwd_p4(int DP, int F); int occ_bwd_p5(int DP, int F);

}  // namespace clica
