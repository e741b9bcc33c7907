/*
 * clica.h -- C ABI of libclica_sm100.so: the B200 (sm_100a) implementation of cl-ica's InfoNCE
 * training-step hot path.  Plain C: device pointers, sizes, a cudaStream_t passed as void*.  No torch
 * types, no C++ types, nothing thrown across the boundary.
 *
 * The reference (brendel-group/cl-ica) is pure Python on PyTorch and has no FFI of its own; each entry
 * point below names the reference code (file:line under the reference root) whose arithmetic it replaces.
 * The reference-side binding (ctypes stub + autograd.Function) is shown in INTEGRATION.md and shipped
 * in cl-ica_b200/_lib.py / functional.py.
 *
 * Conventions
 *   - all matrices are fp32, row-major, unit column stride, explicit leading dimension (in elements);
 *   - every pointer is a DEVICE pointer into caller-owned memory (PyTorch caching allocator) unless
 *     stated otherwise; the library allocates no persistent device memory;
 *   - every call is asynchronous on `stream` (a cudaStream_t), never synchronises, never touches the
 *     default stream, and is CUDA-graph capturable;
 *   - return value: 0 on success; >0 a cudaError_t; <0 a CLICA_E_* code.  clica_last_error() returns a
 *     thread-local description of the last failure on the calling thread.  No silent fallbacks.
 */
#ifndef CLICA_H_
#define CLICA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLICA_ABI_VERSION 1

#if defined(__GNUC__)
#define CLICA_API __attribute__((visibility("default")))
#else
#define CLICA_API
#endif

#define CLICA_E_BADARG      (-1)  /* null pointer, negative size, ld < cols, ...              */
#define CLICA_E_UNSUPPORTED (-2)  /* combination not implemented by the CUDA path (e.g. p < 1) */
#define CLICA_E_WORKSPACE   (-3)  /* workspace too small                                       */
#define CLICA_E_ALIGN       (-4)  /* pointer/ld alignment requirement violated                 */
#define CLICA_E_ARCH        (-5)  /* device is not compute capability 10.x                     */

/* GEMM numeric modes of the encoder kernels */
#define CLICA_GEMM_3XTF32 0       /* tcgen05 kind::tf32, hi/lo split, 3 MMAs: ~fp32 accuracy (parity mode) */
#define CLICA_GEMM_TF32   1       /* tcgen05 kind::tf32 single pass (fast mode, ~1e-3 relative)            */
#define CLICA_GEMM_FP32   3       /* CUDA-core FFMA, exact fp32 products (skinny layers; debug)            */

CLICA_API int         clica_abi_version(void);
CLICA_API const char* clica_last_error(void);
/* Fills sm_count / cc_major / cc_minor of the current device; any pointer may be NULL. */
CLICA_API int         clica_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Lp-InfoNCE loss, forward.         replaces  losses.py:443-477 (LpSimCLRLoss.loss, p >= 1 branch)
 *                                             losses.py:506-510 (_logmeanexp)
 *
 *   D_ik  = sum_c |z1[i,c] - z3[k,c]|^p ,  pos_i = sum_c |z1[i,c] - z2[i,c]|^p      (i < B, k < M)
 *   include_pos = 1 (simclr_compatibility_mode):  lse_i = log( sum_k e^{-D_ik/tau} + e^{-pos_i/tau} )
 *   include_pos = 0:                              lse_i = log( sum_k e^{-D_ik/tau} ) - log M
 *   loss_i = 2 ( alpha pos_i / tau + (1 - alpha) lse_i )
 *
 * The B x M distance matrix is never materialised: anchors stay in registers, negatives stream through
 * shared memory, the row soft-max is evaluated online.
 *
 * outputs  loss_i[B], lse[B] (as defined above), pos[B] (un-scaled pos_i),
 *          rowstat[B][2] = per anchor (m2, ls): the reference maximum of the log2-domain logits and the
 *          log2 of the sum of exp2(logit - m2).  The backward re-derives every soft-max weight as
 *          exp2((-D*log2(e)/tau - m2) - ls) -- the forward's own arithmetic -- so a row's weights sum to 1 to
 *          fp32 accuracy even when |lse| is in the thousands (8-byte aligned),
 *          scalars[3] = { mean_i loss_i, mean_i pos_i/tau, mean_i lse_i }   (losses.py:469-477)
 * p        any real >= 1; p in {1,2,3,4} use multiply-only inner loops, other p use ex2/lg2.
 * use_pow  must be 1 (losses.py:452-454, `pow=True`, the only value any reference script uses).
 * ws       scratch of at least clica_lpnce_workspace_bytes(B, M, d) bytes, 16-byte aligned (split
 *          partials; dead after the call -- the backward is stateless and takes rowstat / pos back).
 *          ZERO CONTRACT (all three loss entry points): the first CLICA_LPNCE_COUNTER_BYTES bytes of a loss
 *          workspace hold arrival counters of the in-kernel split merge.  The caller zero-fills them ONCE
 *          (e.g. when allocating the workspace); every call leaves them zero again, so a workspace can be
 *          reused by consecutive calls on one stream without further memsets (and replayed from a CUDA graph).
 * One kernel launch: pair walk, merge of the column splits, positive pair, per-item loss and the means.
 * ---------------------------------------------------------------------------------------------- */
#define CLICA_LPNCE_COUNTER_BYTES 65536
CLICA_API size_t clica_lpnce_workspace_bytes(int B, int M, int d);

CLICA_API int clica_lpnce_fwd(const float* z1, int ld1, const float* z2, int ld2, const float* z3, int ld3,
                    int B, int M, int d, float p, float tau, float alpha, int include_pos,
                    int use_pow, float* loss_i, float* lse, float* pos, float* rowstat,
                    float* scalars3, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Lp-InfoNCE loss, backward.        replaces  autograd through losses.py:447-477
 *                                   (LogsumexpBackward, CatBackward, PowBackward,
 *                                    LinalgVectorNormBackward, SubBackward)
 *
 * Gradient of  L = g_mean * mean_i loss_i + sum_i g_loss_i[i] * loss_i  with respect to z1, z2, z3.
 *   g_mean    device scalar (nullable => 0):  dL/d scalars3[0]
 *   g_loss_i  device [B]    (nullable => 0):  dL/d loss_i
 *   rowstat, pos  the forward's outputs
 *   g_z1[B,d] (ld = ldg1), g_z2[B,d], g_z3[M,d]: any may be NULL (skipped).  g_z1 receives only the
 *             anchor-role term; a caller whose z3 aliases z1 (torch.roll, main_mlp.py:272) adds g_z3
 *             back through its own autograd graph exactly as the reference does.
 * Zero differences contribute exactly zero (torch's norm backward masks them).
 * ws: >= clica_lpnce_bwd_workspace_bytes(B, M, d) bytes, zero contract as for the forward.
 * g_loss_i == NULL (the training step): one kernel launch; otherwise three (coefficients, pair walk, reduce).
 * ---------------------------------------------------------------------------------------------- */
CLICA_API size_t clica_lpnce_bwd_workspace_bytes(int B, int M, int d);

CLICA_API int clica_lpnce_bwd(const float* z1, int ld1, const float* z2, int ld2, const float* z3, int ld3,
                    int B, int M, int d, float p, float tau, float alpha, int include_pos,
                    int use_pow, const float* rowstat, const float* pos, const float* g_mean,
                    const float* g_loss_i, float* g_z1, int ldg1, float* g_z2, int ldg2,
                    float* g_z3, int ldg3, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-sharded variant (one process per GPU).  The caller all-gathers the encoder
 * outputs (NCCL) so that z_all[M,d] holds every rank's anchors, runs clica_lpnce_fwd on its own rows
 * [row0, row0+B) against z_all, all-gathers rowstat into rowstat_all[M][2], then calls this: it returns, for the
 * LOCAL rows only, the complete gradient of the GLOBAL mean loss  (1/M) sum_i loss_i  with
 * z3 = roll(z_all, 1):   anchor-role term + column-role term (rows i of every rank that use local
 * row k as a negative, weights from rowstat_all[i]) + positive term.  No second collective.
 *   g_scale  device scalar: dL/d(global mean loss)   (nullable => 1)
 * ---------------------------------------------------------------------------------------------- */
CLICA_API size_t clica_lpnce_bwd_sharded_workspace_bytes(int B, int M, int d);

CLICA_API int clica_lpnce_bwd_sharded(const float* z1_local, int ld1, const float* z2_local, int ld2,
                            const float* z_all, int ld3, const float* rowstat_all,
                            const float* pos_local, int B, int M, int d, int row0, float p,
                            float tau, float alpha, int include_pos, const float* g_scale,
                            float* g_z1, int ldg1, float* g_z2, int ldg2,
                            void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Encoder layer kernels.            replace  nn.Linear + nn.LeakyReLU of encoders.py:38-48
 *                                   (ATen addmm / leaky_relu and their autograd: AddmmBackward,
 *                                    LeakyReluBackward)
 *
 *   fwd        y[M,N]  = act( x[M,K] W[N,K]^T + b[N] ),  act = LeakyReLU(slope) or identity (slope = 1)
 *   bwd_data   dx[M,K] = ( dy[M,N] W[N,K] ) * act'(x_act) ,  act'(v) = v > 0 ? 1 : slope_prev
 *              (x_act = this layer's INPUT = previous layer's activation output; nullable => no mask)
 *   bwd_weight dW[N,K] = dy[M,N]^T x[M,K] ,  db[N] = sum_m dy[m,:]         (db nullable)
 *
 * `mode` is one of CLICA_GEMM_*.  Tensor-core modes stage fp32 operands with TMA into 128B-swizzled
 * shared memory and accumulate in TMEM (tcgen05.mma kind::tf32); shapes the tensor-core path cannot
 * address (leading dimension not a multiple of 4 floats, K or N < 16) are routed to the exact-fp32
 * CUDA-core kernel inside the same call -- that is a shape rule documented in DESIGN.md, not a fallback.
 * ws: >= clica_linear_workspace_bytes(M, N, K, mode) bytes, 1024-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
CLICA_API size_t clica_linear_workspace_bytes(int M, int N, int K, int mode);

CLICA_API int clica_linear_act_fwd(const float* x, int ldx, const float* W, int ldw, const float* b,
                         float* y, int ldy, int M, int K, int N, float slope, int mode,
                         void* ws, size_t ws_bytes, void* stream);

CLICA_API int clica_linear_act_bwd_data(const float* dy, int lddy, const float* W, int ldw,
                              const float* x_act, int ldxa, float slope_prev,
                              float* dx, int lddx, int M, int K, int N, int mode,
                              void* ws, size_t ws_bytes, void* stream);

CLICA_API int clica_linear_bwd_weight(const float* dy, int lddy, const float* x, int ldx,
                            float* dW, int lddw, float* db, int M, int K, int N, int mode,
                            void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole-encoder calls (one C call per forward / backward of the Linear+LeakyReLU stack, so the host
 * pays one FFI crossing instead of 7 / 20).            replaces  encoders.py:36-58 forward + autograd
 *
 * Layer l (0 <= l < L): W[l] is [widths[l+1], widths[l]] (ld = widths[l]), b[l] is [widths[l+1]].
 * acts[l] (l = 0..L) are caller-allocated: acts[0] = input x and acts[L] = output are dense
 * [M, widths[l]] fp32; the hidden ones (0 < l < L) are OPAQUE buffers of clica_mlp_act_floats(M,
 * widths[l], mode) floats, 16-byte aligned, holding the activation in the GEMM operand format of `mode`
 * ((hi, lo) planes in 3xTF32 mode) -- they are what the backward needs (saved by the caller's autograd
 * ctx) and are never converted.  LeakyReLU(slope) after every layer but the last.
 * backward: g_out = dL/d acts[L]; dW[l], db[l] are written (not accumulated); g_in nullable
 * (main_mlp.py feeds the frozen mixing net's output, which needs no gradient).
 * The first and last layer (n -> 10n, 10n -> n: rows TMA cannot address, <2% of the flops) always run on
 * the exact-fp32 CUDA-core kernel; hidden layers with M, K, N >= 32 run on tcgen05 in tensor-core modes.
 * ---------------------------------------------------------------------------------------------- */
CLICA_API size_t clica_mlp_act_floats(int M, int width, int mode);
CLICA_API size_t clica_mlp_workspace_bytes(int M, int L, const int* widths, int mode);

/* Weights change once per optimizer step but the encoder runs four times per step (2 forward, 2 backward):
 * clica_mlp_pack_weights converts them once into the GEMM operand format (`packed`: 1024-byte aligned,
 * clica_mlp_packed_weight_bytes bytes) and fwd / bwd take it as `packed_weights` (NULL: packed internally).
 * grads_prezeroed != 0: the caller has already zeroed every dW[l] / db[l] (one memset of a flat buffer). */
CLICA_API size_t clica_mlp_packed_weight_bytes(int L, const int* widths, int mode);
CLICA_API int clica_mlp_pack_weights(int L, const int* widths, const float* const* W, int mode, void* packed,
                  size_t packed_bytes, void* stream);
/* Where layer l lives inside `packed`: byte offsets of its hi / lo planes (-1: layer not packed / no lo plane) and
 * the planes' row pitch in floats (columns [widths[l], ld) are zero padding).  For an optimizer that keeps the
 * planes current itself (clica_adam_step_capturable_packed) instead of re-packing after every step. */
CLICA_API int clica_mlp_packed_weight_layout(int L, const int* widths, int mode, long long* hi_off,
                  long long* lo_off, int* ld);

CLICA_API int clica_mlp_fwd(int L, const int* widths, const float* const* W, const float* const* b,
                  float* const* acts, int M, float slope, int mode, const void* packed_weights,
                  void* ws, size_t ws_bytes, void* stream);

CLICA_API int clica_mlp_bwd(int L, const int* widths, const float* const* W, const float* const* acts,
                  const float* g_out, float* const* dW, float* const* db, float* g_in,
                  int M, float slope, int mode, const void* packed_weights, int grads_prezeroed,
                  void* ws, size_t ws_bytes, void* stream);

/* Layers l_first, l_first-1, ..., l_last (L-1 >= l_first >= l_last >= 0) of the same backward chain; the
 * gradient flowing between two consecutive range calls stays in `ws` (same workspace, same stream).
 * clica_mlp_bwd == range(L-1 .. 0).  No reference analogue: the row-sharded multi-GPU step
 * (semantics of main_3dident.py:373,480-492) issues the backward bucket by bucket so that the
 * NCCL all-reduce of a finished layer's dW/db overlaps the differentiation of the earlier layers. */
CLICA_API int clica_mlp_bwd_range(int L, const int* widths, const float* const* W, const float* const* acts,
                  const float* g_out, float* const* dW, float* const* db, float* g_in,
                  int M, float slope, int mode, const void* packed_weights, int grads_prezeroed,
                  int l_first, int l_last, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-tensor Adam step.     replaces  torch.optim.Adam.step (main_mlp.py:283,312) for the
 *                                   encoder's parameter list; same update rule as torch's default
 *                                   (no amsgrad, no weight decay, eps added after the bias-corrected
 *                                   sqrt).  Host arrays of `n` device pointers / element counts.
 *   step is the 1-based step count AFTER this update.
 * ---------------------------------------------------------------------------------------------- */
CLICA_API int clica_adam_step(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                    float* const* exp_avg_sq, const int64_t* numel, float lr, float beta1,
                    float beta2, float eps, int64_t step, float grad_scale, void* stream);

/* Capturable variant (CUDA-graph replays of the whole training step): the step count lives on the device.
 * step_state: 16-byte, 8-byte aligned device buffer { int64 step; float lr/(1-b1^t); float 1/sqrt(1-b2^t) },
 * zero-initialised by the caller before the first step; each call advances it by one on the device (a
 * one-thread kernel, same double-precision bias corrections as the host variant) and then applies the
 * update, so replaying a captured graph performs consecutive Adam steps with no host-side state. */
CLICA_API int clica_adam_step_capturable(int n, float* const* params, const float* const* grads,
                    float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel, float lr,
                    float beta1, float beta2, float eps, void* step_state, float grad_scale, void* stream);

/* Same update; tensor k with pack_hi[k] != NULL -- a row-major weight matrix with pack_cols[k] columns -- is ALSO
 * written in the tensor-core operand format at pack_hi[k] / pack_lo[k] (row pitch pack_ld[k] floats; pack_lo[k]
 * NULL: one plane with the fp32 value).  Replaces the clica_mlp_pack_weights pass that would otherwise follow
 * every optimizer step (main_mlp.py:283 followed by the next step's encoder forward). */
CLICA_API int clica_adam_step_capturable_packed(int n, float* const* params, const float* const* grads,
                    float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel, float lr,
                    float beta1, float beta2, float eps, void* step_state, float grad_scale,
                    float* const* pack_hi, float* const* pack_lo, const int* pack_cols, const int* pack_ld,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Frozen mixing network g of h = f o g (main_mlp.py:313), forward only.
 *                                   replaces  the nn.Sequential of invertible_network_utils.py:87-123
 *                                   (construct_invertible_mlp: L bias-free n x n Linear layers with
 *                                   LeakyReLU(slope) in between) for CUDA fp32 inputs: one kernel instead
 *                                   of L cuBLAS launches + L-1 elementwise launches.
 *   x [M, n] (ld ldx), W[l] = the l-th layer's [n, n] weight (out x in, contiguous), y [M, n] (ld ldy).
 *   Supported: 1 <= L <= 8, n <= 48, L*n*n*4 <= 48 KB (validated on B200 in round 2: tests/test_gpu_kernels_r2.py;
 *   GraphedTrainStep uses it for the mixing net unless CLICA_FUSED_MIXING=0).
 * ---------------------------------------------------------------------------------------------- */
CLICA_API int clica_mixing_fwd(const float* x, int ldx, const float* const* W, int L, int n, int M, float slope,
                     float* y, int ldy, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Launch accounting (bench.py's `gpu_launches` and per-kernel roofline numbers; no reference analogue).
 *   clica_launch_count(family)  kernels launched by this library in this process (family < 0: all)
 *   clica_prof_enable(on)       when on, every launch scope is bracketed by CUDA events on its stream
 *   clica_prof_collect(ms, n)   host arrays of CLICA_NUM_FAMILIES entries: summed device time (ms) and
 *                               number of scopes per kernel family since the last collect; synchronises.
 * Families: 0 loss fwd, 1 loss bwd, 2 loss finalize/prep/reduce, 3 tcgen05 GEMM, 4 CUDA-core GEMM,
 *           5 Adam, 6 misc (column sums, operand packing).
 * ---------------------------------------------------------------------------------------------- */
/* ------------------------------------------------------------------------------------------------
 * Device-side latent samplers.   replaces, for CUDA devices, the host-side samplers of
 *   spaces.py:47-119 (NRealSpace), :134-231 (NSphereSpace), :273-351 (NBoxSpace) and
 *   spaces_utils.py:82-142 (generalized normal via CPU Gamma draws; rejection loops with a host sync per round)
 * out[rows, n] (ld) <- rows samples;  space: 0 R^n, 1 unit sphere (draw in R^n, project), 2 box [box_lo, box_hi]
 * (every element is redrawn until it lies in the box);  dist: 0 uniform (sphere, box), 1 normal(mean, scale),
 * 2 laplace(mean, scale), 3 generalized normal(mean, scale, p) = mean + scale * sign * Gamma(1/p, 1)^(1/p).
 * mean: NULL / one row (mean_rows = 1) / one row per sample (mean_rows = rows).  Same distributions as the
 * reference, different random stream: Philox4x32-10 keyed by `seed`, every (element, attempt, offset) owns its
 * counter, so a (seed, offset) pair names a reproducible draw; callers advance `offset` per call.  One launch,
 * no host synchronisation.
 * ---------------------------------------------------------------------------------------------- */
CLICA_API int clica_sample_latents(float* out, int ld, int rows, int n, int space, int dist, const float* mean,
                     int ld_mean, int mean_rows, float scale, float p, float box_lo, float box_hi,
                     uint64_t seed, uint64_t offset, void* stream);

/* The persistent tcgen05 GEMMs normally launch one CTA (or CTA pair) per SM.  clica_tc_set_sm_reserve(n) makes
 * every later GEMM launch of this process leave n SMs free (0 <= n <= 64; default 0) -- the multi-GPU step sets it
 * while NCCL's all-reduce of a finished gradient bucket runs concurrently with the remaining backward GEMMs.
 * No reference analogue. */
CLICA_API int clica_tc_set_sm_reserve(int sms);

#define CLICA_NUM_FAMILIES 7
CLICA_API long long clica_launch_count(int family);
CLICA_API int clica_prof_enable(int on);
CLICA_API int clica_prof_collect(float* ms_by_family, int* scopes_by_family);

#ifdef __cplusplus
}
#endif
#endif /* CLICA_H_ */
