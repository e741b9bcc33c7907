// linear_api.cu -- C-ABI entry points of the encoder layers (see include/clica.h): per-layer
// Linear(+LeakyReLU) forward / backward-data / backward-weight and the whole-stack mlp_fwd / mlp_bwd.
// Replaces nn.Linear + nn.LeakyReLU of /root/reference/encoders.py:38-48 and their autograd nodes.
//
// Shape rule (not a fallback): a layer goes to the tcgen05 tensor-core GEMM (gemm_tc.cu) when the
// requested mode is a tensor-core mode AND TMA can address all three matrices (leading dimensions
// multiples of 4 floats, 16-byte aligned bases) AND M, K, N are all >= 32; otherwise -- the n -> 10n and
// 10n -> n layers of the encoder, whose rows are 40 bytes at n = 10 -- it runs on the exact-fp32
// CUDA-core kernel of gemm_simt.cuh.  In the whole-stack calls the first and last layer always take the
// CUDA-core kernel (they are <2% of the flops) and every hidden activation lives in the tensor-core
// operand format ((hi, lo) planes in 3xTF32 mode) so that no conversion pass is ever needed.
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "skinny.cuh"

namespace clica {

int simt_gemm(const SimtGemmParams& q, bool a_kc, bool b_kc, int splits, cudaStream_t st) {
    dim3 grid(ceil_div(q.N, kSBN), ceil_div(q.M, kSBM), splits);
    LaunchScope ls(st, kFamGemmSimt);
    if (a_kc && b_kc) launch_k(gemm_simt_kernel<true, true>, grid, 256, 0, st, q);
    else if (a_kc && !b_kc) launch_k(gemm_simt_kernel<true, false>, grid, 256, 0, st, q);
    else if (!a_kc && b_kc) launch_k(gemm_simt_kernel<false, true>, grid, 256, 0, st, q);
    else launch_k(gemm_simt_kernel<false, false>, grid, 256, 0, st, q);
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

int simt_colsum(const float* hi, const float* lo, int ld, int M, int N, float* db, cudaStream_t st, bool prezeroed) {
    if (!prezeroed) CLICA_CUDA_OK(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
    int ysplit = ceil_div(M, 256);
    if (ysplit > 64) ysplit = 64;
    const int rows_per_block = ceil_div(M, ysplit);
    LaunchScope ls(st, kFamMisc);
    launch_k(colsum_kernel, dim3(ceil_div(N, 32), ysplit), dim3(32, 8), 0, st, hi, lo, ld, M, N, rows_per_block, db);
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

namespace {

// ---- skinny-layer launchers (skinny.cuh) ----------------------------------------------------------------
int launch_skinny_kin(const SkinnyKinParams& q, cudaStream_t st) {
    const dim3 grid(ceil_div(q.M, kKinRows), ceil_div(q.N, kKinCols));
    LaunchScope ls(st, kFamGemmSimt);
    if (q.K <= 16) launch_k(skinny_kin_kernel<16>, grid, 256, 0, st, q);
    else launch_k(skinny_kin_kernel<48>, grid, 256, 0, st, q);
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}
int launch_skinny_nout_small(const SkinnyNoutParams& q, cudaStream_t st) {
    const size_t smem = nout_small_smem_bytes(q.N);
    static PerDeviceOnce attr;
    if (first_on_this_device(attr)) {
        CLICA_CUDA_OK(cudaFuncSetAttribute(skinny_nout_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)nout_small_smem_bytes(kNsMaxN)));
    }
    LaunchScope ls(st, kFamGemmSimt);
    launch_k(skinny_nout_small_kernel, ceil_div(q.M, kNsRows), 256, smem, st, q);
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}
bool skinny_nout_ok(int N, int K) { return N <= 48 && (size_t)N * K * sizeof(float) <= 96 * 1024; }
int launch_skinny_nout(const SkinnyNoutParams& q, int sm_count, cudaStream_t st) {
    const size_t smem = (size_t)q.N * q.K * sizeof(float);
    int grid = ceil_div(q.M, 8);
    if (grid > 8 * sm_count) grid = 8 * sm_count;
    static PerDeviceOnce attr;
    if (first_on_this_device(attr)) {
        CLICA_CUDA_OK(cudaFuncSetAttribute(skinny_nout_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CLICA_CUDA_OK(cudaFuncSetAttribute(skinny_nout_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    }
    LaunchScope ls(st, kFamGemmSimt);
    if (q.N <= 16) launch_k(skinny_nout_kernel<16>, grid, 256, smem, st, q);
    else launch_k(skinny_nout_kernel<48>, grid, 256, smem, st, q);
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}
bool skinny_dw_ok(int N, int K) { return (size_t)kSkRows * (N + K + 1) * sizeof(float) <= 64 * 1024 && (size_t)N * (K + 1) <= 32768; }
int launch_skinny_dw(const SkinnyDwParams& q, int sm_count, cudaStream_t st) {
    const size_t smem = (size_t)kSkRows * (q.N + q.K + 1) * sizeof(float);
    static PerDeviceOnce attr;
    if (first_on_this_device(attr)) {
        CLICA_CUDA_OK(cudaFuncSetAttribute(skinny_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CLICA_CUDA_OK(cudaFuncSetAttribute(skinny_dw_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    }
    LaunchScope ls(st, kFamGemmSimt);
    const int tiles = ceil_div(q.M, kSkRows);
    if ((size_t)q.N * (q.K + 1) <= (size_t)kDwsMaxOut) {
        // two tiles per CTA (halves the atomics of the tile-at-a-time kernel) unless that leaves > 2 CTAs per SM
        int per_cta = ceil_div(tiles, 2 * sm_count);
        if (per_cta < 2) per_cta = 2;
        launch_k(skinny_dw_small_kernel, ceil_div(tiles, per_cta), 256, smem, st, q);
    } else {
        launch_k(skinny_dw_kernel, tiles, 256, smem, st, q);
    }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

// y = act(x W^T + b); x, y given as (hi, lo) pairs (lo nullable)
int simt_fwd(PlanesIn x, const float* W, int ldw, const float* b, PlanesOut y, int M, int K, int N, float slope,
             int sm_count, cudaStream_t st) {
    if (K <= 48) {
        SkinnyKinParams s = {};
        s.x_hi = x.hi; s.x_lo = x.lo; s.ldx = x.ld; s.B = W; s.b_sk = 1; s.b_sn = ldw; s.bias = b; s.slope = slope;
        s.o_hi = y.hi; s.o_lo = y.lo; s.ldo = y.ld; s.M = M; s.N = N; s.K = K; s.epi = 0;
        return launch_skinny_kin(s, st);
    }
    if (y.lo == nullptr && (N <= kNsMaxN || skinny_nout_ok(N, K))) {
        SkinnyNoutParams s = {};
        s.x_hi = x.hi; s.x_lo = x.lo; s.ldx = x.ld; s.W = W; s.ldw = ldw; s.bias = b; s.out = y.hi; s.ldo = y.ld;
        s.slope = slope; s.M = M; s.N = N; s.K = K;
        if (N <= kNsMaxN) return launch_skinny_nout_small(s, st);
        return launch_skinny_nout(s, sm_count, st);
    }
    SimtGemmParams q = {};
    q.A = x.hi; q.A_lo = x.lo; q.a_sm = x.ld; q.a_sk = 1;
    q.B = W; q.B_lo = nullptr; q.b_sk = 1; q.b_sn = ldw;
    q.C = y.hi; q.C_lo = y.lo; q.ldc = y.ld; q.M = M; q.N = N; q.K = K; q.k_chunk = K;
    q.bias = b; q.slope = slope; q.epilogue = kEpiBiasAct;
    return simt_gemm(q, true, true, 1, st);
}
// dx = (dy W) * mask(aux)
int simt_bwd_data(PlanesIn dy, const float* W, int ldw, const float* aux, int ldaux, float slope_prev, PlanesOut dx,
                  int M, int K, int N, float* colsum, cudaStream_t st) {
    if (N <= 48) {     // the reduction runs over the layer's (small) output width
        SkinnyKinParams s = {};
        s.x_hi = dy.hi; s.x_lo = dy.lo; s.ldx = dy.ld; s.B = W; s.b_sk = ldw; s.b_sn = 1; s.aux = aux; s.ldaux = ldaux;
        s.slope = slope_prev; s.o_hi = dx.hi; s.o_lo = dx.lo; s.ldo = dx.ld; s.colsum = colsum;
        s.M = M; s.N = K; s.K = N; s.epi = 1;
        return launch_skinny_kin(s, st);
    }
    SimtGemmParams q = {};
    q.A = dy.hi; q.A_lo = dy.lo; q.a_sm = dy.ld; q.a_sk = 1;      // [M x N], reduce over N
    q.B = W; q.B_lo = nullptr; q.b_sk = ldw; q.b_sn = 1;          // B(n, k) = W[n][k]
    q.C = dx.hi; q.C_lo = dx.lo; q.ldc = dx.ld; q.M = M; q.N = K; q.K = N; q.k_chunk = N;
    q.aux = aux; q.ldaux = ldaux; q.slope = slope_prev; q.epilogue = kEpiMask; q.colsum = colsum;
    return simt_gemm(q, true, false, 1, st);
}
// dW = dy^T x (split-K over the M rows, atomics into the zeroed dW), db = column sums of dy
int simt_bwd_weight(PlanesIn dy, PlanesIn x, float* dW, int lddw, float* db, int M, int K, int N, int sm_count,
                    cudaStream_t st, bool prezeroed = false) {
    if (skinny_dw_ok(N, K)) {
        if (!prezeroed) {
            CLICA_CUDA_OK(cudaMemset2DAsync(dW, (size_t)lddw * sizeof(float), 0, (size_t)K * sizeof(float), N, st));
            if (db) CLICA_CUDA_OK(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
        }
        SkinnyDwParams s = {};
        s.dy_hi = dy.hi; s.dy_lo = dy.lo; s.lddy = dy.ld; s.x_hi = x.hi; s.x_lo = x.lo; s.ldx = x.ld;
        s.dW = dW; s.lddw = lddw; s.db = db; s.M = M; s.N = N; s.K = K;
        return launch_skinny_dw(s, sm_count, st);
    }
    const int tiles = ceil_div(N, kSBM) * ceil_div(K, kSBN);
    int splits = (2 * sm_count + tiles - 1) / tiles;
    const int max_splits = ceil_div(M, 4 * kSBK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int chunk = ceil_div(ceil_div(M, splits), kSBK) * kSBK;
    splits = ceil_div(M, chunk);
    if (!prezeroed) CLICA_CUDA_OK(cudaMemset2DAsync(dW, (size_t)lddw * sizeof(float), 0, (size_t)K * sizeof(float), N, st));
    SimtGemmParams q = {};
    q.A = dy.hi; q.A_lo = dy.lo; q.a_sm = 1; q.a_sk = dy.ld;      // A(n, m) = dy[m][n]
    q.B = x.hi; q.B_lo = x.lo; q.b_sk = x.ld; q.b_sn = 1;         // B(m, k) = x[m][k]
    q.C = dW; q.C_lo = nullptr; q.ldc = lddw; q.M = N; q.N = K; q.K = M; q.k_chunk = chunk;
    q.slope = 1.f; q.epilogue = kEpiAtomic;
    int rc = simt_gemm(q, false, false, splits, st);
    if (rc) return rc;
    if (db) return simt_colsum(dy.hi, dy.lo, dy.ld, M, N, db, st, prezeroed);
    return 0;
}

bool tc_mode(int mode) { return mode == CLICA_GEMM_3XTF32 || mode == CLICA_GEMM_TF32; }
bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

int check_mode(int mode) {
    CLICA_REQUIRE(mode == CLICA_GEMM_3XTF32 || mode == CLICA_GEMM_TF32 || mode == CLICA_GEMM_FP32,
                  CLICA_E_BADARG, "unknown GEMM mode %d", mode);
    return 0;
}

// ---- whole-stack planning ------------------------------------------------------------------------------
struct MlpPlan {
    int L; const int* w; int M; int mode;
    int nplanes;                     // planes per hidden activation (2 in 3xTF32 mode, else 1)
    bool planar;                     // hidden activations use plane_ld() pitches
    // Hidden activations and the gradients flowing between layers: stored (hi, lo) planes in 3xTF32 mode (default), or
    // ONE fp32 plane that the tensor-core kernel splits in shared memory (CLICA_TC_SINGLE_PLANE=1; gemm_tc.cu, converter
    // warps) -- half the HBM bytes but measured slower on B200 (profiles/r2_gemm_ab.md).  The dense input of the first
    // layer and the dense output gradient of the last one always go through the converter kernels when those layers
    // run on the tensor cores.  Weights are pre-split once per optimizer step.
    bool layer_eligible(int l) const {      // M-independent: decides the packed-weight layout
        if (!tc_mode(mode) || !tc_shape_ok(32, w[l + 1], w[l])) return false;
        // first / last layer: the dense input / output rows must be legal TMA strides (multiples of 16 bytes)
        if (l == 0 && (w[0] % 4) != 0) return false;
        if (l == L - 1 && (w[L] % 4) != 0) return false;
        return true;
    }
    bool layer_tc(int l) const { return layer_eligible(l) && tc_shape_ok(M, w[l + 1], w[l]); }
    int act_ld(int l) const { return planar ? plane_ld(w[l]) : w[l]; }
    int act_planes() const { return (mode == CLICA_GEMM_3XTF32 && env_flag("CLICA_TC_SINGLE_PLANE", 0) == 0) ? 2 : 1; }
    size_t act_floats(int l) const { return (size_t)act_planes() * M * act_ld(l); }
    size_t wplane_floats(int l) const { return layer_eligible(l) ? (size_t)nplanes * w[l + 1] * plane_ld(w[l]) : 0; }
    int nterms() const { return mode == CLICA_GEMM_3XTF32 ? 3 : 1; }
};
MlpPlan make_plan(int L, const int* widths, int M, int mode) {
    MlpPlan p;
    p.L = L; p.w = widths; p.M = M; p.mode = mode;
    p.planar = tc_mode(mode);
    p.nplanes = (mode == CLICA_GEMM_3XTF32) ? 2 : 1;
    return p;
}
PlanesIn act_in(const MlpPlan& p, const float* const* acts, int l) {
    PlanesIn a;
    a.hi = acts[l]; a.lo = nullptr;
    a.ld = (l == 0 || l == p.L) ? p.w[l] : p.act_ld(l);
    if (l != 0 && l != p.L && p.act_planes() == 2) a.lo = acts[l] + (size_t)p.M * a.ld;
    return a;
}
PlanesOut as_out(PlanesIn a) { PlanesOut o; o.hi = (float*)a.hi; o.lo = (float*)a.lo; o.ld = a.ld; return o; }

struct MlpWs { float* wplanes[64]; float* gbuf[64]; void* flags; size_t flag_bytes; size_t bytes; size_t packed_bytes; };
// `packed` (nullable): caller-provided packed weight planes (clica_mlp_pack_weights); otherwise they live in ws.
// gbuf[l] holds dL/d(pre-activation of layer l), l = 0 .. L-2 -- one buffer per layer, because the chained backward
// (gemm_tc.cu, ChainParams) lets the GEMMs of neighbouring layers overlap: a ping-pong pair would be overwritten while
// the weight-gradient GEMM of an earlier layer still reads it.
MlpWs carve_mlp(const MlpPlan& p, void* ws, const float* packed = nullptr) {
    MlpWs w;
    char* base = (char*)ws;
    size_t off = 0, woff = 0;
    for (int l = 0; l < p.L && l < 64; ++l) {
        w.wplanes[l] = packed ? (float*)((char*)packed + woff) : (float*)(base + woff);
        woff += align_up(p.wplane_floats(l) * sizeof(float), 1024);
    }
    w.packed_bytes = woff;
    off = woff;          // (the region stays reserved in ws either way: one size formula for every caller)
    for (int l = 0; l < 64; ++l) w.gbuf[l] = nullptr;
    for (int l = 0; l + 1 < p.L && l < 64; ++l) {
        w.gbuf[l] = (float*)(base + off);
        off += align_up(p.act_floats(l + 1) * sizeof(float), 1024);
    }
    w.flags = (void*)(base + off);
    w.flag_bytes = align_up(tc_chain_flag_bytes(16, p.M), 1024);
    off += w.flag_bytes;
    w.bytes = off;
    return w;
}
PlanesIn weight_planes(const MlpPlan& p, const MlpWs& w, int l) {
    PlanesIn a;
    a.ld = plane_ld(p.w[l]);
    a.hi = w.wplanes[l];
    a.lo = (p.nplanes == 2) ? w.wplanes[l] + (size_t)p.w[l + 1] * a.ld : nullptr;
    return a;
}
int pack_weights(const MlpPlan& p, const MlpWs& w, const float* const* W, cudaStream_t st) {
    if (env_flag("CLICA_PACK_FUSED", 1) != 0) {      // all eligible layers in one launch
        SplitJob jobs[64];
        int n = 0;
        for (int l = 0; l < p.L && l < 64; ++l) {
            if (!p.layer_eligible(l)) continue;
            PlanesIn wp = weight_planes(p, w, l);
            jobs[n++] = SplitJob{W[l], p.w[l], p.w[l + 1], p.w[l], (float*)wp.hi, (float*)wp.lo, wp.ld};
        }
        return tc_split_planes_multi(jobs, n, st);
    }
    for (int l = 0; l < p.L; ++l) {
        if (!p.layer_eligible(l)) continue;
        PlanesIn wp = weight_planes(p, w, l);
        int rc = tc_split_planes(W[l], p.w[l], p.w[l + 1], p.w[l], (float*)wp.hi, (float*)wp.lo, wp.ld, st);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace
}  // namespace clica

using namespace clica;

extern "C" size_t clica_linear_workspace_bytes(int M, int N, int K, int mode) {
    if (!tc_mode(mode)) return 0;
    return tc_workspace_bytes(M, N, K, mode);
}

extern "C" int clica_linear_act_fwd(const float* x, int ldx, const float* W, int ldw, const float* b,
                                    float* y, int ldy, int M, int K, int N, float slope, int mode,
                                    void* ws, size_t ws_bytes, void* stream) {
    int rc = check_mode(mode);
    if (rc) return rc;
    CLICA_REQUIRE(x && W && y, CLICA_E_BADARG, "linear_act_fwd: null pointer");
    CLICA_REQUIRE(M >= 1 && K >= 1 && N >= 1 && ldx >= K && ldw >= K && ldy >= N, CLICA_E_BADARG,
                  "linear_act_fwd: bad shape M=%d K=%d N=%d ldx=%d ldw=%d ldy=%d", M, K, N, ldx, ldw, ldy);
    DeviceInfo di;
    if ((rc = get_device_info(&di))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (tc_mode(mode) && tc_shape_ok(M, N, K) && ldy % 4 == 0 && aligned16(y)) {
        CLICA_REQUIRE(ws && ws_bytes >= tc_workspace_bytes(M, N, K, mode), CLICA_E_WORKSPACE, "linear_act_fwd: workspace too small");
        return tc_linear_fwd(x, ldx, W, ldw, b, y, ldy, M, K, N, slope, mode, ws, ws_bytes, di.sm_count, st);
    }
    return simt_fwd(PlanesIn{x, nullptr, ldx}, W, ldw, b, PlanesOut{y, nullptr, ldy}, M, K, N, slope, di.sm_count, st);
}

extern "C" int clica_linear_act_bwd_data(const float* dy, int lddy, const float* W, int ldw,
                                         const float* x_act, int ldxa, float slope_prev,
                                         float* dx, int lddx, int M, int K, int N, int mode,
                                         void* ws, size_t ws_bytes, void* stream) {
    int rc = check_mode(mode);
    if (rc) return rc;
    CLICA_REQUIRE(dy && W && dx, CLICA_E_BADARG, "linear_act_bwd_data: null pointer");
    CLICA_REQUIRE(M >= 1 && K >= 1 && N >= 1 && lddy >= N && ldw >= K && lddx >= K && (!x_act || ldxa >= K),
                  CLICA_E_BADARG, "linear_act_bwd_data: bad shape");
    DeviceInfo di;
    if ((rc = get_device_info(&di))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (tc_mode(mode) && tc_shape_ok(M, K, N) && lddx % 4 == 0 && aligned16(dx)) {
        CLICA_REQUIRE(ws && ws_bytes >= tc_workspace_bytes(M, N, K, mode), CLICA_E_WORKSPACE, "linear_act_bwd_data: workspace too small");
        return tc_linear_bwd_data(dy, lddy, W, ldw, x_act, ldxa, slope_prev, dx, lddx, M, K, N, mode, ws,
                                  ws_bytes, di.sm_count, st);
    }
    return simt_bwd_data(PlanesIn{dy, nullptr, lddy}, W, ldw, x_act, ldxa, slope_prev, PlanesOut{dx, nullptr, lddx}, M, K, N, nullptr, st);
}

extern "C" int clica_linear_bwd_weight(const float* dy, int lddy, const float* x, int ldx,
                                       float* dW, int lddw, float* db, int M, int K, int N, int mode,
                                       void* ws, size_t ws_bytes, void* stream) {
    int rc = check_mode(mode);
    if (rc) return rc;
    CLICA_REQUIRE(dy && x && dW, CLICA_E_BADARG, "linear_bwd_weight: null pointer");
    CLICA_REQUIRE(M >= 1 && K >= 1 && N >= 1 && lddy >= N && ldx >= K && lddw >= K, CLICA_E_BADARG,
                  "linear_bwd_weight: bad shape");
    DeviceInfo di;
    if ((rc = get_device_info(&di))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (tc_mode(mode) && tc_shape_ok(N, K, M)) {
        CLICA_REQUIRE(ws && ws_bytes >= tc_workspace_bytes(M, N, K, mode), CLICA_E_WORKSPACE, "linear_bwd_weight: workspace too small");
        return tc_linear_bwd_weight(dy, lddy, x, ldx, dW, lddw, db, M, K, N, mode, ws, ws_bytes, di.sm_count, st);
    }
    return simt_bwd_weight(PlanesIn{dy, nullptr, lddy}, PlanesIn{x, nullptr, ldx}, dW, lddw, db, M, K, N, di.sm_count, st);
}

// ---- whole-stack calls ----------------------------------------------------------------------------
extern "C" size_t clica_mlp_act_floats(int M, int width, int mode) {
    if (M < 1 || width < 1) return 0;
    if (!tc_mode(mode)) return (size_t)M * width;
    const int planes = (mode == CLICA_GEMM_3XTF32 && env_flag("CLICA_TC_SINGLE_PLANE", 0) == 0) ? 2 : 1;
    return (size_t)planes * M * plane_ld(width);        // (hi, lo) planes by default (see MlpPlan)
}

extern "C" size_t clica_mlp_workspace_bytes(int M, int L, const int* widths, int mode) {
    if (L < 1 || L > 64 || !widths || M < 1) return 0;
    MlpPlan p = make_plan(L, widths, M, mode);
    return carve_mlp(p, nullptr).bytes + 1024;
}

extern "C" size_t clica_mlp_packed_weight_bytes(int L, const int* widths, int mode) {
    if (L < 1 || L > 64 || !widths) return 0;
    MlpPlan p = make_plan(L, widths, 32, mode);
    return carve_mlp(p, nullptr).packed_bytes + 1024;
}

// Where clica_mlp_pack_weights puts layer l inside the packed buffer: byte offsets of its hi / lo planes (-1: the layer is
// not packed / has no lo plane) and the planes' row pitch in floats.  Lets an optimizer keep the planes current itself
// (clica_adam_step_capturable_packed).
extern "C" int clica_mlp_packed_weight_layout(int L, const int* widths, int mode, long long* hi_off, long long* lo_off, int* ld) {
    int rc = check_mode(mode);
    if (rc) return rc;
    CLICA_REQUIRE(L >= 1 && L <= 64 && widths && hi_off && lo_off && ld, CLICA_E_BADARG, "mlp_packed_weight_layout: bad arguments");
    MlpPlan p = make_plan(L, widths, 32, mode);
    MlpWs w = carve_mlp(p, nullptr, (const float*)nullptr);
    for (int l = 0; l < L; ++l) {
        hi_off[l] = -1; lo_off[l] = -1; ld[l] = 0;
        if (!p.layer_eligible(l)) continue;
        PlanesIn wp = weight_planes(p, w, l);
        hi_off[l] = (long long)(uintptr_t)wp.hi;
        lo_off[l] = wp.lo ? (long long)(uintptr_t)wp.lo : -1;
        ld[l] = wp.ld;
    }
    return 0;
}

extern "C" int clica_mlp_pack_weights(int L, const int* widths, const float* const* W, int mode, void* packed,
                                      size_t packed_bytes, void* stream) {
    int rc = check_mode(mode);
    if (rc) return rc;
    CLICA_REQUIRE(L >= 1 && L <= 64 && widths && W && packed, CLICA_E_BADARG, "mlp_pack_weights: bad arguments");
    CLICA_REQUIRE((((uintptr_t)packed) & 1023u) == 0, CLICA_E_ALIGN, "mlp_pack_weights: buffer must be 1024-byte aligned");
    MlpPlan p = make_plan(L, widths, 32, mode);
    MlpWs w = carve_mlp(p, nullptr, (const float*)packed);
    CLICA_REQUIRE(packed_bytes >= w.packed_bytes, CLICA_E_WORKSPACE, "mlp_pack_weights: buffer %zu < %zu bytes", packed_bytes, w.packed_bytes);
    return pack_weights(p, w, W, (cudaStream_t)stream);
}

extern "C" int clica_mlp_fwd(int L, const int* widths, const float* const* W, const float* const* b,
                             float* const* acts, int M, float slope, int mode, const void* packed_weights,
                             void* ws, size_t ws_bytes, void* stream) {
    int rc = check_mode(mode);
    if (rc) return rc;
    CLICA_REQUIRE(L >= 1 && L <= 64 && widths && W && b && acts && M >= 1, CLICA_E_BADARG, "mlp_fwd: bad arguments");
    DeviceInfo di;
    if ((rc = get_device_info(&di))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    MlpPlan p = make_plan(L, widths, M, mode);
    void* wsa = (void*)align_up((size_t)(uintptr_t)ws, 1024);
    MlpWs w = carve_mlp(p, wsa, (const float*)packed_weights);
    CLICA_REQUIRE(ws && ws_bytes >= w.bytes + 1024, CLICA_E_WORKSPACE, "mlp_fwd: workspace %zu < %zu bytes", ws_bytes, w.bytes + 1024);
    if (!packed_weights && (rc = pack_weights(p, w, W, st))) return rc;
    // consecutive tensor-core layers are handed to tc_chain_launch together: ONE persistent launch walks their tiles, a
    // layer's tile waiting only for the 256-row block of the previous layer's output it reads (gemm_tc.cu, ChainParams)
    static thread_local TcChainLink links[kTcMaxLinks];
    int nl = 0, prev_layer = -2;
    auto flush = [&]() -> int {
        const int r = nl ? tc_chain_launch(links, nl, di.sm_count, w.flags, w.flag_bytes, st) : 0;
        nl = 0; prev_layer = -2;
        return r;
    };
    for (int l = 0; l < L; ++l) {
        const int K = widths[l], N = widths[l + 1];
        const float s = (l == L - 1) ? 1.f : slope;
        PlanesIn x = act_in(p, acts, l);
        PlanesOut y = as_out(act_in(p, acts, l + 1));
        if (p.layer_tc(l) && aligned16(x.hi) && aligned16(y.hi)) {
            TcGemm g = {};
            g.A = x; g.a_mn_major = 0; g.B = weight_planes(p, w, l); g.b_mn_major = 0; g.nterms = p.nterms();
            g.Mo = M; g.No = N; g.Kr = K; g.epi = kTcBiasAct; g.bias = b[l]; g.slope = s;
            g.outp = y;        // one fp32 plane through the TMA (the dense [M, N] output of the last layer as well)
            if (nl == kTcMaxLinks && (rc = flush())) return rc;
            links[nl] = TcChainLink{g, (prev_layer == l - 1) ? nl - 1 : -1, 0};
            ++nl; prev_layer = l;
        } else {
            if ((rc = flush())) return rc;
            rc = simt_fwd(x, W[l], K, b[l], y, M, K, N, s, di.sm_count, st);
        }
        if (rc) return rc;
    }
    return flush();
}

extern "C" int clica_mlp_bwd(int L, const int* widths, const float* const* W, const float* const* acts,
                             const float* g_out, float* const* dW, float* const* db, float* g_in,
                             int M, float slope, int mode, const void* packed_weights, int grads_prezeroed,
                             void* ws, size_t ws_bytes, void* stream) {
    return clica_mlp_bwd_range(L, widths, W, acts, g_out, dW, db, g_in, M, slope, mode, packed_weights,
                               grads_prezeroed, L - 1, 0, ws, ws_bytes, stream);
}

// Layers l_first, l_first - 1, ..., l_last of the backward chain.  The gradient w.r.t. the pre-activation of layer
// l_first comes from g_out (l_first == L - 1) or from the workspace, where the previous range call left it: a whole
// backward may be issued as consecutive ranges with the SAME workspace on the SAME stream (the multi-GPU step does so
// to start the all-reduce of a finished layer's gradients while the earlier layers are still being differentiated).
extern "C" int clica_mlp_bwd_range(int L, const int* widths, const float* const* W, const float* const* acts,
                                   const float* g_out, float* const* dW, float* const* db, float* g_in,
                                   int M, float slope, int mode, const void* packed_weights, int grads_prezeroed,
                                   int l_first, int l_last, void* ws, size_t ws_bytes, void* stream) {
    int rc = check_mode(mode);
    if (rc) return rc;
    CLICA_REQUIRE(L >= 1 && L <= 64 && widths && W && acts && g_out && dW && db && M >= 1, CLICA_E_BADARG, "mlp_bwd: bad arguments");
    CLICA_REQUIRE(l_first <= L - 1 && l_last >= 0 && l_last <= l_first, CLICA_E_BADARG,
                  "mlp_bwd_range: layers [%d .. %d] outside [%d .. 0]", l_first, l_last, L - 1);
    DeviceInfo di;
    if ((rc = get_device_info(&di))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    MlpPlan p = make_plan(L, widths, M, mode);
    void* wsa = (void*)align_up((size_t)(uintptr_t)ws, 1024);
    MlpWs w = carve_mlp(p, wsa, (const float*)packed_weights);
    CLICA_REQUIRE(ws && ws_bytes >= w.bytes + 1024, CLICA_E_WORKSPACE, "mlp_bwd: workspace %zu < %zu bytes", ws_bytes, w.bytes + 1024);
    if (!packed_weights && (rc = pack_weights(p, w, W, st))) return rc;
    const bool pz = grads_prezeroed != 0;

    PlanesIn g = {g_out, nullptr, widths[L]};     // dL/d(pre-activation of layer l), starts as dL/d(output)
    bool db_done = false;                         // db[l] already accumulated by the epilogue that produced g
    if (l_first < L - 1) {                        // continue where the previous range stopped
        g.ld = p.act_ld(l_first + 1);
        g.hi = w.gbuf[l_first];
        g.lo = (p.act_planes() == 2) ? g.hi + (size_t)M * g.ld : nullptr;
        db_done = true;
    }
    // the tensor-core GEMMs of consecutive layers -- dX_l (produces the next layer's gradient) then dW_l, l descending --
    // are collected and issued as ONE chained launch (gemm_tc.cu, ChainParams): dX_{l-1} and dW_{l-1} wait per 256-row
    // block for dX_l's tiles, with dW_l's tiles between them in the work list to cover that latency
    static thread_local TcChainLink links[kTcMaxLinks];
    int nl = 0, g_link = -1;                      // g_link: the link whose output is the current g (-1: not in the list)
    auto flush = [&]() -> int {
        const int r = nl ? tc_chain_launch(links, nl, di.sm_count, w.flags, w.flag_bytes, st) : 0;
        nl = 0; g_link = -1;
        return r;
    };
    for (int l = l_first; l >= l_last; --l) {
        const int K = widths[l], N = widths[l + 1];
        PlanesIn x = act_in(p, acts, l);
        const bool tc_l = p.layer_tc(l) && aligned16(g.hi) && aligned16(x.hi);
        if (!tc_l && (rc = flush())) return rc;
        if (nl + 2 > kTcMaxLinks && (rc = flush())) return rc;
        // g_prev = (g W[l]) * LeakyReLU'(pre-activation of layer l-1); acts[l] = LeakyReLU output: same sign
        PlanesOut gp = {};
        int dx_link = -1;
        if (l > 0) {
            gp.ld = p.act_ld(l);
            gp.hi = w.gbuf[l - 1];
            gp.lo = (p.act_planes() == 2) ? gp.hi + (size_t)M * gp.ld : nullptr;
            // the epilogue that writes g_prev also accumulates its column sums = db[l-1] (no separate reduction pass)
            if (!pz) CLICA_CUDA_OK(cudaMemsetAsync(db[l - 1], 0, (size_t)K * sizeof(float), st));
            if (tc_l) {
                TcGemm t = {};
                t.A = g; t.a_mn_major = 0; t.B = weight_planes(p, w, l); t.b_mn_major = 1; t.nterms = p.nterms();
                t.Mo = M; t.No = K; t.Kr = N; t.epi = kTcMask; t.aux = x.hi; t.ldaux = x.ld; t.slope = slope; t.outp = gp;
                t.colsum = db[l - 1];
                links[nl] = TcChainLink{t, g_link, 0};
                dx_link = nl++;
            }
        }
        // dW[l] = g^T x ; db[l] = column sums of g
        if (tc_l) {
            if (!pz) CLICA_CUDA_OK(cudaMemsetAsync(dW[l], 0, (size_t)N * K * sizeof(float), st));
            TcGemm t = {};
            t.A = g; t.a_mn_major = 1; t.B = x; t.b_mn_major = 1; t.nterms = p.nterms();
            t.Mo = N; t.No = K; t.Kr = M; t.epi = kTcAtomic; t.out = dW[l]; t.ldo = K; t.allow_split_k = 1;
            links[nl++] = TcChainLink{t, g_link, 1};
            if (!db_done) {
                // g is the caller's dense output gradient here (never a link's output): its column sums can run right away
                if ((rc = simt_colsum(g.hi, g.lo, g.ld, M, N, db[l], st, pz))) return rc;
            }
        } else {
            if ((rc = simt_bwd_weight(g, x, dW[l], K, db_done ? nullptr : db[l], M, K, N, di.sm_count, st, pz))) return rc;
        }
        if (l == 0) {
            if (g_in) {
                if ((rc = flush())) return rc;
                rc = simt_bwd_data(g, W[0], K, nullptr, 0, 1.f, PlanesOut{g_in, nullptr, K}, M, K, N, nullptr, st);
            }
            if (rc) return rc;
            break;
        }
        if (!tc_l) {
            if ((rc = simt_bwd_data(g, W[l], K, x.hi, x.ld, slope, gp, M, K, N, db[l - 1], st))) return rc;
        }
        g_link = dx_link;
        db_done = true;
        g.hi = gp.hi; g.lo = gp.lo; g.ld = gp.ld;
    }
    return flush();
}
