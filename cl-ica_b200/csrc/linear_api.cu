// linear_api.cu -- C-ABI entry points of the encoder layers (see include/clica.h): per-layer
// Linear(+LeakyReLU) forward / backward-data / backward-weight and the whole-stack mlp_fwd / mlp_bwd.
// Replaces nn.Linear + nn.LeakyReLU of /root/reference/encoders.py:38-48 and their autograd nodes.
//
// Shape rule (not a fallback): a layer goes to the tcgen05 tensor-core GEMM (gemm_tc.cu) when the
// requested mode is a tensor-core mode AND TMA can address all three matrices (leading dimensions
// multiples of 4 floats, 16-byte aligned bases) AND both K and N are >= 16; otherwise -- the n -> 10n and
// 10n -> n layers of the encoder, whose rows are 40 bytes at n = 10 -- it runs on the exact-fp32
// CUDA-core kernel of gemm_simt.cuh.
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace clica {

int simt_linear_fwd(const float* x, int ldx, const float* W, int ldw, const float* b, float* y, int ldy,
                    int M, int K, int N, float slope, cudaStream_t st) {
    SimtGemmParams q;
    q.A = x; q.a_sm = ldx; q.a_sk = 1;
    q.B = W; q.b_sk = 1; q.b_sn = ldw;
    q.C = y; q.ldc = ldy; q.M = M; q.N = N; q.K = K; q.k_chunk = K;
    q.bias = b; q.aux = nullptr; q.ldaux = 0; q.slope = slope; q.epilogue = kEpiBiasAct;
    dim3 grid(ceil_div(N, kSBN), ceil_div(M, kSBM), 1);
    { LaunchScope ls(st, kFamGemmSimt); gemm_simt_kernel<true, true><<<grid, 256, 0, st>>>(q); }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

int simt_linear_bwd_data(const float* dy, int lddy, const float* W, int ldw, const float* x_act, int ldxa,
                         float slope_prev, float* dx, int lddx, int M, int K, int N, cudaStream_t st) {
    SimtGemmParams q;
    q.A = dy; q.a_sm = lddy; q.a_sk = 1;          // [M x N], reduce over N
    q.B = W; q.b_sk = ldw; q.b_sn = 1;            // B(n, k) = W[n][k]
    q.C = dx; q.ldc = lddx; q.M = M; q.N = K; q.K = N; q.k_chunk = N;
    q.bias = nullptr; q.aux = x_act; q.ldaux = ldxa; q.slope = slope_prev; q.epilogue = kEpiMask;
    dim3 grid(ceil_div(K, kSBN), ceil_div(M, kSBM), 1);
    { LaunchScope ls(st, kFamGemmSimt); gemm_simt_kernel<true, false><<<grid, 256, 0, st>>>(q); }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

int simt_linear_bwd_weight(const float* dy, int lddy, const float* x, int ldx, float* dW, int lddw, float* db,
                           int M, int K, int N, int sm_count, cudaStream_t st) {
    // dW[N x K] = dy^T x, reduction over the M rows: split-K so that the grid fills the SMs
    const int tiles = ceil_div(N, kSBM) * ceil_div(K, kSBN);
    int splits = (2 * sm_count + tiles - 1) / tiles;
    const int max_splits = ceil_div(M, 4 * kSBK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int chunk = ceil_div(ceil_div(M, splits), kSBK) * kSBK;
    splits = ceil_div(M, chunk);
    CLICA_CUDA_OK(cudaMemset2DAsync(dW, (size_t)lddw * sizeof(float), 0, (size_t)K * sizeof(float), N, st));
    SimtGemmParams q;
    q.A = dy; q.a_sm = 1; q.a_sk = lddy;          // A(n, m) = dy[m][n]
    q.B = x; q.b_sk = ldx; q.b_sn = 1;            // B(m, k) = x[m][k]
    q.C = dW; q.ldc = lddw; q.M = N; q.N = K; q.K = M; q.k_chunk = chunk;
    q.bias = nullptr; q.aux = nullptr; q.ldaux = 0; q.slope = 1.f; q.epilogue = kEpiAtomic;
    dim3 grid(ceil_div(K, kSBN), ceil_div(N, kSBM), splits);
    { LaunchScope ls(st, kFamGemmSimt); gemm_simt_kernel<false, false><<<grid, 256, 0, st>>>(q); }
    CLICA_CUDA_OK(cudaGetLastError());
    if (db) {
        { LaunchScope ls(st, kFamMisc); colsum_kernel<<<ceil_div(N, 32), dim3(32, 8), 0, st>>>(dy, lddy, M, N, db); }
        CLICA_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

namespace {

bool tc_mode(int mode) { return mode == CLICA_GEMM_3XTF32 || mode == CLICA_GEMM_TF32; }
bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }

int check_mode(int mode) {
    CLICA_REQUIRE(mode == CLICA_GEMM_3XTF32 || mode == CLICA_GEMM_TF32 || mode == CLICA_GEMM_FP32,
                  CLICA_E_BADARG, "unknown GEMM mode %d", mode);
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
    if (tc_mode(mode) && tc_shape_ok(M, N, K) && ldx % 4 == 0 && ldw % 4 == 0 && ldy % 4 == 0 &&
        aligned16(x) && aligned16(W) && aligned16(y))
        return tc_linear_fwd(x, ldx, W, ldw, b, y, ldy, M, K, N, slope, mode, ws, ws_bytes, di.sm_count, st);
    return simt_linear_fwd(x, ldx, W, ldw, b, y, ldy, M, K, N, slope, st);
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
    if (tc_mode(mode) && tc_shape_ok(M, K, N) && lddy % 4 == 0 && ldw % 4 == 0 && lddx % 4 == 0 &&
        (!x_act || ldxa % 4 == 0) && aligned16(dy) && aligned16(W) && aligned16(dx))
        return tc_linear_bwd_data(dy, lddy, W, ldw, x_act, ldxa, slope_prev, dx, lddx, M, K, N, mode, ws,
                                  ws_bytes, di.sm_count, st);
    return simt_linear_bwd_data(dy, lddy, W, ldw, x_act, ldxa, slope_prev, dx, lddx, M, K, N, st);
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
    if (tc_mode(mode) && tc_shape_ok(N, K, M) && lddy % 4 == 0 && ldx % 4 == 0 && lddw % 4 == 0 &&
        aligned16(dy) && aligned16(x) && aligned16(dW))
        return tc_linear_bwd_weight(dy, lddy, x, ldx, dW, lddw, db, M, K, N, mode, ws, ws_bytes, di.sm_count, st);
    return simt_linear_bwd_weight(dy, lddy, x, ldx, dW, lddw, db, M, K, N, di.sm_count, st);
}

// ---- whole-stack calls ----------------------------------------------------------------------------
extern "C" size_t clica_mlp_workspace_bytes(int M, int L, const int* widths, int mode) {
    size_t need = 0;
    size_t gbuf = 0;
    for (int l = 0; l < L; ++l) {
        size_t w = clica_linear_workspace_bytes(M, widths[l + 1], widths[l], mode);
        if (w > need) need = w;
        size_t g = (size_t)M * (size_t)widths[l] * sizeof(float);
        if (l > 0 && g > gbuf) gbuf = g;
    }
    // backward ping-pong buffers for dL/d(acts[l]) + the per-layer GEMM workspace
    return align_up(need, 1024) + 2 * align_up(gbuf, 1024);
}

extern "C" int clica_mlp_fwd(int L, const int* widths, const float* const* W, const float* const* b,
                             float* const* acts, int M, float slope, int mode,
                             void* ws, size_t ws_bytes, void* stream) {
    CLICA_REQUIRE(L >= 1 && widths && W && b && acts, CLICA_E_BADARG, "mlp_fwd: null pointer / L < 1");
    for (int l = 0; l < L; ++l) {
        const int K = widths[l], N = widths[l + 1];
        const float s = (l == L - 1) ? 1.f : slope;
        int rc = clica_linear_act_fwd(acts[l], K, W[l], K, b[l], acts[l + 1], N, M, K, N, s, mode, ws, ws_bytes, stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int clica_mlp_bwd(int L, const int* widths, const float* const* W, const float* const* acts,
                             const float* g_out, float* const* dW, float* const* db, float* g_in,
                             int M, float slope, int mode, void* ws, size_t ws_bytes, void* stream) {
    CLICA_REQUIRE(L >= 1 && widths && W && acts && g_out && dW && db, CLICA_E_BADARG, "mlp_bwd: null pointer / L < 1");
    size_t need = clica_mlp_workspace_bytes(M, L, widths, mode);
    CLICA_REQUIRE(ws_bytes >= need && (need == 0 || ws), CLICA_E_WORKSPACE, "mlp_bwd: workspace %zu < %zu bytes", ws_bytes, need);
    size_t gbuf = 0, lin = 0;
    for (int l = 0; l < L; ++l) {
        size_t w = clica_linear_workspace_bytes(M, widths[l + 1], widths[l], mode);
        if (w > lin) lin = w;
        size_t g = (size_t)M * (size_t)widths[l] * sizeof(float);
        if (l > 0 && g > gbuf) gbuf = g;
    }
    char* base = (char*)ws;
    void* lin_ws = base;
    float* gb[2] = {(float*)(base + align_up(lin, 1024)), (float*)(base + align_up(lin, 1024) + align_up(gbuf, 1024))};
    const float* g = g_out;   // dL/d acts[l+1] (already through the activation mask)
    for (int l = L - 1; l >= 0; --l) {
        const int K = widths[l], N = widths[l + 1];
        int rc = clica_linear_bwd_weight(g, N, acts[l], K, dW[l], K, db[l], M, K, N, mode, lin_ws, lin, stream);
        if (rc) return rc;
        if (l > 0) {
            float* gx = gb[l & 1];
            // acts[l] is the LeakyReLU output of layer l-1: its sign is the activation mask
            rc = clica_linear_act_bwd_data(g, N, W[l], K, acts[l], K, slope, gx, K, M, K, N, mode, lin_ws, lin, stream);
            if (rc) return rc;
            g = gx;
        } else if (g_in) {
            rc = clica_linear_act_bwd_data(g, N, W[0], K, nullptr, 0, 1.f, g_in, K, M, K, N, mode, lin_ws, lin, stream);
            if (rc) return rc;
        }
    }
    return 0;
}
