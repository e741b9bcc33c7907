// lpnce_api.cu -- C-ABI entry points of the fused Lp-InfoNCE loss (see include/clica.h) plus the small
// O(B*d) kernels around the pair-walking kernels of lpnce_kernels.cuh: finalize (merge split partials,
// positive pair, per-item loss, the three means), prep (per-anchor backward coefficients) and reduce
// (sum split partials, add the positive-pair gradient).
#include "lpnce_kernels.cuh"

namespace clica {
namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

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

struct FinParams {
    const float* part_m; const float* part_s; int part_stride; int nsplit;
    const float* z1; int ld1; const float* z2; int ld2;
    int B; int M; int d; float p; float tau; float alpha; int include_pos;
    float* loss_i; float* lse; float* pos; float2* rowstat; float* scalars;
    double* block_sums; int* counter;
    const float* z3; int ld3; int fast;   // EXPERIMENTAL fast forward: negatives, for the underflow fallback
};

__global__ void __launch_bounds__(256) lpnce_finalize_kernel(const FinParams q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float v_loss = 0.f, v_pos = 0.f, v_lse = 0.f;
    if (i < q.B) {
        // loads are issued in independent batches of 8 (a load -> use loop pays one memory latency per iteration)
        float ps = 0.f;
        const float* a = q.z1 + (size_t)i * q.ld1;
        const float* b = q.z2 + (size_t)i * q.ld2;
        for (int c0 = 0; c0 < q.d; c0 += 8) {
            float av[8], bv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = c0 + u < q.d;
                av[u] = ok ? __ldg(a + c0 + u) : 0.f;
                bv[u] = ok ? __ldg(b + c0 + u) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) ps += abs_pow(av[u] - bv[u], q.p);     // |0|^p = 0 for the padding
        }
        const float coef = kLog2e / q.tau;
        const float xp = -ps * coef;
        float M = q.include_pos ? xp : -INFINITY;
        float S = 0.f;
        // pass 1: row maximum over the split partials; pass 2 (partials now in L1/L2): rescaled sum, fixed order
        for (int s0 = 0; s0 < q.nsplit; s0 += 8) {
            float pm[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                pm[u] = (s0 + u < q.nsplit) ? q.part_m[(size_t)(s0 + u) * q.part_stride + i] : -INFINITY;
#pragma unroll
            for (int u = 0; u < 8; ++u) M = fmaxf(M, pm[u]);
        }
        if (q.include_pos) S = exp2f(xp - M);
        for (int s0 = 0; s0 < q.nsplit; s0 += 8) {
            float pm[8], psum[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const bool ok = s0 + u < q.nsplit;
                pm[u] = ok ? q.part_m[(size_t)(s0 + u) * q.part_stride + i] : 0.f;
                psum[u] = ok ? q.part_s[(size_t)(s0 + u) * q.part_stride + i] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (s0 + u < q.nsplit) S += psum[u] * exp2f(pm[u] - M);
        }
        if (q.fast && !(S >= 0x1p-80f)) {
            // fast forward (reference point 0): this row's soft-max sum is so small that terms flushed to zero could
            // matter (or everything underflowed).  Recompute it robustly: maximum first, then the re-scaled sum.
            float mx = q.include_pos ? xp : -INFINITY;
            for (int j = 0; j < q.M; ++j) {
                const float* c3 = q.z3 + (size_t)j * q.ld3;
                float D = 0.f;
                for (int c = 0; c < q.d; ++c) D += abs_pow(__ldg(a + c) - __ldg(c3 + c), q.p);
                mx = fmaxf(mx, -D * coef);
            }
            float ssum = q.include_pos ? exp2f(xp - mx) : 0.f;
            for (int j = 0; j < q.M; ++j) {
                const float* c3 = q.z3 + (size_t)j * q.ld3;
                float D = 0.f;
                for (int c = 0; c < q.d; ++c) D += abs_pow(__ldg(a + c) - __ldg(c3 + c), q.p);
                ssum += exp2f(-D * coef - mx);
            }
            M = mx; S = ssum;
        }
        const float ls = log2f(S);
        q.rowstat[i] = make_float2(M, ls);
        float l = (M + ls) * kLn2;
        if (!q.include_pos) l -= logf((float)q.M);
        const float li = 2.f * (q.alpha * ps / q.tau + (1.f - q.alpha) * l);
        q.loss_i[i] = li;
        q.lse[i] = l;
        q.pos[i] = ps;
        v_loss = li; v_pos = ps / q.tau; v_lse = l;
    }
    // deterministic: per-block sums, then the last block to finish adds them in block order
    __shared__ float red[3][8];
    __shared__ int is_last;
    v_loss = warp_sum(v_loss); v_pos = warp_sum(v_pos); v_lse = warp_sum(v_lse);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = v_loss; red[1][warp] = v_pos; red[2][warp] = v_lse; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0, s1 = 0, s2 = 0;
        for (int w = 0; w < 8; ++w) { s0 += red[0][w]; s1 += red[1][w]; s2 += red[2][w]; }
        q.block_sums[3 * blockIdx.x + 0] = s0;
        q.block_sums[3 * blockIdx.x + 1] = s1;
        q.block_sums[3 * blockIdx.x + 2] = s2;
        __threadfence();
        is_last = (atomicAdd(q.counter, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        // the last block adds the per-block sums in block order: thread t takes blocks t, t + 256, ... (parallel
        // loads), thread 0 then adds the 256 per-thread sums in thread order -- a fixed order, so deterministic
        __shared__ double fin[3][256];
        __threadfence();
        const volatile double* bs = q.block_sums;
        double s0 = 0, s1 = 0, s2 = 0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += 256) { s0 += bs[3 * b]; s1 += bs[3 * b + 1]; s2 += bs[3 * b + 2]; }
        fin[0][threadIdx.x] = s0; fin[1][threadIdx.x] = s1; fin[2][threadIdx.x] = s2;
        __syncthreads();
        if (threadIdx.x < 3) {
            double t = 0;
            const int nb = gridDim.x < 256 ? (int)gridDim.x : 256;
            for (int k = 0; k < nb; ++k) t += fin[threadIdx.x][k];
            q.scalars[threadIdx.x] = (float)(t / q.B);
        }
    }
}

// per-anchor coefficients of the backward:  E = 2 gl (1-alpha)/tau,  CP = 2 gl (alpha - (1-alpha) w+)/tau,
// gl_i = g_mean * inv_count + g_loss_i[i],  w+ = exp2((-pos*coef - m2) - ls) from the forward's row statistics
struct PrepParams {
    const float2* rowstat; const float* pos; const float* g_mean; const float* g_loss_i;
    int n; float inv_count; float tau; float alpha; int include_pos;
    float* E; float* CP;              // either may be null (CP null: pos / rowstat are not read)
    float default_g;                  // used when g_mean == nullptr
};
__global__ void lpnce_prep_kernel(const PrepParams q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.n) return;
    float gl = (q.g_mean ? *q.g_mean : q.default_g) * q.inv_count;
    if (q.g_loss_i) gl += q.g_loss_i[i];
    if (q.E) q.E[i] = 2.f * gl * (1.f - q.alpha) / q.tau;
    if (q.CP) {
        const float2 st = q.rowstat[i];
        const float wpos = q.include_pos ? exp2f(fmaf(q.pos[i], -(kLog2e / q.tau), -st.x) - st.y) : 0.f;
        q.CP[i] = 2.f * gl * (q.alpha - (1.f - q.alpha) * wpos) / q.tau;
    }
}

// g_out[row, c] = p * ( E[row] * sum_s partA[s][row][c] + sum_s partB[s][row][c] ) + CP[row] * G'(z1 - z2)
struct ReduceParams {
    const float* partA; int nsA; int rowsA;   // anchor role, scaled by E; nullable
    const float* partB; int nsB; int rowsB;   // column role, already weighted; nullable
    const float* E; const float* CP;          // CP nullable (no positive term)
    const float* z1; int ld1; const float* z2; int ld2;
    float* g_out; int ldg; float* g_z2; int ldg2;   // both nullable
    int rows; int d; int TW; float p;
};
__global__ void lpnce_reduce_kernel(const ReduceParams q) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)q.rows * q.d) return;
    const int row = (int)(idx / q.d), c = (int)(idx - (long long)row * q.d);
    float g = 0.f;
    if (q.partA) {
        float a = 0.f;
        for (int s = 0; s < q.nsA; ++s) a += q.partA[((size_t)s * q.rowsA + row) * q.TW + c];
        g = q.E[row] * a;
    }
    if (q.partB) {
        float b = 0.f;
        for (int s = 0; s < q.nsB; ++s) b += q.partB[((size_t)s * q.rowsB + row) * q.TW + c];
        g += b;
    }
    g *= q.p;
    if (q.CP) {
        const float gp = q.CP[row] * dabs_pow(q.z1[(size_t)row * q.ld1 + c] - q.z2[(size_t)row * q.ld2 + c], q.p);
        g += gp;
        if (q.g_z2) q.g_z2[(size_t)row * q.ldg2 + c] = -gp;
    }
    if (q.g_out) q.g_out[(size_t)row * q.ldg + c] = g;
}

// ---- host-side planning ---------------------------------------------------------------------------
// feature pairs per lane (DP) and lanes per pair (F): d <= 40 is held by one lane; wider d is split over
// F in {2, 4, 8} adjacent lanes with DP in {16, 20} (d <= 320)
struct Shape { int DP; int F; };
const int kDpList[] = {2, 3, 4, 5, 8, 12, 16, 20};
Shape pick_shape(int d) {
    const int need = (d + 1) / 2;
    for (int v : kDpList) if (v >= need) return Shape{v, 1};
    for (int f = 2; f <= 8; f *= 2)
        for (int v : {16, 20}) if (v * f >= need) return Shape{v, f};
    return Shape{-1, 0};
}
int pick_dp(int d) { return pick_shape(d).DP; }
int p_code(float p) {
    if (p == 1.f) return 1;
    if (p == 2.f) return 2;
    if (p == 3.f) return 3;
    if (p == 4.f) return 4;
    return 0;
}

struct SplitPlan { int row_tiles; int tiles_per_split; int nsplit; };
SplitPlan plan_splits(int rows, int rows_cta, int MS, int tile_cols, int sm_count, int ctas_per_sm) {
    SplitPlan pl;
    pl.row_tiles = ceil_div(rows, rows_cta);
    const int col_tiles = ceil_div(MS, tile_cols);
    int want = (sm_count * ctas_per_sm) / (pl.row_tiles > 0 ? pl.row_tiles : 1);
    if (want < 1) want = 1;
    if (want > col_tiles) want = col_tiles;
    if (want > 64) want = 64;
    pl.tiles_per_split = ceil_div(col_tiles, want);
    pl.nsplit = ceil_div(col_tiles, pl.tiles_per_split);
    return pl;
}
// resident CTAs per SM of the kernel that will run (cudaOccupancy..., cached per instantiation)
int occ_fwd(int pc, Shape sh, int fast) {
    int n;
    switch (pc) {
        case 1: n = occ_fwd_p1(sh.DP, sh.F, fast); break;
        case 2: n = occ_fwd_p2(sh.DP, sh.F, fast); break;
        case 3: n = occ_fwd_p3(sh.DP, sh.F, fast); break;
        case 4: n = occ_fwd_p4(sh.DP, sh.F, fast); break;
        default: n = occ_fwd_p0(sh.DP, sh.F, fast); break;
    }
    return n < 1 ? 1 : n;
}
int occ_bwd(int pc, Shape sh) {
    int n;
    switch (pc) {
        case 1: n = occ_bwd_p1(sh.DP, sh.F); break;
        case 2: n = occ_bwd_p2(sh.DP, sh.F); break;
        case 3: n = occ_bwd_p3(sh.DP, sh.F); break;
        case 4: n = occ_bwd_p4(sh.DP, sh.F); break;
        default: n = occ_bwd_p0(sh.DP, sh.F); break;
    }
    return n < 1 ? 1 : n;
}
// The split plan only depends on (rows, MS, DP, #SMs) -- NOT on the exponent -- so that the workspace queries
// (which do not know p) and the launches agree: the most common occupancy of the family is used (p = 2).
// (the experimental fast forward -- CLICA_LPNCE_FAST, read here so that the workspace query and the launch agree --
// owns more rows per thread at small d and has its own occupancy)
int fwd_fast_enabled() { return env_flag("CLICA_LPNCE_FAST", 0) != 0; }
SplitPlan plan_fwd(int B, int M, Shape sh, int sms) {
    const int fast = fwd_fast_enabled();
    const int R = fwd_rows_per_thread(sh.DP, fast != 0);
    return plan_splits(B, rows_per_cta(R, sh.F), M, tile_rows(sh.F), sms, occ_fwd(2, sh, fast));
}
SplitPlan plan_bwd(int rows, int MS, Shape sh, int sms) {
    const int R = bwd_rows_per_thread(sh.DP);
    return plan_splits(rows, rows_per_cta(R, sh.F), MS, tile_rows(sh.F), sms, occ_bwd(2, sh));
}

int dispatch_fwd(int pc, Shape sh, const FwdParams& q, dim3 g, cudaStream_t s) {
    switch (pc) {
        case 1: return launch_fwd_p1(sh.DP, sh.F, q, g, s);
        case 2: return launch_fwd_p2(sh.DP, sh.F, q, g, s);
        case 3: return launch_fwd_p3(sh.DP, sh.F, q, g, s);
        case 4: return launch_fwd_p4(sh.DP, sh.F, q, g, s);
        default: return launch_fwd_p0(sh.DP, sh.F, q, g, s);
    }
}
int dispatch_bwd(int pc, Shape sh, const BwdParams& q, dim3 g, cudaStream_t s) {
    switch (pc) {
        case 1: return launch_bwd_p1(sh.DP, sh.F, q, g, s);
        case 2: return launch_bwd_p2(sh.DP, sh.F, q, g, s);
        case 3: return launch_bwd_p3(sh.DP, sh.F, q, g, s);
        case 4: return launch_bwd_p4(sh.DP, sh.F, q, g, s);
        default: return launch_bwd_p0(sh.DP, sh.F, q, g, s);
    }
}

int check_common(int B, int M, int d, float p, float tau, int use_pow, Shape* sh, DeviceInfo* di) {
    CLICA_REQUIRE(B >= 1 && M >= 1 && d >= 1, CLICA_E_BADARG, "lpnce: need B, M, d >= 1 (got %d, %d, %d)", B, M, d);
    CLICA_REQUIRE(tau > 0.f, CLICA_E_BADARG, "lpnce: tau must be > 0 (got %g)", (double)tau);
    CLICA_REQUIRE(p >= 1.f, CLICA_E_UNSUPPORTED,
                  "lpnce: p = %g < 1 (losses.py:433-442 branch) is not implemented by the CUDA path", (double)p);
    CLICA_REQUIRE(use_pow == 1, CLICA_E_UNSUPPORTED, "lpnce: pow=False is not implemented by the CUDA path");
    *sh = pick_shape(d);
    CLICA_REQUIRE(sh->DP > 0, CLICA_E_UNSUPPORTED, "lpnce: feature width d = %d > 320 is not implemented", d);
    int rc = get_device_info(di);
    if (rc) return rc;
    return 0;
}

struct FwdWs { int* counter; double* block_sums; float* part_m; float* part_s; size_t bytes; };
FwdWs carve_fwd(void* ws, int B, int nsplit) {
    FwdWs w;
    char* p = (char*)ws;
    size_t off = 0;
    w.counter = (int*)(p + off); off += 16;
    w.block_sums = (double*)(p + off); off += align_up(3ull * ceil_div(B, 256) * sizeof(double), 16);
    w.part_m = (float*)(p + off); off += align_up((size_t)nsplit * B * sizeof(float), 16);
    w.part_s = (float*)(p + off); off += align_up((size_t)nsplit * B * sizeof(float), 16);
    w.bytes = off;
    return w;
}

struct BwdWs { float* E; float* CP; float* partA; float* partB; size_t bytes; };
// nL = entries of E (B for the plain backward, M for the sharded one)
BwdWs carve_bwd(void* ws, int nL, int B, int rowsA, int nsA, int rowsB, int nsB, int TW) {
    BwdWs w;
    char* p = (char*)ws;
    size_t off = 0;
    w.E = (float*)(p + off); off += align_up((size_t)nL * sizeof(float), 16);
    w.CP = (float*)(p + off); off += align_up((size_t)B * sizeof(float), 16);
    w.partA = (float*)(p + off); off += align_up((size_t)nsA * rowsA * TW * sizeof(float), 16);
    w.partB = (float*)(p + off); off += align_up((size_t)nsB * rowsB * TW * sizeof(float), 16);
    w.bytes = off;
    return w;
}

inline int is_flat16(const float* S, int ldS, int d, Shape sh) {
    return (sh.F == 1) && (d == 2 * sh.DP) && (ldS == d) && (((uintptr_t)S & 15u) == 0);
}

}  // namespace
}  // namespace clica

using namespace clica;

extern "C" size_t clica_lpnce_workspace_bytes(int B, int M, int d) {
    DeviceInfo di;
    Shape sh = pick_shape(d);
    if (B < 1 || M < 1 || sh.DP < 0 || get_device_info(&di)) return 0;
    SplitPlan pl = plan_fwd(B, M, sh, di.sm_count);
    return carve_fwd(nullptr, B, pl.nsplit).bytes;
}

extern "C" int clica_lpnce_fwd(const float* z1, int ld1, const float* z2, int ld2, const float* z3, int ld3,
                               int B, int M, int d, float p, float tau, float alpha, int include_pos,
                               int use_pow, float* loss_i, float* lse, float* pos, float* rowstat,
                               float* scalars3, void* ws, size_t ws_bytes, void* stream) {
    Shape sh; DeviceInfo di;
    int rc = check_common(B, M, d, p, tau, use_pow, &sh, &di);
    if (rc) return rc;
    CLICA_REQUIRE(z1 && z2 && z3 && loss_i && lse && pos && rowstat && scalars3 && ws, CLICA_E_BADARG, "lpnce_fwd: null pointer");
    CLICA_REQUIRE(((uintptr_t)rowstat & 7u) == 0, CLICA_E_ALIGN, "lpnce_fwd: rowstat must be 8-byte aligned");
    CLICA_REQUIRE(ld1 >= d && ld2 >= d && ld3 >= d, CLICA_E_BADARG, "lpnce_fwd: leading dimension < d");
    CLICA_REQUIRE(((uintptr_t)ws & 15u) == 0, CLICA_E_ALIGN, "lpnce_fwd: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    SplitPlan pl = plan_fwd(B, M, sh, di.sm_count);
    FwdWs w = carve_fwd(ws, B, pl.nsplit);
    CLICA_REQUIRE(ws_bytes >= w.bytes, CLICA_E_WORKSPACE, "lpnce_fwd: workspace %zu < %zu bytes", ws_bytes, w.bytes);

    FwdParams q;
    q.O = z1; q.ldO = ld1; q.BO = B; q.S = z3; q.ldS = ld3; q.MS = M; q.d = d;
    q.coef = kLog2e / tau; q.pg = p;
    q.tiles_per_split = pl.tiles_per_split; q.flat16 = is_flat16(z3, ld3, d, sh);
    q.part_m = w.part_m; q.part_s = w.part_s; q.part_stride = B; q.counter = w.counter;
    q.fast = fwd_fast_enabled();                           // EXPERIMENTAL, off by default
    { LaunchScope ls(st, kFamLossFwd); rc = dispatch_fwd(p_code(p), sh, q, dim3(pl.row_tiles, pl.nsplit, 1), st); }
    if (rc) return rc;

    FinParams f;
    f.part_m = w.part_m; f.part_s = w.part_s; f.part_stride = B; f.nsplit = pl.nsplit;
    f.z1 = z1; f.ld1 = ld1; f.z2 = z2; f.ld2 = ld2; f.B = B; f.M = M; f.d = d; f.p = p; f.tau = tau;
    f.alpha = alpha; f.include_pos = include_pos;
    f.loss_i = loss_i; f.lse = lse; f.pos = pos; f.rowstat = (float2*)rowstat; f.scalars = scalars3;
    f.block_sums = w.block_sums; f.counter = w.counter;
    f.z3 = z3; f.ld3 = ld3; f.fast = q.fast;
    { LaunchScope ls(st, kFamLossAux); lpnce_finalize_kernel<<<ceil_div(B, 256), 256, 0, st>>>(f); }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" size_t clica_lpnce_bwd_workspace_bytes(int B, int M, int d) {
    DeviceInfo di;
    Shape sh = pick_shape(d);
    if (B < 1 || M < 1 || sh.DP < 0 || get_device_info(&di)) return 0;
    SplitPlan a = plan_bwd(B, M, sh, di.sm_count), b = plan_bwd(M, B, sh, di.sm_count);
    return carve_bwd(nullptr, B, B, B, a.nsplit, M, b.nsplit, 2 * sh.DP * sh.F).bytes;
}

extern "C" int clica_lpnce_bwd(const float* z1, int ld1, const float* z2, int ld2, const float* z3, int ld3,
                               int B, int M, int d, float p, float tau, float alpha, int include_pos,
                               int use_pow, const float* rowstat, const float* pos, const float* g_mean,
                               const float* g_loss_i, float* g_z1, int ldg1, float* g_z2, int ldg2,
                               float* g_z3, int ldg3, void* ws, size_t ws_bytes, void* stream) {
    Shape sh; DeviceInfo di;
    int rc = check_common(B, M, d, p, tau, use_pow, &sh, &di);
    if (rc) return rc;
    CLICA_REQUIRE(z1 && z2 && z3 && rowstat && pos && ws, CLICA_E_BADARG, "lpnce_bwd: null pointer");
    CLICA_REQUIRE(((uintptr_t)rowstat & 7u) == 0, CLICA_E_ALIGN, "lpnce_bwd: rowstat must be 8-byte aligned");
    CLICA_REQUIRE(ld1 >= d && ld2 >= d && ld3 >= d, CLICA_E_BADARG, "lpnce_bwd: leading dimension < d");
    CLICA_REQUIRE((!g_z1 || ldg1 >= d) && (!g_z2 || ldg2 >= d) && (!g_z3 || ldg3 >= d), CLICA_E_BADARG,
                  "lpnce_bwd: gradient leading dimension < d");
    CLICA_REQUIRE(((uintptr_t)ws & 15u) == 0, CLICA_E_ALIGN, "lpnce_bwd: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int TW = 2 * sh.DP * sh.F;
    SplitPlan pa = plan_bwd(B, M, sh, di.sm_count), pb = plan_bwd(M, B, sh, di.sm_count);
    BwdWs w = carve_bwd(ws, B, B, B, pa.nsplit, M, pb.nsplit, TW);
    CLICA_REQUIRE(ws_bytes >= w.bytes, CLICA_E_WORKSPACE, "lpnce_bwd: workspace %zu < %zu bytes", ws_bytes, w.bytes);

    PrepParams pp;
    pp.rowstat = (const float2*)rowstat; pp.pos = pos; pp.g_mean = g_mean; pp.g_loss_i = g_loss_i; pp.n = B;
    pp.inv_count = 1.f / (float)B;
    pp.tau = tau; pp.alpha = alpha; pp.include_pos = include_pos;
    pp.E = w.E; pp.CP = w.CP; pp.default_g = 0.f;
    { LaunchScope ls(st, kFamLossAux); lpnce_prep_kernel<<<ceil_div(B, 256), 256, 0, st>>>(pp); }
    CLICA_CUDA_OK(cudaGetLastError());

    const bool needA = (g_z1 != nullptr), needB = (g_z3 != nullptr);
    if (needA || needB) {
        BwdParams q;
        q.d = d; q.coef = kLog2e / tau; q.pg = p; q.nroles = 0;
        int gx = 0, gy = 0;
        if (needA) {   // anchors own, negatives stream
            BwdRole& r = q.role[q.nroles++];
            r.O = z1; r.ldO = ld1; r.BO = B; r.S = z3; r.ldS = ld3; r.MS = M;
            r.LO = (const float2*)rowstat; r.LS = nullptr; r.ES = nullptr; r.EO = nullptr;
            r.tiles_per_split = pa.tiles_per_split; r.nsplit = pa.nsplit; r.flat16 = is_flat16(z3, ld3, d, sh);
            r.part = w.partA; r.part_rows = B; r.row_tiles = pa.row_tiles;
            gx = max(gx, pa.row_tiles); gy = max(gy, pa.nsplit);
        }
        if (needB) {   // negatives own, anchors (with their lse and coefficient) stream
            BwdRole& r = q.role[q.nroles++];
            r.O = z3; r.ldO = ld3; r.BO = M; r.S = z1; r.ldS = ld1; r.MS = B;
            r.LO = nullptr; r.LS = (const float2*)rowstat; r.ES = w.E; r.EO = nullptr;
            r.tiles_per_split = pb.tiles_per_split; r.nsplit = pb.nsplit; r.flat16 = is_flat16(z1, ld1, d, sh);
            r.part = w.partB; r.part_rows = M; r.row_tiles = pb.row_tiles;
            gx = max(gx, pb.row_tiles); gy = max(gy, pb.nsplit);
        }
        { LaunchScope ls(st, kFamLossBwd); rc = dispatch_bwd(p_code(p), sh, q, dim3(gx, gy, q.nroles), st); }
        if (rc) return rc;
    }
    if (g_z1 || g_z2) {
        ReduceParams r;
        r.partA = needA ? w.partA : nullptr; r.nsA = pa.nsplit; r.rowsA = B;
        r.partB = nullptr; r.nsB = 0; r.rowsB = 0;
        r.E = w.E; r.CP = w.CP; r.z1 = z1; r.ld1 = ld1; r.z2 = z2; r.ld2 = ld2;
        r.g_out = g_z1; r.ldg = ldg1; r.g_z2 = g_z2; r.ldg2 = ldg2;
        r.rows = B; r.d = d; r.TW = TW; r.p = p;
        const long long n = (long long)B * d;
        { LaunchScope ls(st, kFamLossAux); lpnce_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(r); }
        CLICA_CUDA_OK(cudaGetLastError());
    }
    if (g_z3) {
        ReduceParams r;
        r.partA = nullptr; r.nsA = 0; r.rowsA = 0;
        r.partB = w.partB; r.nsB = pb.nsplit; r.rowsB = M;
        r.E = nullptr; r.CP = nullptr; r.z1 = nullptr; r.ld1 = 0; r.z2 = nullptr; r.ld2 = 0;
        r.g_out = g_z3; r.ldg = ldg3; r.g_z2 = nullptr; r.ldg2 = 0;
        r.rows = M; r.d = d; r.TW = TW; r.p = p;
        const long long n = (long long)M * d;
        { LaunchScope ls(st, kFamLossAux); lpnce_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(r); }
        CLICA_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

extern "C" size_t clica_lpnce_bwd_sharded_workspace_bytes(int B, int M, int d) {
    DeviceInfo di;
    Shape sh = pick_shape(d);
    if (B < 1 || M < 1 || sh.DP < 0 || get_device_info(&di)) return 0;
    SplitPlan a = plan_bwd(B, M, sh, di.sm_count);
    return carve_bwd(nullptr, M, B, B, a.nsplit, B, a.nsplit, 2 * sh.DP * sh.F).bytes;
}

extern "C" int clica_lpnce_bwd_sharded(const float* z1_local, int ld1, const float* z2_local, int ld2,
                                       const float* z_all, int ld3, const float* rowstat_all,
                                       const float* pos_local, int B, int M, int d, int row0, float p,
                                       float tau, float alpha, int include_pos, const float* g_scale,
                                       float* g_z1, int ldg1, float* g_z2, int ldg2,
                                       void* ws, size_t ws_bytes, void* stream) {
    Shape sh; DeviceInfo di;
    int rc = check_common(B, M, d, p, tau, 1, &sh, &di);
    if (rc) return rc;
    CLICA_REQUIRE(z1_local && z2_local && z_all && rowstat_all && pos_local && g_z1 && ws, CLICA_E_BADARG,
                  "lpnce_bwd_sharded: null pointer");
    CLICA_REQUIRE(((uintptr_t)rowstat_all & 7u) == 0, CLICA_E_ALIGN, "lpnce_bwd_sharded: rowstat must be 8-byte aligned");
    CLICA_REQUIRE(row0 >= 0 && row0 + B <= M, CLICA_E_BADARG, "lpnce_bwd_sharded: rows [%d, %d) outside [0, %d)", row0, row0 + B, M);
    CLICA_REQUIRE(ld1 >= d && ld2 >= d && ld3 >= d && ldg1 >= d && (!g_z2 || ldg2 >= d), CLICA_E_BADARG,
                  "lpnce_bwd_sharded: leading dimension < d");
    CLICA_REQUIRE(((uintptr_t)ws & 15u) == 0, CLICA_E_ALIGN, "lpnce_bwd_sharded: workspace must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int TW = 2 * sh.DP * sh.F;
    SplitPlan pa = plan_bwd(B, M, sh, di.sm_count);
    BwdWs w = carve_bwd(ws, M, B, B, pa.nsplit, B, pa.nsplit, TW);
    CLICA_REQUIRE(ws_bytes >= w.bytes, CLICA_E_WORKSPACE, "lpnce_bwd_sharded: workspace %zu < %zu bytes", ws_bytes, w.bytes);

    // coefficients of every global anchor (they all stream through the column-role pass) ...
    PrepParams pp;
    const float2* stat_all = (const float2*)rowstat_all;
    pp.rowstat = stat_all; pp.pos = nullptr; pp.g_mean = g_scale; pp.g_loss_i = nullptr; pp.n = M;
    pp.inv_count = 1.f / (float)M;
    pp.tau = tau; pp.alpha = alpha; pp.include_pos = include_pos;
    pp.E = w.E; pp.CP = nullptr; pp.default_g = 1.f;
    { LaunchScope ls(st, kFamLossAux); lpnce_prep_kernel<<<ceil_div(M, 256), 256, 0, st>>>(pp); }
    CLICA_CUDA_OK(cudaGetLastError());
    // ... and the positive-pair coefficient of the local rows
    pp.rowstat = stat_all + row0; pp.pos = pos_local; pp.n = B; pp.E = nullptr; pp.CP = w.CP;
    { LaunchScope ls(st, kFamLossAux); lpnce_prep_kernel<<<ceil_div(B, 256), 256, 0, st>>>(pp); }
    CLICA_CUDA_OK(cudaGetLastError());

    // ONE merged pass: the local rows are owners, every global row streams by once; the pair (i, j) contributes
    // [E_i w(i->j) + E_j w(j->i)] G'(z_j - z_i) -- anchor role and column role share the (symmetric) distance.
    BwdParams q;
    q.d = d; q.coef = kLog2e / tau; q.pg = p; q.nroles = 1;
    {
        BwdRole& r = q.role[0];
        r.O = z1_local; r.ldO = ld1; r.BO = B; r.S = z_all; r.ldS = ld3; r.MS = M;
        r.tiles_per_split = pa.tiles_per_split; r.nsplit = pa.nsplit; r.flat16 = is_flat16(z_all, ld3, d, sh);
        r.part_rows = B; r.row_tiles = pa.row_tiles;
        r.LO = stat_all + row0; r.EO = w.E + row0; r.LS = stat_all; r.ES = w.E; r.part = w.partB;
    }
    { LaunchScope ls(st, kFamLossBwd); rc = dispatch_bwd(p_code(p), sh, q, dim3(pa.row_tiles, pa.nsplit, 1), st); }
    if (rc) return rc;

    ReduceParams r;
    r.partA = nullptr; r.nsA = 0; r.rowsA = 0;
    r.partB = w.partB; r.nsB = pa.nsplit; r.rowsB = B;
    r.E = nullptr; r.CP = w.CP; r.z1 = z1_local; r.ld1 = ld1; r.z2 = z2_local; r.ld2 = ld2;
    r.g_out = g_z1; r.ldg = ldg1; r.g_z2 = g_z2; r.ldg2 = ldg2;
    r.rows = B; r.d = d; r.TW = TW; r.p = p;
    const long long n = (long long)B * d;
    { LaunchScope ls(st, kFamLossAux); lpnce_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(r); }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}
