// lpnce_api.cu -- C-ABI entry points of the fused Lp-InfoNCE loss (see include/clica.h): host-side planning of
// the pair-walking kernels of lpnce_kernels.cuh.  The training step's calls (no per-item upstream gradient) are ONE
// launch forward and ONE launch backward; per-item upstream gradients (dL/d loss_i given) take two small extra
// O(B*d) kernels around the backward: prep (per-anchor coefficients) and reduce (sum split partials, positive pair).
#include "lpnce_kernels.cuh"

namespace clica {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

// per-anchor coefficients of the backward:  E = 2 gl (1-alpha)/tau,  CP = 2 gl (alpha - (1-alpha) w+)/tau,
// gl_i = g_mean * inv_count + g_loss_i[i],  w+ = exp2((-pos*coef - m2) - ls) from the forward's row statistics
struct PrepParams {
    const float2* rowstat; const float* pos; const float* g_mean; const float* g_loss_i;
    int n; float inv_count; float tau; float alpha; int include_pos;
    float* E; float* CP;              // either may be null (CP null: pos / rowstat are not read)
    float default_g;                  // used when g_mean == nullptr
};
__global__ void lpnce_prep_kernel(const PrepParams q) {
    pdl_enter();
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
    int rows; int d; int TW; float p; int sim;
};
__global__ void lpnce_reduce_kernel(const ReduceParams q) {
    pdl_enter();
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
    if (!q.sim) g *= q.p;
    if (q.CP) {
        const float ov = q.z1[(size_t)row * q.ld1 + c], zv = q.z2[(size_t)row * q.ld2 + c];
        if (q.sim) {
            g -= q.CP[row] * zv;
            if (q.g_z2) q.g_z2[(size_t)row * q.ldg2 + c] = -q.CP[row] * ov;
        } else {
            const float gp = q.CP[row] * dabs_pow(ov - zv, q.p);
            g += gp;
            if (q.g_z2) q.g_z2[(size_t)row * q.ldg2 + c] = -gp;
        }
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
int p_code(float p) {
    if (p == 0.f) return kSim;        // dot-product similarity (SimCLRLoss; main_mlp.py selects it with --p 0)
    if (p == 1.f) return 1;
    if (p == 2.f) return 2;
    if (p == 3.f) return 3;
    if (p == 4.f) return 4;
    return 0;
}

struct SplitPlan { int row_tiles; int tiles_per_split; int nsplit; };
SplitPlan plan_splits(int rows, int rows_cta, int MS, int tile_cols, int sm_count, int ctas_per_sm) {
    // Column splits fill the SMs.  A CTA costs its streamed tiles plus a fixed part (owner rows, merge, partial write:
    // about a third of a tile) and the grid runs in waves of #SMs x resident CTAs: take the split count with the
    // shortest makespan, fewer splits unless more are >= 3 % better (e.g. 128 row tiles on 148 one-CTA SMs: 8 splits =
    // 7 waves of 8 tiles beat 1 split = 1 wave of 64 tiles with 20 SMs idle).
    SplitPlan pl;
    pl.row_tiles = ceil_div(rows, rows_cta);
    const int col_tiles = ceil_div(MS, tile_cols);
    const long long slots = (long long)sm_count * ctas_per_sm;
    const int max_ns = col_tiles < 64 ? col_tiles : 64;
    double best = 1e300;
    pl.tiles_per_split = col_tiles > 0 ? col_tiles : 1;
    pl.nsplit = 1;
    for (int ns = 1; ns <= max_ns; ++ns) {
        const int tps = ceil_div(col_tiles, ns);
        if (ceil_div(col_tiles, tps) != ns) continue;          // same division as a smaller split count
        const long long ctas = (long long)pl.row_tiles * ns;
        const long long waves = (ctas + slots - 1) / (slots > 0 ? slots : 1);
        const double cost = (double)waves * (tps + 0.35);
        if (cost < best * 0.97) { best = cost; pl.tiles_per_split = tps; pl.nsplit = ns; }
    }
    return pl;
}
// resident CTAs per SM of the kernel that will run (cudaOccupancy..., cached per instantiation)
int occ_fwd(int pc, Shape sh, int r4) {
    int n;
    switch (pc) {
        case 1: n = occ_fwd_p1(sh.DP, sh.F, r4); break;
        case 2: n = occ_fwd_p2(sh.DP, sh.F, r4); break;
        case 3: n = occ_fwd_p3(sh.DP, sh.F, r4); break;
        case 4: n = occ_fwd_p4(sh.DP, sh.F, r4); break;
        case 5: n = occ_fwd_p5(sh.DP, sh.F, r4); break;
        default: n = occ_fwd_p0(sh.DP, sh.F, r4); break;
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
        case 5: n = occ_bwd_p5(sh.DP, sh.F); break;
        default: n = occ_bwd_p0(sh.DP, sh.F); break;
    }
    return n < 1 ? 1 : n;
}
// The split plan depends on (rows, MS, DP, #SMs) and on the rows-per-thread variant of the forward -- NOT on the
// exponent (the most common occupancy of the family, p = 2, is used) -- so that the workspace queries, which do not
// know p, can bound what any launch will need.
// Forward variant with 4 owner rows per thread (dot form, p = 2, d <= 10): CLICA_LPNCE_R4 (default on)
int fwd_r4_enabled() { return env_flag("CLICA_LPNCE_R4", 1) != 0; }
// dot-form bound (see lpnce_kernels.cuh); CLICA_LPNCE_DOT=0 disables the form
float dot_limit() {
    if (env_flag("CLICA_LPNCE_DOT", 1) == 0) return 0.f;
    const int tenths = env_flag("CLICA_LPNCE_DOT_LIMIT_X10", 80);      // the bound in tenths (default 8.0)
    return tenths > 0 ? 0.1f * (float)tenths : 0.f;
}
// backward: measured on B200 (profiles/r2_loss_probe.md) the dot form is SLOWER there than subtract-then-square (the
// weights need both operands anyway and the kernel is latency-bound at 16 warps per SM), so it is opt-in
float dot_limit_bwd() { return env_flag("CLICA_LPNCE_DOT_BWD", 0) != 0 ? dot_limit() : 0.f; }
bool r4_applies(int pc, Shape sh) { return pc == 2 && sh.F == 1 && fwd_rows_per_thread(sh.DP, true) != fwd_rows_per_thread(sh.DP, false); }
SplitPlan plan_fwd(int B, int M, Shape sh, int sms, int r4) {
    const int R = fwd_rows_per_thread(sh.DP, r4 != 0);
    return plan_splits(B, rows_per_cta(R, sh.F), M, tile_rows(sh.F), sms, occ_fwd(2, sh, r4));
}
SplitPlan plan_bwd(int rows, int MS, Shape sh, int sms) {
    const int R = bwd_rows_per_thread(sh.DP);
    return plan_splits(rows, rows_per_cta(R, sh.F), MS, tile_rows(sh.F), sms, occ_bwd(2, sh));
}

int dispatch_fwd(int pc, Shape sh, const FwdParams& q, dim3 g, cudaStream_t s, int r4) {
    switch (pc) {
        case 1: return launch_fwd_p1(sh.DP, sh.F, q, g, s, r4);
        case 2: return launch_fwd_p2(sh.DP, sh.F, q, g, s, r4);
        case 3: return launch_fwd_p3(sh.DP, sh.F, q, g, s, r4);
        case 4: return launch_fwd_p4(sh.DP, sh.F, q, g, s, r4);
        case 5: return launch_fwd_p5(sh.DP, sh.F, q, g, s, r4);
        default: return launch_fwd_p0(sh.DP, sh.F, q, g, s, r4);
    }
}
int dispatch_bwd(int pc, Shape sh, const BwdParams& q, dim3 g, cudaStream_t s) {
    switch (pc) {
        case 1: return launch_bwd_p1(sh.DP, sh.F, q, g, s);
        case 2: return launch_bwd_p2(sh.DP, sh.F, q, g, s);
        case 3: return launch_bwd_p3(sh.DP, sh.F, q, g, s);
        case 4: return launch_bwd_p4(sh.DP, sh.F, q, g, s);
        case 5: return launch_bwd_p5(sh.DP, sh.F, q, g, s);
        default: return launch_bwd_p0(sh.DP, sh.F, q, g, s);
    }
}

int check_common(int B, int M, int d, float p, float tau, int use_pow, Shape* sh, DeviceInfo* di) {
    CLICA_REQUIRE(B >= 1 && M >= 1 && d >= 1, CLICA_E_BADARG, "lpnce: need B, M, d >= 1 (got %d, %d, %d)", B, M, d);
    CLICA_REQUIRE(tau > 0.f, CLICA_E_BADARG, "lpnce: tau must be > 0 (got %g)", (double)tau);
    CLICA_REQUIRE(p >= 1.f || p == 0.f, CLICA_E_UNSUPPORTED,
                  "lpnce: 0 < p = %g < 1 (losses.py:433-442 branch) is not implemented by the CUDA path", (double)p);
    CLICA_REQUIRE(use_pow == 1, CLICA_E_UNSUPPORTED, "lpnce: pow=False is not implemented by the CUDA path");
    *sh = pick_shape(d);
    CLICA_REQUIRE(sh->DP > 0, CLICA_E_UNSUPPORTED, "lpnce: feature width d = %d > 320 is not implemented", d);
    int rc = get_device_info(di);
    if (rc) return rc;
    return 0;
}

struct FwdWs { int* done_counter; int* tile_counter; double* block_sums; float* part_m; float* part_s; size_t bytes; };
FwdWs carve_fwd(void* ws, int B, int row_tiles, int nsplit) {
    FwdWs w;
    char* p = (char*)ws;
    size_t off = 0;
    w.done_counter = (int*)(p + off);
    w.tile_counter = (int*)(p + off) + 4;
    off += kCounterBytes;                                  // zero on entry, zero on exit
    w.block_sums = (double*)(p + off); off += align_up(3ull * row_tiles * sizeof(double), 16);
    w.part_m = (float*)(p + off); off += align_up((size_t)nsplit * B * sizeof(float), 16);
    w.part_s = (float*)(p + off); off += align_up((size_t)nsplit * B * sizeof(float), 16);
    w.bytes = off;
    return w;
}

struct BwdWs { int* counterA; int* counterB; float* E; float* CP; float* partA; float* partB; size_t bytes; };
// nL = entries of E (B for the plain backward, M for the sharded one)
BwdWs carve_bwd(void* ws, int nL, int B, int rowsA, int nsA, int rowsB, int nsB, int TW) {
    BwdWs w;
    char* p = (char*)ws;
    size_t off = 0;
    w.counterA = (int*)(p + off);
    w.counterB = (int*)(p + off + kCounterBytes / 2);
    off += kCounterBytes;                                  // zero on entry, zero on exit
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

void init_role(BwdRole& r) {
    r.O = nullptr; r.ldO = 0; r.BO = 0; r.S = nullptr; r.ldS = 0; r.MS = 0;
    r.LO = nullptr; r.LS = nullptr; r.ES = nullptr; r.EO = nullptr; r.merged = 0; r.stream_weighted = 0;
    r.tiles_per_split = 1; r.nsplit = 1; r.flat16 = 0; r.part = nullptr; r.part_rows = 0; r.row_tiles = 0;
    r.tile_counter = nullptr; r.g_out = nullptr; r.ldg = 0; r.scale_by_E = 0;
    r.Z2 = nullptr; r.ld2 = 0; r.LP = nullptr; r.POS = nullptr; r.g_z2 = nullptr; r.ldg2 = 0; r.with_pos = 0;
}

}  // namespace
}  // namespace clica

using namespace clica;

extern "C" size_t clica_lpnce_workspace_bytes(int B, int M, int d) {
    DeviceInfo di;
    Shape sh = pick_shape(d);
    if (B < 1 || M < 1 || sh.DP < 0 || get_device_info(&di)) return 0;
    size_t bytes = 0;
    for (int r4 = 0; r4 < 2; ++r4) {          // the launch picks its variant from p, which the query does not know
        SplitPlan pl = plan_fwd(B, M, sh, di.sm_count, r4);
        const size_t b = carve_fwd(nullptr, B, pl.row_tiles, pl.nsplit).bytes;
        if (b > bytes) bytes = b;
    }
    return bytes;
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
    const int pc = p_code(p);
    const int r4 = (fwd_r4_enabled() && r4_applies(pc, sh)) ? 1 : 0;
    SplitPlan pl = plan_fwd(B, M, sh, di.sm_count, r4);
    CLICA_REQUIRE(pl.row_tiles <= kMaxRowTilesFwd, CLICA_E_UNSUPPORTED, "lpnce_fwd: B = %d needs %d row tiles (> %d)", B, pl.row_tiles, kMaxRowTilesFwd);
    FwdWs w = carve_fwd(ws, B, pl.row_tiles, pl.nsplit);
    CLICA_REQUIRE(ws_bytes >= w.bytes, CLICA_E_WORKSPACE, "lpnce_fwd: workspace %zu < %zu bytes", ws_bytes, w.bytes);

    FwdParams q;
    q.O = z1; q.ldO = ld1; q.BO = B; q.S = z3; q.ldS = ld3; q.MS = M; q.Z2 = z2; q.ld2 = ld2; q.d = d;
    q.coef = kLog2e / tau; q.pg = p; q.tau = tau; q.alpha = alpha; q.include_pos = include_pos;
    q.tiles_per_split = pl.tiles_per_split; q.nsplit = pl.nsplit; q.flat16 = is_flat16(z3, ld3, d, sh);
    q.dot_limit = dot_capable(pc, sh.F) ? dot_limit() : 0.f;
    q.part_m = w.part_m; q.part_s = w.part_s; q.part_stride = B;
    q.done_counter = w.done_counter; q.tile_counter = w.tile_counter; q.block_sums = w.block_sums;
    q.loss_i = loss_i; q.lse = lse; q.pos = pos; q.rowstat = (float2*)rowstat; q.scalars = scalars3;
    { LaunchScope ls(st, kFamLossFwd); rc = dispatch_fwd(pc, sh, q, dim3(pl.row_tiles, pl.nsplit, 1), st, r4); }
    return rc;
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
    const int pc = p_code(p);
    SplitPlan pa = plan_bwd(B, M, sh, di.sm_count), pb = plan_bwd(M, B, sh, di.sm_count);
    CLICA_REQUIRE(pa.row_tiles <= kMaxRowTilesBwd && pb.row_tiles <= kMaxRowTilesBwd, CLICA_E_UNSUPPORTED,
                  "lpnce_bwd: %d / %d row tiles (> %d)", pa.row_tiles, pb.row_tiles, kMaxRowTilesBwd);
    BwdWs w = carve_bwd(ws, B, B, B, pa.nsplit, M, pb.nsplit, TW);
    CLICA_REQUIRE(ws_bytes >= w.bytes, CLICA_E_WORKSPACE, "lpnce_bwd: workspace %zu < %zu bytes", ws_bytes, w.bytes);
    const bool fused = (g_loss_i == nullptr);     // one upstream scalar: everything happens inside the pair kernel

    if (!fused) {
        PrepParams pp;
        pp.rowstat = (const float2*)rowstat; pp.pos = pos; pp.g_mean = g_mean; pp.g_loss_i = g_loss_i; pp.n = B;
        pp.inv_count = 1.f / (float)B;
        pp.tau = tau; pp.alpha = alpha; pp.include_pos = include_pos;
        pp.E = w.E; pp.CP = w.CP; pp.default_g = 0.f;
        { LaunchScope ls(st, kFamLossAux); launch_k(lpnce_prep_kernel, ceil_div(B, 256), 256, 0, st, pp); }
        CLICA_CUDA_OK(cudaGetLastError());
    }

    const bool needA = (g_z1 != nullptr) || (fused && g_z2 != nullptr), needB = (g_z3 != nullptr);
    if (needA || needB) {
        BwdParams q;
        q.d = d; q.coef = kLog2e / tau; q.pg = p; q.nroles = 0;
        q.tau = tau; q.alpha = alpha; q.include_pos = include_pos;
        q.fused = fused ? 1 : 0; q.g_mean = g_mean; q.default_g = 0.f; q.inv_count = 1.f / (float)B;
        q.dot_limit = dot_capable(pc, sh.F) ? dot_limit_bwd() : 0.f;
        init_role(q.role[0]); init_role(q.role[1]);
        int gx = 0, gy = 0;
        if (needA) {   // anchors own, negatives stream
            BwdRole& r = q.role[q.nroles++];
            r.O = z1; r.ldO = ld1; r.BO = B; r.S = z3; r.ldS = ld3; r.MS = M;
            r.LO = (const float2*)rowstat;
            r.tiles_per_split = pa.tiles_per_split; r.nsplit = pa.nsplit; r.flat16 = is_flat16(z3, ld3, d, sh);
            r.part = w.partA; r.part_rows = B; r.row_tiles = pa.row_tiles;
            r.tile_counter = w.counterA; r.g_out = g_z1; r.ldg = ldg1; r.scale_by_E = 1;
            r.Z2 = z2; r.ld2 = ld2; r.LP = (const float2*)rowstat; r.POS = pos; r.g_z2 = g_z2; r.ldg2 = ldg2; r.with_pos = 1;
            gx = max(gx, pa.row_tiles); gy = max(gy, pa.nsplit);
        }
        if (needB) {   // negatives own, anchors (with their lse and coefficient) stream
            BwdRole& r = q.role[q.nroles++];
            r.O = z3; r.ldO = ld3; r.BO = M; r.S = z1; r.ldS = ld1; r.MS = B;
            r.LS = (const float2*)rowstat; r.ES = fused ? nullptr : w.E; r.stream_weighted = 1;
            r.tiles_per_split = pb.tiles_per_split; r.nsplit = pb.nsplit; r.flat16 = is_flat16(z1, ld1, d, sh);
            r.part = w.partB; r.part_rows = M; r.row_tiles = pb.row_tiles;
            r.tile_counter = w.counterB; r.g_out = g_z3; r.ldg = ldg3; r.scale_by_E = 0;
            gx = max(gx, pb.row_tiles); gy = max(gy, pb.nsplit);
        }
        { LaunchScope ls(st, kFamLossBwd); rc = dispatch_bwd(pc, sh, q, dim3(gx, gy, q.nroles), st); }
        if (rc) return rc;
    }
    if (fused) return 0;
    if (g_z1 || g_z2) {
        ReduceParams r;
        r.partA = (g_z1 != nullptr) ? w.partA : nullptr; r.nsA = pa.nsplit; r.rowsA = B;
        r.partB = nullptr; r.nsB = 0; r.rowsB = 0;
        r.E = w.E; r.CP = w.CP; r.z1 = z1; r.ld1 = ld1; r.z2 = z2; r.ld2 = ld2;
        r.g_out = g_z1; r.ldg = ldg1; r.g_z2 = g_z2; r.ldg2 = ldg2;
        r.rows = B; r.d = d; r.TW = TW; r.p = p; r.sim = (pc == kSim) ? 1 : 0;
        const long long n = (long long)B * d;
        { LaunchScope ls(st, kFamLossAux); launch_k(lpnce_reduce_kernel, (unsigned)((n + 255) / 256), 256, 0, st, r); }
        CLICA_CUDA_OK(cudaGetLastError());
    }
    if (g_z3) {
        ReduceParams r;
        r.partA = nullptr; r.nsA = 0; r.rowsA = 0;
        r.partB = w.partB; r.nsB = pb.nsplit; r.rowsB = M;
        r.E = nullptr; r.CP = nullptr; r.z1 = nullptr; r.ld1 = 0; r.z2 = nullptr; r.ld2 = 0;
        r.g_out = g_z3; r.ldg = ldg3; r.g_z2 = nullptr; r.ldg2 = 0;
        r.rows = M; r.d = d; r.TW = TW; r.p = p; r.sim = (pc == kSim) ? 1 : 0;
        const long long n = (long long)M * d;
        { LaunchScope ls(st, kFamLossAux); launch_k(lpnce_reduce_kernel, (unsigned)((n + 255) / 256), 256, 0, st, r); }
        CLICA_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

extern "C" size_t clica_lpnce_bwd_sharded_workspace_bytes(int B, int M, int d) {
    DeviceInfo di;
    Shape sh = pick_shape(d);
    if (B < 1 || M < 1 || sh.DP < 0 || get_device_info(&di)) return 0;
    SplitPlan a = plan_bwd(B, M, sh, di.sm_count);
    return carve_bwd(nullptr, 0, 0, 0, 0, B, a.nsplit, 2 * sh.DP * sh.F).bytes;
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
    const int pc = p_code(p);
    SplitPlan pa = plan_bwd(B, M, sh, di.sm_count);
    CLICA_REQUIRE(pa.row_tiles <= kMaxRowTilesBwd, CLICA_E_UNSUPPORTED, "lpnce_bwd_sharded: %d row tiles (> %d)", pa.row_tiles, kMaxRowTilesBwd);
    BwdWs w = carve_bwd(ws, 0, 0, 0, 0, B, pa.nsplit, TW);
    CLICA_REQUIRE(ws_bytes >= w.bytes, CLICA_E_WORKSPACE, "lpnce_bwd_sharded: workspace %zu < %zu bytes", ws_bytes, w.bytes);
    const float2* stat_all = (const float2*)rowstat_all;

    // ONE merged pass, ONE launch: the local rows are owners, every global row streams by once; the pair (i, j)
    // contributes E [w(i->j) + w(j->i)] G'(z_j - z_i) -- anchor role and column role share the (symmetric) distance;
    // the last split CTA of each row tile sums the splits and adds the positive-pair term
    BwdParams q;
    q.d = d; q.coef = kLog2e / tau; q.pg = p; q.nroles = 1;
    q.tau = tau; q.alpha = alpha; q.include_pos = include_pos;
    q.fused = 1; q.g_mean = g_scale; q.default_g = 1.f; q.inv_count = 1.f / (float)M;
    q.dot_limit = dot_capable(pc, sh.F) ? dot_limit_bwd() : 0.f;
    init_role(q.role[0]); init_role(q.role[1]);
    {
        BwdRole& r = q.role[0];
        r.O = z1_local; r.ldO = ld1; r.BO = B; r.S = z_all; r.ldS = ld3; r.MS = M;
        r.tiles_per_split = pa.tiles_per_split; r.nsplit = pa.nsplit; r.flat16 = is_flat16(z_all, ld3, d, sh);
        r.part_rows = B; r.row_tiles = pa.row_tiles; r.part = w.partB;
        r.LO = stat_all + row0; r.LS = stat_all; r.merged = 1; r.stream_weighted = 1;
        r.tile_counter = w.counterA; r.g_out = g_z1; r.ldg = ldg1; r.scale_by_E = 0;
        r.Z2 = z2_local; r.ld2 = ld2; r.LP = stat_all + row0; r.POS = pos_local; r.g_z2 = g_z2; r.ldg2 = ldg2; r.with_pos = 1;
    }
    { LaunchScope ls(st, kFamLossBwd); rc = dispatch_bwd(pc, sh, q, dim3(pa.row_tiles, pa.nsplit, 1), st); }
    return rc;
}
