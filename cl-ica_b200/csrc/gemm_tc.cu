// gemm_tc.cu -- hand-written tcgen05 GEMM for the encoder's Linear layers (sm_100a).
//
//   D[128 x 128 tile] = sum_r A(m, r) B(n, r)          fp32 operands consumed as TF32 by tcgen05.mma kind::tf32
//
// Replaces the cuBLAS SGEMMs behind nn.Linear fwd / AddmmBackward of /root/reference/encoders.py:39.
//
// Structure (one output tile per CTA, 192 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor.2d (128B-swizzled boxes) -> smem ring, mbarrier complete_tx
//   warp 1      MMA issuer    : one elected thread issues tcgen05.mma (M=128, N=128, K=8) from smem descriptors,
//                               accumulator in TMEM (128 columns x 128 lanes fp32); tcgen05.commit frees the stage
//   warps 2..5  epilogue      : tcgen05.ld (32 lanes x 32 columns per instruction) -> registers -> fused
//                               bias + LeakyReLU | activation mask | split-K atomic accumulate -> global
// Both operand majors are supported natively (instruction-descriptor a_major / b_major bits, MN-major
// 128B-swizzle shared-memory descriptors), so forward (x W^T), backward-data (dY W) and backward-weight
// (dY^T X) all read the row-major tensors as they lie in HBM: no transposes are ever materialised.
// 3xTF32: every stage holds A_hi, A_lo, B_hi, B_lo; three MMAs per k-step (hi*hi + lo*hi + hi*lo).
#include "gemm_tc.cuh"
#include "gemm_simt.cuh"

#include <cuda.h>   // CUtensorMap + enums only; the driver entry point is fetched through the runtime

#include <stdlib.h>

#include <mutex>
#include <unordered_map>

namespace clica {
namespace {

constexpr int BM = 128, BN = 128, BK = 32;             // BK fp32 = 128 bytes = one swizzle span
constexpr int kTileBytes = BM * BK * 4;                // 16 KiB per operand plane per stage
constexpr int kTcThreads = 192;
constexpr int kTmemCols = 128;
constexpr int kMaxStages = 8;

struct TcKernelParams {
    int Mo, No;
    int kb_total, kb_per_split;
    int a_mn, b_mn, nterms, stages;
    int epi;
    const float* bias; float slope;
    const float* aux; int ldaux;
    float* out; int ldo;
    float* out_hi; float* out_lo; int ldp;
    uint32_t mn_lbo, mn_sbo, mn_lt;   // MN-major descriptor fields (defaults: BK*128, 512, 1)
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, descriptor version 1 (sm_100)
//   K-major : SWIZZLE_128B (layout_type 2, 16-byte chunks XOR row%8): rows of 128 B (32 fp32 along K);
//             8-row groups 1024 B apart (SBO); LBO unused (1)
//   MN-major: tf32 operands that are contiguous along M/N must use SWIZZLE_128B_BASE32B (layout_type 1,
//             32-byte chunks XOR row%4; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 128 B
//             (32 fp32 along M/N) indexed by k; 4-k groups 512 B apart (SBO); consecutive 32-wide M/N
//             groups BK*128 B apart (LBO)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;   // version
    d |= (uint64_t)layout_type << 61;
    return d;
}
// instruction descriptor: D fp32, A/B tf32, M = 128, N = 128
__device__ __forceinline__ uint32_t make_idesc(int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// one operand tile of a stage: K-major = one {32 x 128} box; MN-major = four {32 x 32} boxes
__device__ __forceinline__ void load_operand(const CUtensorMap* tm, uint32_t dst, int mn_major, int mn0, int k0, uint64_t* bar) {
    if (!mn_major) {
        tma_load_2d(dst, tm, k0, mn0, bar);
    } else {
#pragma unroll
        for (int j = 0; j < BM / 32; ++j) tma_load_2d(dst + j * (BK * 128), tm, mn0 + 32 * j, k0, bar);
    }
}

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
               const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl,
               const TcKernelParams q) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B needs 1024-byte alignment
    const int nplanes = (q.nterms == 3) ? 2 : 1;
    const uint32_t stage_bytes = 2u * nplanes * kTileBytes;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kb0 = blockIdx.z * q.kb_per_split;
    const int kb1 = min(q.kb_total, kb0 + q.kb_per_split);
    const int nkb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < q.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, kTmemCols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int i = 0; i < nkb; ++i) {
                const int s = i % q.stages;
                const uint32_t ph = (uint32_t)(i / q.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
                const uint32_t sa = tiles + s * stage_bytes;
                const int k0 = (kb0 + i) * BK;
                load_operand(&tmAh, sa, q.a_mn, m0, k0, &full_bar[s]);
                if (nplanes == 2) load_operand(&tmAl, sa + kTileBytes, q.a_mn, m0, k0, &full_bar[s]);
                load_operand(&tmBh, sa + nplanes * kTileBytes, q.b_mn, n0, k0, &full_bar[s]);
                if (nplanes == 2) load_operand(&tmBl, sa + (nplanes + 1) * kTileBytes, q.b_mn, n0, k0, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            const uint32_t idesc = make_idesc(q.a_mn, q.b_mn);
            const uint32_t a_step = q.a_mn ? 1024u : 32u, b_step = q.b_mn ? 1024u : 32u;   // bytes per K = 8 step
            const uint32_t a_lbo = q.a_mn ? q.mn_lbo : 16u, b_lbo = q.b_mn ? q.mn_lbo : 16u;
            const uint32_t a_sbo = q.a_mn ? q.mn_sbo : 1024u, b_sbo = q.b_mn ? q.mn_sbo : 1024u;
            const uint32_t a_lt = q.a_mn ? q.mn_lt : 2u, b_lt = q.b_mn ? q.mn_lt : 2u;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % q.stages;
                const uint32_t ph = (uint32_t)(i / q.stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t sa = tiles + s * stage_bytes;
                const uint32_t a_hi = sa, a_lo = sa + kTileBytes;
                const uint32_t b_hi = sa + nplanes * kTileBytes, b_lo = sa + (nplanes + 1) * kTileBytes;
#pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t dah = make_smem_desc(a_hi + ks * a_step, a_lbo, a_sbo, a_lt);
                    const uint64_t dbh = make_smem_desc(b_hi + ks * b_step, b_lbo, b_sbo, b_lt);
                    if (nplanes == 2) {
                        const uint64_t dal = make_smem_desc(a_lo + ks * a_step, a_lbo, a_sbo, a_lt);
                        const uint64_t dbl = make_smem_desc(b_lo + ks * b_step, b_lbo, b_sbo, b_lt);
                        umma_tf32(tmem_base, dal, dbh, idesc, (i > 0 || ks > 0) ? 1u : 0u);   // small terms first
                        umma_tf32(tmem_base, dah, dbl, idesc, 1u);
                        umma_tf32(tmem_base, dah, dbh, idesc, 1u);
                    } else {
                        umma_tf32(tmem_base, dah, dbh, idesc, (i > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&empty_bar[s]);        // stage reusable once these MMAs have read it
            }
            umma_commit(&tmem_full_bar);           // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        // tcgen05.ld hands every thread 32 consecutive columns of ITS row; storing that directly would touch 32
        // different cache lines per instruction.  Each warp therefore transposes its 32 x 32 chunk through a
        // padded shared-memory tile (the pipeline stages are dead by now) and walks it row by row, so that every
        // global access of the epilogue (bias, mask, plain / hi / lo stores, split-K atomics) is one 128-byte line.
        const int wq = warp & 3;
        mbar_wait(&tmem_full_bar, 0u);
        tcgen05_fence_after();
        float* stg = reinterpret_cast<float*>(smem_raw + (tiles - smem_u32(smem_raw))) + wq * (32 * 33);
        const int row0 = m0 + wq * 32;
        const int nrows = min(32, q.Mo - row0);
#pragma unroll 1
        for (int chunk = 0; chunk < BN / 32; ++chunk) {
            float v[32];
            __syncwarp();   // previous chunk's readers are done with stg; tcgen05.ld is .sync.aligned
            tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(chunk * 32), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];      // bank (lane + j) % 32: conflict-free
            __syncwarp();
            const int col = n0 + chunk * 32 + lane;
            if (col >= q.No || nrows <= 0) continue;
            const float bias_v = (q.epi == kTcBiasAct && q.bias != nullptr) ? __ldg(q.bias + col) : 0.f;
            for (int r = 0; r < nrows; ++r) {
                float t = stg[r * 33 + lane];
                const size_t grow = (size_t)(row0 + r);
                if (q.epi == kTcAtomic) {
                    atomicAdd(q.out + grow * q.ldo + col, t);
                    continue;
                }
                if (q.epi == kTcBiasAct) {
                    t += bias_v;
                    t = t > 0.f ? t : t * q.slope;
                } else if (q.aux != nullptr) {
                    t *= (__ldg(q.aux + grow * q.ldaux + col) > 0.f) ? 1.f : q.slope;
                }
                if (q.out != nullptr) q.out[grow * q.ldo + col] = t;
                if (q.out_hi != nullptr) {
                    if (q.out_lo != nullptr) {
                        const float h = round_to_tf32(t);
                        q.out_hi[grow * q.ldp + col] = h;
                        q.out_lo[grow * q.ldp + col] = t - h;
                    } else {
                        q.out_hi[grow * q.ldp + col] = t;
                    }
                }
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// split a plain matrix into planes; columns [cols, ld_dst) of the planes are zero-filled
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, int ld_src, int rows, int cols,
                                                            float* __restrict__ hi, float* __restrict__ lo, int ld_dst) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * ld_dst) return;
    const int r = (int)(idx / ld_dst), c = (int)(idx - (long long)r * ld_dst);
    const float v = (c < cols) ? __ldg(src + (size_t)r * ld_src + c) : 0.f;
    if (lo != nullptr) {
        const float h = round_to_tf32(v);
        hi[idx] = h;
        lo[idx] = v - h;
    } else {
        hi[idx] = v;
    }
}

// ---- host: tensor maps ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

struct MapKey {
    const void* ptr; int rows, cols, ld, box_rows;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.cols;
        h = h * 1000003u ^ (size_t)k.ld; h = h * 1000003u ^ (size_t)k.box_rows;
        return h;
    }
};

// 2-D fp32 row-major tensor [rows][cols] with row pitch ld; boxes of 32 columns x box_rows rows, 128B swizzle,
// out-of-bounds elements read as zero (ragged M / N / K tails need no special casing in the kernel).
int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

int get_tensor_map(const float* ptr, int rows, int cols, int ld, int box_rows, bool mn_major, CUtensorMap* out) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const int mn_swz = env_int("CLICA_TC_MN_SWZ", (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);   // debug override
    MapKey key{ptr, rows, cols, ld, mn_major ? -(box_rows + 1000 * mn_swz) : box_rows};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return 0; }
    }
    EncodeTiledFn enc = get_encode_fn();
    CLICA_REQUIRE(enc != nullptr, CLICA_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
    CLICA_REQUIRE((((uintptr_t)ptr) & 15u) == 0 && (ld % 4) == 0, CLICA_E_ALIGN,
                  "TMA operand must be 16-byte aligned with a leading dimension that is a multiple of 4 floats");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? (CUtensorMapSwizzle)mn_swz : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CLICA_REQUIRE(r == CUDA_SUCCESS, CLICA_E_BADARG, "cuTensorMapEncodeTiled failed (%d) for [%d x %d] ld %d", (int)r, rows, cols, ld);
    {
        std::lock_guard<std::mutex> lk(mu);
        if (cache.size() > 8192) cache.clear();
        cache.emplace(key, tm);
    }
    *out = tm;
    return 0;
}

}  // namespace

bool tc_shape_ok(int M_out, int N_out, int K_red) { return M_out >= 32 && N_out >= 32 && K_red >= 32; }

int tc_split_planes(const float* src, int ld_src, int rows, int cols, float* hi, float* lo, int ld_dst, cudaStream_t st) {
    const long long n = (long long)rows * ld_dst;
    { LaunchScope ls(st, kFamMisc); split_planes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, ld_src, rows, cols, hi, lo, ld_dst); }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

int tc_gemm_launch(const TcGemm& g, int sm_count, cudaStream_t st) {
    CLICA_REQUIRE(g.Mo >= 1 && g.No >= 1 && g.Kr >= 1, CLICA_E_BADARG, "tc_gemm: empty problem");
    CLICA_REQUIRE(g.A.hi && g.B.hi, CLICA_E_BADARG, "tc_gemm: null operand");
    CLICA_REQUIRE((g.A.lo != nullptr) == (g.B.lo != nullptr), CLICA_E_BADARG, "tc_gemm: operands must both be split or both plain");
    const int nterms = g.A.lo ? 3 : 1;
    CUtensorMap tAh, tAl, tBh, tBl;
    int rc;
    // storage shape of each operand: K-major [MN rows][Kr cols] (box 128 rows); MN-major [Kr rows][MN cols] (box 32 rows)
    const int a_rows = g.a_mn_major ? g.Kr : g.Mo, a_cols = g.a_mn_major ? g.Mo : g.Kr, a_box = g.a_mn_major ? BK : BM;
    const int b_rows = g.b_mn_major ? g.Kr : g.No, b_cols = g.b_mn_major ? g.No : g.Kr, b_box = g.b_mn_major ? BK : BN;
    if ((rc = get_tensor_map(g.A.hi, a_rows, a_cols, g.A.ld, a_box, g.a_mn_major != 0, &tAh))) return rc;
    if ((rc = get_tensor_map(g.B.hi, b_rows, b_cols, g.B.ld, b_box, g.b_mn_major != 0, &tBh))) return rc;
    if (nterms == 3) {
        if ((rc = get_tensor_map(g.A.lo, a_rows, a_cols, g.A.ld, a_box, g.a_mn_major != 0, &tAl))) return rc;
        if ((rc = get_tensor_map(g.B.lo, b_rows, b_cols, g.B.ld, b_box, g.b_mn_major != 0, &tBl))) return rc;
    } else {
        tAl = tAh; tBl = tBh;
    }
    TcKernelParams q;
    q.Mo = g.Mo; q.No = g.No;
    q.kb_total = ceil_div(g.Kr, BK);
    q.a_mn = g.a_mn_major; q.b_mn = g.b_mn_major; q.nterms = nterms;
    q.stages = (nterms == 3) ? 3 : 6;
    q.epi = g.epi; q.bias = g.bias; q.slope = g.slope; q.aux = g.aux; q.ldaux = g.ldaux;
    q.out = g.out; q.ldo = g.ldo; q.out_hi = g.outp.hi; q.out_lo = g.outp.lo; q.ldp = g.outp.ld;
    q.mn_lbo = (uint32_t)env_int("CLICA_TC_MN_LBO", BK * 128);   // debug overrides of the MN-major descriptor
    q.mn_sbo = (uint32_t)env_int("CLICA_TC_MN_SBO", 512);
    q.mn_lt = (uint32_t)env_int("CLICA_TC_MN_LT", 1);
    CLICA_REQUIRE(g.epi != kTcAtomic || g.out != nullptr, CLICA_E_BADARG, "tc_gemm: atomic epilogue needs a plain output");
    const int tiles = ceil_div(g.Mo, BM) * ceil_div(g.No, BN);
    int splits = 1;
    if (g.epi == kTcAtomic && g.allow_split_k) {
        splits = (sm_count + tiles - 1) / tiles;
        const int max_splits = q.kb_total / 4 > 0 ? q.kb_total / 4 : 1;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
    }
    q.kb_per_split = ceil_div(q.kb_total, splits);
    splits = ceil_div(q.kb_total, q.kb_per_split);
    const size_t smem = (size_t)q.stages * 2 * (nterms == 3 ? 2 : 1) * kTileBytes + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        CLICA_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    dim3 grid(ceil_div(g.No, BN), ceil_div(g.Mo, BM), splits);
    { LaunchScope ls(st, kFamGemmTc); gemm_tc_kernel<<<grid, kTcThreads, smem, st>>>(tAh, tAl, tBh, tBl, q); }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- per-layer entry points with plain fp32 operands: split into planes in the workspace, then GEMM -----------
namespace {
struct WsCarver {
    char* base; size_t off, cap;
    float* take(size_t floats) {
        off = align_up(off, 1024);
        float* p = (float*)(base + off);
        off += floats * sizeof(float);
        return p;
    }
};
// (hi, lo) planes of a plain matrix; in TF32 mode an aligned matrix is used in place
int make_planes(const float* src, int ld, int rows, int cols, int mode, WsCarver& ws, PlanesIn* out, cudaStream_t st) {
    if (mode == CLICA_GEMM_TF32 && (ld % 4) == 0 && (((uintptr_t)src) & 15u) == 0) {
        out->hi = src; out->lo = nullptr; out->ld = ld;
        return 0;
    }
    const int pl = plane_ld(cols);
    float* hi = ws.take((size_t)rows * pl);
    float* lo = (mode == CLICA_GEMM_3XTF32) ? ws.take((size_t)rows * pl) : nullptr;
    CLICA_REQUIRE(ws.off <= ws.cap, CLICA_E_WORKSPACE, "tensor-core GEMM workspace too small (%zu > %zu)", ws.off, ws.cap);
    out->hi = hi; out->lo = lo; out->ld = pl;
    return tc_split_planes(src, ld, rows, cols, hi, lo, pl, st);
}
}  // namespace

size_t tc_workspace_bytes(int M, int N, int K, int mode) {
    const size_t planes = (mode == CLICA_GEMM_3XTF32) ? 2 : 1;
    const size_t f = (size_t)M * plane_ld(K) + (size_t)N * plane_ld(K) + (size_t)M * plane_ld(N);
    return planes * f * sizeof(float) + 8 * 1024;
}

int tc_linear_fwd(const float* x, int ldx, const float* W, int ldw, const float* b, float* y, int ldy,
                  int M, int K, int N, float slope, int mode, void* ws, size_t ws_bytes, int sm_count, cudaStream_t st) {
    WsCarver c{(char*)ws, 0, ws_bytes};
    TcGemm g = {};
    int rc;
    if ((rc = make_planes(x, ldx, M, K, mode, c, &g.A, st))) return rc;
    if ((rc = make_planes(W, ldw, N, K, mode, c, &g.B, st))) return rc;
    g.a_mn_major = 0; g.b_mn_major = 0; g.Mo = M; g.No = N; g.Kr = K;
    g.epi = kTcBiasAct; g.bias = b; g.slope = slope; g.out = y; g.ldo = ldy;
    return tc_gemm_launch(g, sm_count, st);
}

int tc_linear_bwd_data(const float* dy, int lddy, const float* W, int ldw, const float* x_act, int ldxa,
                       float slope_prev, float* dx, int lddx, int M, int K, int N, int mode, void* ws,
                       size_t ws_bytes, int sm_count, cudaStream_t st) {
    WsCarver c{(char*)ws, 0, ws_bytes};
    TcGemm g = {};
    int rc;
    if ((rc = make_planes(dy, lddy, M, N, mode, c, &g.A, st))) return rc;     // [M][N]: K-major, reduce over N
    if ((rc = make_planes(W, ldw, N, K, mode, c, &g.B, st))) return rc;       // [N][K]: rows are the reduction -> MN-major
    g.a_mn_major = 0; g.b_mn_major = 1; g.Mo = M; g.No = K; g.Kr = N;
    g.epi = kTcMask; g.aux = x_act; g.ldaux = ldxa; g.slope = slope_prev; g.out = dx; g.ldo = lddx;
    return tc_gemm_launch(g, sm_count, st);
}

int tc_linear_bwd_weight(const float* dy, int lddy, const float* x, int ldx, float* dW, int lddw, float* db,
                         int M, int K, int N, int mode, void* ws, size_t ws_bytes, int sm_count, cudaStream_t st) {
    WsCarver c{(char*)ws, 0, ws_bytes};
    TcGemm g = {};
    int rc;
    if ((rc = make_planes(dy, lddy, M, N, mode, c, &g.A, st))) return rc;     // [M][N]: rows are the reduction -> MN-major
    if ((rc = make_planes(x, ldx, M, K, mode, c, &g.B, st))) return rc;       // [M][K]: rows are the reduction -> MN-major
    g.a_mn_major = 1; g.b_mn_major = 1; g.Mo = N; g.No = K; g.Kr = M;
    g.epi = kTcAtomic; g.out = dW; g.ldo = lddw; g.allow_split_k = 1;
    CLICA_CUDA_OK(cudaMemset2DAsync(dW, (size_t)lddw * sizeof(float), 0, (size_t)K * sizeof(float), N, st));
    if ((rc = tc_gemm_launch(g, sm_count, st))) return rc;
    if (db) return simt_colsum(dy, nullptr, lddy, M, N, db, st);
    return 0;
}

}  // namespace clica
