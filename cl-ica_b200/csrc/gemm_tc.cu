// gemm_tc.cu -- hand-written tcgen05 GEMM for the encoder's Linear layers (sm_100a).
//
//   D[128 x BN tile] = sum_r A(m, r) B(n, r)           fp32 operands consumed as TF32 by tcgen05.mma kind::tf32
//
// Replaces the cuBLAS SGEMMs behind nn.Linear fwd / AddmmBackward of /root/reference/encoders.py:39.
//
// Structure (persistent: one CTA per SM loops over output tiles; 192 threads):
//   warp 0      TMA producer  : cp.async.bulk.tensor.2d (128B-swizzled boxes) -> smem ring, mbarrier complete_tx
//   warp 1      MMA issuer    : one elected thread issues tcgen05.mma (M=128, N=BN in {128, 256}, K=8) from smem
//                               descriptors; fp32 accumulators double-buffered in TMEM (2 x BN columns x 128
//                               lanes); tcgen05.commit recycles smem stages and publishes finished accumulators
//   warps 2..9  epilogue      : tcgen05.ld (32 lanes x 32 columns per instruction) -> registers -> fused
//                               bias + LeakyReLU | activation mask | split-K atomic accumulate -> global,
//                               overlapped with the next tile's main loop
// Both operand majors are supported natively (instruction-descriptor a_major / b_major bits, MN-major
// 128B-swizzle shared-memory descriptors), so forward (x W^T), backward-data (dY W) and backward-weight
// (dY^T X) all read the row-major tensors as they lie in HBM: no transposes are ever materialised.
// 3xTF32: every stage holds A_hi, A_lo, B_hi, B_lo; three MMAs per k-step (hi*hi + lo*hi + hi*lo).
#include "gemm_tc.cuh"
#include "gemm_simt.cuh"

#include <cuda.h>   // CUtensorMap + enums only; the driver entry point is fetched through the runtime

#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

namespace clica {
namespace {

constexpr int BM = 128, BK = 32;                       // BK fp32 = 128 bytes = one swizzle span; BN = 128 or 256 (template)
constexpr int kConvWarps = 4;                            // operand-converter warps (in-kernel hi/lo split)
constexpr int kTcThreads = 320 + 32 * kConvWarps;        // TMA warp + MMA warp + 8 epilogue warps + converter warps
constexpr int kMaxStages = 8;
constexpr int kPairDefault = 1;                          // CLICA_TC_PAIR default (1: CTA pairs, tcgen05 cta_group::2)
constexpr size_t kEpiStageBytes = 8 * 4096;                 // one 32 x 32 fp32 staging block per epilogue warp

constexpr int kMaxChain = 12;                          // GEMMs one launch can chain (see ChainParams)
constexpr int kMaxBN = 256;

struct TcKernelParams {
    int Mo, No;
    int kb_total, kb_per_split, splits;
    int a_mn, b_mn, nterms;
    int conv_a, conv_b;               // 3xTF32 with a single stored fp32 plane: the converter warps split it in shared memory
    int conv_trunc;                   // split by truncation (hi = the landed word as the tensor core reads it); 0: by rounding
    int epi;
    const float* bias; float slope;
    const float* aux; int ldaux;
    float* out; int ldo;
    float* out_hi; float* out_lo; int ldp;
    uint32_t mn_lbo, mn_sbo, mn_lt;   // MN-major descriptor fields (defaults: BK*128, 512, 1)
    float* colsum;                    // optional [No]: += column sums of the stored values (pre-zeroed by the caller)
    int out_tma;                      // kTcAtomic: tmOh describes `out`, partial sums leave as TMA reduce-adds
    int kchunk;                       // > 0: the tensor core accumulates at most this many k-blocks into one TMEM accumulator;
                                      // the epilogue warps add the chunks' partial sums in fp32 registers (see the MMA issuer)
};

// One launch runs a CHAIN of GEMMs (the hidden layers of the encoder forward, or the dX / dW GEMMs of its backward) as
// one persistent grid: the work items of all GEMMs form one list that every CTA pair walks with stride #pairs, so the
// ramp (pipeline fill), the tail (last epilogue) and the partial last wave of a GEMM overlap the next GEMM's main loop
// instead of being paid per kernel.  A GEMM whose operand rows are produced by an earlier GEMM of the chain waits -- per
// 256-row block, in its TMA producer thread -- for that block's arrival counter (`flags`), which the epilogue warps of the
// producing tiles bump once their TMA stores have completed.  Every CTA is resident (grid <= #SMs, one CTA per SM) and
// items are processed in list order, so the earliest unfinished item can always proceed: no deadlock.
struct ChainGemm {
    CUtensorMap tmAh, tmAl, tmBh, tmBl, tmOh, tmOl;
    TcKernelParams q;
    int bn;                           // tile width of this GEMM (multiple of 32 * CTAS, <= kMaxBN)
    int num_m, num_n, item_begin;     // tiles and the index of its first work item in the chain's list
    int n_fastest;                    // item order: 1 = the n-tiles of one m-tile are consecutive (forward, dX: a block of
                                      // output rows completes early); 0 = m fastest, k-split slowest (dW)
    int dep;                          // >= 0: GEMM of the chain whose OUTPUT rows this GEMM reads as reduction-side input
    int dep_rows;                     // 0: the rows of the item's own m-tile (A = that output); 1: the item's reduction rows
    int dep_count;                    // arrivals of a 256-row block of `dep` that mean "all its tiles are stored"
    int publish;                      // bump flags[this GEMM][m-tile] after each tile's stores completed
};
struct ChainParams {
    ChainGemm g[kMaxChain];
    int n, total_items, stages;
    uint32_t stage_stride;            // bytes between ring stages (the widest GEMM's stage)
    unsigned* flags; int flag_stride; // [n][flag_stride] arrival counters, zero at launch
    unsigned* err;                    // set, and the kernel trapped, when a dependency wait timed out (logic error: fail loudly)
    int debug;                        // CLICA_TC_DEBUG bits (timing experiments only; results are wrong): 1 no output stores,
                                      // 2 no column sums, 4 no mask loads
};

// ---- PTX wrappers -----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// same, acquiring at cluster scope: the arrivals (and the shared-memory writes they publish) may come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP_C:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE_C;\n\t"
        "bra WAIT_LOOP_C;\n\t"
        "WAIT_DONE_C:\n\t"
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
// smem -> global tile store through the TMA (bulk async-group completion); the tensor map clips rows / columns
// that fall outside the tensor, so ragged tiles need no predication
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tm), "r"(src), "r"(x), "r"(y) : "memory");
}
// smem tile added (fp32) into the global tensor by the TMA: split-K partial sums, one request per 128-byte line
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, uint32_t src, int x, int y) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tm), "r"(src), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int CTAS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    if constexpr (CTAS == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {   // executed by one warp of EACH CTA of the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CTAS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (CTAS == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CTAS == 1) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {   // issued by the leader CTA only: M = 256 over the pair, each CTA's tensor core produces its 128 rows
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// ---- CTA-pair (cluster of 2) helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on an mbarrier that may live in the PEER CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tm), "r"(bar_cluster_addr), "r"(x), "r"(y) : "memory");
}
// tcgen05.commit that arrives on the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps only
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
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
// instruction descriptor: D fp32, A/B tf32, M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc(int a_mn, int b_mn, int n, int m = BM) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// MMA width of the n-tile that starts `rem` columns before the edge of the output: a ragged last tile issues a
// narrower MMA (multiples of 32 columns per CTA) instead of multiplying TMA zero-fill
template <int CTAS>
__device__ __forceinline__ int tile_width(int bn, int rem) {
    constexpr int G = 32 * CTAS;
    const int w = (rem + G - 1) / G * G;
    return w < bn ? w : bn;
}

// one operand tile of a stage: K-major = one {32 x rows} box (the box height is baked into the tensor map);
// MN-major = rows/32 boxes of {32 x 32}
// REMOTE: pair mode without converters -- the completion is signalled on the LEADER's barrier (cluster address)
template <bool REMOTE>
__device__ __forceinline__ void load_operand(const CUtensorMap* tm, uint32_t dst, int mn_major, int rows, int mn0, int k0,
                                             uint64_t* bar, uint32_t bar_cluster) {
    if (!mn_major) {
        if constexpr (!REMOTE) tma_load_2d(dst, tm, k0, mn0, bar);
        else tma_load_2d_pair(dst, tm, k0, mn0, bar_cluster);
    } else {
        for (int j = 0; j < rows / 32; ++j) {
            if constexpr (!REMOTE) tma_load_2d(dst + j * (BK * 128), tm, mn0 + 32 * j, k0, bar);
            else tma_load_2d_pair(dst + j * (BK * 128), tm, mn0 + 32 * j, k0, bar_cluster);
        }
    }
}

// ---- work list of a chain -----------------------------------------------------------------------------------
struct Item { int m_idx, n_idx, kb0, kb1; };
__device__ __forceinline__ Item decode_item(const ChainGemm& G, int lw) {
    Item it;
    int split;
    if (G.n_fastest) { it.n_idx = lw % G.num_n; const int r = lw / G.num_n; it.m_idx = r % G.num_m; split = r / G.num_m; }
    else { it.m_idx = lw % G.num_m; const int r = lw / G.num_m; it.n_idx = r % G.num_n; split = r / G.num_n; }
    it.kb0 = split * G.q.kb_per_split;
    it.kb1 = min(G.q.kb_total, it.kb0 + G.q.kb_per_split);
    return it;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// TMA producer: the 256-row blocks [b0, b1] of the output of GEMM G.dep are completely stored (see ChainParams)
__device__ __forceinline__ void wait_blocks(const ChainParams& P, const ChainGemm& G, int b0, int b1) {
    const unsigned* f = P.flags + (size_t)G.dep * P.flag_stride;
    const unsigned need = (unsigned)G.dep_count;
    for (int b = b0; b <= b1; ++b) {
        if (ld_acquire_u32(f + b) >= need) continue;
        const long long t0 = clock64();
        while (ld_acquire_u32(f + b) < need) {
            __nanosleep(32);
            if (clock64() - t0 > (1LL << 35)) {      // ~18 s of SM clocks: a logic error must neither hang the device nor pass silently
                atomicExch(P.err, 1u);
                __trap();                             // the launch fails; the next CUDA call of the host reports it
            }
        }
    }
    fence_proxy_async_all();       // the acquire above (generic proxy) orders the TMA reads (async proxy) that follow
}

// Persistent kernel: grid = min(#work items, #SMs); the chain's work items -- (GEMM, split, n-tile, m-tile) -- are
// walked with stride #units by every unit (a CTA or a CTA pair).  The accumulator is double-buffered in TMEM (2 x 256
// columns), so the epilogue of item i overlaps the main loop of item i+1; the smem ring runs continuously across items
// and across the GEMMs of the chain.
//
// CTAS == 2 (CTA pair, cluster of two CTAs on the two SMs of a TPC, tcgen05 cta_group::2): a work item is a
// 256 x bn tile.  Each CTA loads ITS 128 rows of A and ITS half (bn/2 rows) of B -- the pair's tensor cores read the
// other half of B from the peer's shared memory -- so a CTA pulls (128 + bn/2) operand rows per k-block from L2
// instead of (128 + bn): the kernel is paced by L2->SM operand delivery (two fp32 planes per operand), not by the
// tensor pipe.  Only the leader CTA (cluster rank 0) issues MMAs; its "stage full" barrier collects the TMA bytes of
// both CTAs, tcgen05.commit multicasts "stage free" / "accumulator ready" to both, and the epilogue warps of both
// CTAs (each drains its own 128 TMEM lanes) report to the leader's "accumulator free" barrier.
// CONV: four extra converter warps split single-plane 3xTF32 operands in shared memory (see the converter role below);
// the stage hand-over then is  TMA -> full barrier (per CTA) -> converters -> ready barrier (leader) -> MMA.  Measured on
// B200 (profiles/r2_gemm_ab.md) that extra hop costs ~25 % of the main loop's throughput at 3 stages in flight, more
// than the halved activation traffic returns, so CONV kernels are only used where an operand exists as a plain fp32
// matrix anyway (the dense input / output side of the first and last encoder layer); hidden activations keep stored
// (hi, lo) planes and the CONV == false kernel:  TMA -> full barrier (leader, both CTAs' bytes) -> MMA.
template <int CTAS, bool CONV>
__global__ void __launch_bounds__(CONV ? kTcThreads : kTcThreads - 32 * kConvWarps, 1)
gemm_tc_kernel(const __grid_constant__ ChainParams P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];     // this CTA's TMA bytes of the stage have landed
    __shared__ __align__(8) uint64_t ready_bar[kMaxStages];    // (leader CTA) the stage is converted in BOTH CTAs: MMA may read
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float bias_s[kMaxBN];

    constexpr uint32_t kABytes = BM * BK * 4;
    constexpr uint32_t kAccCols = kMaxBN;                   // TMEM columns per accumulator buffer
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
    const int unit_id = blockIdx.x / CTAS, num_units = gridDim.x / CTAS;     // a unit = one CTA or one CTA pair
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;     // swizzled tiles need 1024-byte alignment
    const int stages = P.stages;
    const int total = P.total_items;

    if (warp == 0 && lane == 0) {
        // every CTA's loads complete on its OWN full barrier; its converter warps wait there, split the fp32 planes
        // that were stored un-split, and arrive (remotely, for the peer) on the LEADER's ready barrier, which the MMA
        // warp waits on; the leader's accumulator-free barrier takes the 8 epilogue warps of each CTA
        // (CONV == false, pair mode: the leader's full barrier expects the TMA bytes of BOTH CTAs -- one arrive.expect_tx
        // by the leader's producer; the peer's loads complete_tx on it)
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], kConvWarps * CTAS); mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 8 * CTAS); }
        fence_barrier_init();
    }
    constexpr uint32_t kTmemCols = 2 * kAccCols;            // two accumulators
    if (warp == 1) tmem_alloc<CTAS>(&tmem_slot, kTmemCols);
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (CTAS == 2) cluster_sync_all();        // the peer's barriers are initialised before any remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    // programmatic dependent launch: everything above (barriers, TMEM, cluster handshake) overlapped the tail of the
    // preceding kernel; its results are complete and visible from here on
    pdl_enter();

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            uint32_t it = 0;
            int gi = 0;
            for (int w = unit_id; w < total; w += num_units) {
                while (gi + 1 < P.n && w >= P.g[gi + 1].item_begin) ++gi;
                const ChainGemm& G = P.g[gi];
                const TcKernelParams q = G.q;            // by value: registers, not re-read after every asm memory clobber
                const Item im = decode_item(G, w - G.item_begin);
                const int bnl = G.bn / CTAS;                                      // B rows this CTA loads per k-block
                const uint32_t kBBytes = (uint32_t)bnl * BK * 4;
                const uint32_t nplanes = (q.nterms == 3) ? 2u : 1u;
                const uint32_t stage_bytes = nplanes * (kABytes + kBBytes);
                const int m0 = im.m_idx * (BM * CTAS) + (int)cta_rank * BM;       // this CTA's 128 rows of A
                const int nt0 = im.n_idx * G.bn;
                const int n0 = nt0 + (int)cta_rank * (tile_width<CTAS>(G.bn, q.No - nt0) / CTAS);   // this CTA's share of B
                if (G.dep >= 0) {
                    if (G.dep_rows == 0) wait_blocks(P, G, im.m_idx * (BM * CTAS) / 256, (im.m_idx * (BM * CTAS) + BM * CTAS - 1) / 256);
                    else wait_blocks(P, G, im.kb0 * BK / 256, (im.kb1 * BK - 1) / 256);
                }
                for (int kb = im.kb0; kb < im.kb1; ++kb, ++it) {
                    const uint32_t s = it % (uint32_t)stages;
                    const uint32_t ph = (it / (uint32_t)stages) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    const uint32_t sa = tiles + s * P.stage_stride;
                    const uint32_t sb = sa + nplanes * kABytes;
                    const int k0 = kb * BK;
                    if constexpr (CONV) {
                        const uint32_t a_planes = (nplanes == 2 && !q.conv_a) ? 2u : 1u, b_planes = (nplanes == 2 && !q.conv_b) ? 2u : 1u;
                        mbar_arrive_expect_tx(&full_bar[s], a_planes * kABytes + b_planes * kBBytes);
                        load_operand<false>(&G.tmAh, sa, q.a_mn, BM, m0, k0, &full_bar[s], 0u);
                        if (a_planes == 2) load_operand<false>(&G.tmAl, sa + kABytes, q.a_mn, BM, m0, k0, &full_bar[s], 0u);
                        load_operand<false>(&G.tmBh, sb, q.b_mn, bnl, n0, k0, &full_bar[s], 0u);
                        if (b_planes == 2) load_operand<false>(&G.tmBl, sb + kBBytes, q.b_mn, bnl, n0, k0, &full_bar[s], 0u);
                    } else {
                        constexpr bool REMOTE = (CTAS == 2);
                        uint32_t fb = 0;                                 // pair mode: the LEADER's full barrier
                        if constexpr (CTAS == 1) {
                            mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
                        } else {
                            // the peer's loads for this phase cannot start before its own empty barrier flipped, i.e. before
                            // the leader's full barrier finished the previous phase: a complete_tx that overtakes this
                            // expect_tx only drives the (signed) tx-count negative for a moment
                            fb = mapa_u32(smem_u32(&full_bar[s]), 0u);
                            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], stage_bytes * CTAS);
                        }
                        load_operand<REMOTE>(&G.tmAh, sa, q.a_mn, BM, m0, k0, &full_bar[s], fb);
                        if (nplanes == 2) load_operand<REMOTE>(&G.tmAl, sa + kABytes, q.a_mn, BM, m0, k0, &full_bar[s], fb);
                        load_operand<REMOTE>(&G.tmBh, sb, q.b_mn, bnl, n0, k0, &full_bar[s], fb);
                        if (nplanes == 2) load_operand<REMOTE>(&G.tmBl, sb + kBBytes, q.b_mn, bnl, n0, k0, &full_bar[s], fb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            // ===== MMA issuer (pair mode: the leader CTA drives both tensor cores) =====
            uint32_t it = 0, acc_iter = 0;          // acc_iter: uses of the two TMEM accumulator buffers so far
            int gi = 0;
            for (int w = unit_id; w < total; w += num_units) {
                while (gi + 1 < P.n && w >= P.g[gi + 1].item_begin) ++gi;
                const ChainGemm& G = P.g[gi];
                const TcKernelParams q = G.q;            // by value: registers, not re-read after every asm memory clobber
                const Item im = decode_item(G, w - G.item_begin);
                const uint32_t kBBytes = (uint32_t)(G.bn / CTAS) * BK * 4;
                const uint32_t nplanes = (q.nterms == 3) ? 2u : 1u;
                const uint32_t a_step = q.a_mn ? 1024u : 32u, b_step = q.b_mn ? 1024u : 32u;   // bytes per K = 8 step
                const uint32_t a_lbo = q.a_mn ? q.mn_lbo : 16u, b_lbo = q.b_mn ? q.mn_lbo : 16u;
                const uint32_t a_sbo = q.a_mn ? q.mn_sbo : 1024u, b_sbo = q.b_mn ? q.mn_sbo : 1024u;
                const uint32_t a_lt = q.a_mn ? q.mn_lt : 2u, b_lt = q.b_mn ? q.mn_lt : 2u;
                const uint32_t idesc = make_idesc(q.a_mn, q.b_mn, tile_width<CTAS>(G.bn, q.No - im.n_idx * G.bn), BM * CTAS);
                // K chunks: the tensor core's fp32 accumulator TRUNCATES when it aligns addends, a bias that grows with the
                // number of MMAs accumulated in TMEM (K = 2000: 756 of them, ~6e-5 through the encoder stack).  With
                // q.kchunk > 0 every kchunk k-blocks go to the OTHER accumulator buffer and the epilogue warps add the
                // finished chunk into fp32 registers (rounded adds) while the next chunk is being multiplied.
                const int kc = (q.kchunk > 0) ? q.kchunk : (im.kb1 - im.kb0);
                for (int c0 = im.kb0; c0 < im.kb1; c0 += kc, ++acc_iter) {
                    const int c1 = min(im.kb1, c0 + kc);
                    const uint32_t as = acc_iter & 1u;
                    mbar_wait(&tmem_empty_bar[as], ((acc_iter >> 1) & 1u) ^ 1u);     // epilogue has drained this accumulator
                    tcgen05_fence_after();
                    const uint32_t tmem_acc = tmem_base + as * kAccCols;
                    for (int kb = c0; kb < c1; ++kb, ++it) {
                        const uint32_t s = it % (uint32_t)stages;
                        const uint32_t ph = (it / (uint32_t)stages) & 1u;
                        if constexpr (!CONV) mbar_wait(&full_bar[s], ph);
                        else if constexpr (CTAS == 1) mbar_wait(&ready_bar[s], ph);
                        else mbar_wait_cluster(&ready_bar[s], ph);
                        tcgen05_fence_after();
                        const uint32_t sa = tiles + s * P.stage_stride;
                        const uint32_t a_hi = sa, a_lo = sa + kABytes;
                        const uint32_t b_hi = sa + nplanes * kABytes, b_lo = b_hi + kBBytes;
#pragma unroll
                        for (int ks = 0; ks < BK / 8; ++ks) {
                            const uint32_t first = (kb > c0 || ks > 0) ? 1u : 0u;
                            const uint64_t dah = make_smem_desc(a_hi + ks * a_step, a_lbo, a_sbo, a_lt);
                            const uint64_t dbh = make_smem_desc(b_hi + ks * b_step, b_lbo, b_sbo, b_lt);
                            if (nplanes == 2) {
                                const uint64_t dal = make_smem_desc(a_lo + ks * a_step, a_lbo, a_sbo, a_lt);
                                const uint64_t dbl = make_smem_desc(b_lo + ks * b_step, b_lbo, b_sbo, b_lt);
                                umma_tf32<CTAS>(tmem_acc, dal, dbh, idesc, first);   // small terms first
                                umma_tf32<CTAS>(tmem_acc, dah, dbl, idesc, 1u);
                                umma_tf32<CTAS>(tmem_acc, dah, dbh, idesc, 1u);
                            } else {
                                umma_tf32<CTAS>(tmem_acc, dah, dbh, idesc, first);
                            }
                        }
                        // stage reusable once these MMAs have read it (pair mode: in both CTAs)
                        if constexpr (CTAS == 1) umma_commit(&empty_bar[s]); else umma_commit_pair(&empty_bar[s]);
                    }
                    // accumulator (chunk) complete (pair mode: each CTA's epilogue drains its own 128 TMEM lanes)
                    if constexpr (CTAS == 1) umma_commit(&tmem_full_bar[as]); else umma_commit_pair(&tmem_full_bar[as]);
                }
            }
        }
    } else if (CONV && warp >= 10) {
        // ===== operand converters: warps 10..13 (CONV kernels only).  3xTF32 operands that live in HBM as ONE fp32 plane (activations and
        // their gradients: half the DRAM / L2 bytes of a stored (hi, lo) pair) are split here, in shared memory, right
        // after the TMA delivered them; lo goes to the stage's lo slot at the same offset -- the layout (K-major /
        // MN-major swizzle) is irrelevant to an element-wise pass.
        const int ctid = threadIdx.x - 320;
        uint8_t* const tile_base = smem_raw + (tiles - smem_u32(smem_raw));
        // q.conv_trunc (default): the tensor core reads an fp32 word as tf32 by IGNORING its low 13 mantissa bits, so the
        // landed tile already is the hi operand (hi = v with those bits cleared) and only lo = v - hi has to be produced:
        // one AND + one subtraction per element, one store per 16 bytes; lo's own low bits are dropped by the same
        // truncation (2^-24 relative).  conv_trunc == 0 reproduces split_planes_kernel bit for bit (hi = tf32-round(v)
        // written back in place, lo = tf32-round(v - hi)) at four times the instruction count.
        constexpr uint32_t kStep = (uint32_t)kConvWarps * 32u * 16u;
        auto convert = [&](uint8_t* src, uint32_t bytes, uint32_t lo_off, int trunc) {
            if (trunc) {
                for (uint32_t i = (uint32_t)ctid * 16u; i < bytes; i += 4u * kStep) {     // 4 independent chunks in flight
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i + u * kStep < bytes) v[u] = *reinterpret_cast<const float4*>(src + i + u * kStep);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (i + u * kStep >= bytes) continue;
                        float4 l;
                        l.x = v[u].x - __uint_as_float(__float_as_uint(v[u].x) & 0xFFFFE000u);
                        l.y = v[u].y - __uint_as_float(__float_as_uint(v[u].y) & 0xFFFFE000u);
                        l.z = v[u].z - __uint_as_float(__float_as_uint(v[u].z) & 0xFFFFE000u);
                        l.w = v[u].w - __uint_as_float(__float_as_uint(v[u].w) & 0xFFFFE000u);
                        *reinterpret_cast<float4*>(src + i + u * kStep + lo_off) = l;
                    }
                }
                return;
            }
            for (uint32_t i = (uint32_t)ctid * 16u; i < bytes; i += kStep) {
                const float4 v = *reinterpret_cast<const float4*>(src + i);
                float4 h, l;
                h.x = round_to_tf32(v.x); h.y = round_to_tf32(v.y); h.z = round_to_tf32(v.z); h.w = round_to_tf32(v.w);
                l.x = round_to_tf32(v.x - h.x); l.y = round_to_tf32(v.y - h.y);
                l.z = round_to_tf32(v.z - h.z); l.w = round_to_tf32(v.w - h.w);
                *reinterpret_cast<float4*>(src + i) = h;
                *reinterpret_cast<float4*>(src + i + lo_off) = l;
            }
        };
        const uint32_t ready_addr = (CTAS == 2) ? mapa_u32(smem_u32(&ready_bar[0]), 0u) : 0u;
        uint32_t it = 0;
        int gi = 0;
        for (int w = unit_id; w < total; w += num_units) {
            while (gi + 1 < P.n && w >= P.g[gi + 1].item_begin) ++gi;
            const ChainGemm& G = P.g[gi];
            const TcKernelParams q = G.q;            // by value: registers, not re-read after every asm memory clobber
            const Item im = decode_item(G, w - G.item_begin);
            const uint32_t kBBytes = (uint32_t)(G.bn / CTAS) * BK * 4;
            for (int kb = im.kb0; kb < im.kb1; ++kb, ++it) {
                const uint32_t s = it % (uint32_t)stages;
                const uint32_t ph = (it / (uint32_t)stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                if (q.nterms == 3 && (q.conv_a || q.conv_b)) {
                    uint8_t* sa = tile_base + s * P.stage_stride;
                    if (q.conv_a) convert(sa, kABytes, kABytes, q.conv_trunc);
                    if (q.conv_b) convert(sa + 2 * kABytes, kBBytes, kBBytes, q.conv_trunc);
                    fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's reads
                }
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CTAS == 1) mbar_arrive(&ready_bar[s]);
                    else mbar_arrive_cluster(ready_addr + s * (uint32_t)sizeof(uint64_t));
                }
            }
        }
    } else {
        // ===== epilogue: warps 2..9 (8 warps).  TMEM lane quarter = warp % 4; the two warps of a quarter take the
        // even / odd 32-column chunks.  A thread reads one accumulator row (32 consecutive columns per tcgen05.ld).
        // Planar outputs leave through the TMA: the warp stages its 32 x 32 block in a private 4 KB shared-memory
        // buffer (128B-swizzled: conflict-free STS.128) and one lane issues a bulk tensor store -- whole 128-byte
        // lines per request instead of 32 partial-sector writes per STG instruction, and rows / columns past the
        // tensor edge are clipped by the tensor map.  The staged block also yields the column sums (bias
        // gradients) with 32 conflict-free LDS per lane instead of 160 shuffles.
        const int ew = warp - 2;
        const int wq = warp & 3;
        const int half = ew >> 2;
        const int etid = threadIdx.x - 64;                     // 0..255 among the epilogue threads
        uint8_t* const sbuf = smem_raw + (tiles - smem_u32(smem_raw)) + (uint32_t)stages * P.stage_stride + (uint32_t)ew * 4096u;
        const uint32_t sbuf_u32 = smem_u32(sbuf);
        float4* const srow = reinterpret_cast<float4*>(sbuf + lane * 128);
        // stage the warp's 32 x 32 block (thread = row, u = its 32 columns) once the previous store has read the buffer
        auto stage = [&](const float (&u)[32]) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) srow[j ^ (lane & 7)] = make_float4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
        };
        // column `lane` of the staged block summed over its 32 rows
        auto staged_colsum = [&]() {
            float t = 0.f;
            const uint32_t cw = (uint32_t)(lane >> 2), ci = (uint32_t)(lane & 3) * 4u;
#pragma unroll
            for (int r = 0; r < 32; ++r) t += *reinterpret_cast<const float*>(sbuf + r * 128 + ((cw ^ (uint32_t)(r & 7)) << 4) + ci);
            return t;
        };
        uint32_t acc_iter = 0;                  // uses of the two TMEM accumulator buffers so far (same count as the MMA issuer's)
        float ksum[(kMaxBN / 64) * 32];          // K-chunked items: this thread's running sums (4 column chunks x 32 columns)
        int gi = 0;
        for (int w = unit_id; w < total; w += num_units) {
            while (gi + 1 < P.n && w >= P.g[gi + 1].item_begin) ++gi;
            const ChainGemm& G = P.g[gi];
            const TcKernelParams q = G.q;            // by value: registers, not re-read after every asm memory clobber
            const Item im = decode_item(G, w - G.item_begin);
            const int bn = G.bn, publish = G.publish;
            const CUtensorMap* const tmOh = &G.tmOh;
            const CUtensorMap* const tmOl = &G.tmOl;
            const bool aux_vec = (q.aux != nullptr) && ((q.ldaux & 3) == 0) && ((reinterpret_cast<uintptr_t>(q.aux) & 15u) == 0);
            const int m0 = im.m_idx * (BM * CTAS) + (int)cta_rank * BM;
            const int n0 = im.n_idx * bn;
            const int row = m0 + wq * 32 + lane;
            const bool row_ok = row < q.Mo;
            const int rowc = row_ok ? row : (q.Mo - 1);         // clamped: loads stay in bounds, stores are predicated
            if (q.epi == kTcBiasAct) {
                epi_bar_sync();                                 // previous tile's readers are done with bias_s
                if (etid < bn) bias_s[etid] = (q.bias != nullptr && n0 + etid < q.No) ? __ldg(q.bias + n0 + etid) : 0.f;
                epi_bar_sync();
            }
            const float* arow = (q.aux != nullptr) ? q.aux + (size_t)rowc * q.ldaux : nullptr;
            float4 av[8];
            auto load_aux = [&](int chunk) {                    // mask source for columns [n0 + 32*chunk, +32) of this row
                const int c0 = n0 + chunk * 32;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c0 + 4 * j;
                    if (aux_vec) {
                        av[j] = (c < q.ldaux) ? __ldg(reinterpret_cast<const float4*>(arow + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    } else {
                        av[j].x = (c < q.No) ? __ldg(arow + c) : 0.f;
                        av[j].y = (c + 1 < q.No) ? __ldg(arow + c + 1) : 0.f;
                        av[j].z = (c + 2 < q.No) ? __ldg(arow + c + 2) : 0.f;
                        av[j].w = (c + 3 < q.No) ? __ldg(arow + c + 3) : 0.f;
                    }
                }
            };
            const bool use_aux = (q.epi == kTcMask) && (q.aux != nullptr) && !(P.debug & 4);
            if (use_aux) load_aux(half);
            // K-chunked item (see the MMA issuer): add every finished chunk of k-blocks into fp32 registers and hand its
            // accumulator buffer straight back; the epilogue below then works on the register sums
            const int kc = (q.kchunk > 0) ? q.kchunk : (im.kb1 - im.kb0);
            const bool multi = (im.kb1 - im.kb0) > kc;
            uint32_t as = acc_iter & 1u;
            if (multi) {
                for (int c0 = im.kb0; c0 < im.kb1; c0 += kc, ++acc_iter) {
                    const uint32_t asc = acc_iter & 1u;
                    mbar_wait(&tmem_full_bar[asc], (acc_iter >> 1) & 1u);
                    tcgen05_fence_after();
                    int ci = 0;
#pragma unroll 1
                    for (int chunk = half; chunk < bn / 32; chunk += 2, ++ci) {
                        float t[32];
                        __syncwarp();
                        tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + asc * kAccCols + (uint32_t)(chunk * 32), t);
                        if (c0 == im.kb0) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) ksum[ci * 32 + j] = t[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) ksum[ci * 32 + j] += t[j];
                        }
                    }
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CTAS == 1) mbar_arrive(&tmem_empty_bar[asc]);
                        else mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[asc]), 0u));
                    }
                }
            } else {
                mbar_wait(&tmem_full_bar[as], (acc_iter >> 1) & 1u);
                tcgen05_fence_after();
            }
            int ci_out = 0;
#pragma unroll 1
            for (int chunk = half; chunk < bn / 32; chunk += 2, ++ci_out) {
                float v[32];
                __syncwarp();   // tcgen05.ld is .sync.aligned: re-converge after the predicated stores below
                if (multi) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = ksum[ci_out * 32 + j];
                } else {
                    tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + as * kAccCols + (uint32_t)(chunk * 32), v);
                }
                const int col0 = n0 + chunk * 32;
                const int nvalid = min(32, q.No - col0);
                if (q.epi == kTcBiasAct) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bv = *reinterpret_cast<const float4*>(&bias_s[chunk * 32 + j]);
                        v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * q.slope;
                } else if (use_aux) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[4 * j] *= av[j].x > 0.f ? 1.f : q.slope; v[4 * j + 1] *= av[j].y > 0.f ? 1.f : q.slope;
                        v[4 * j + 2] *= av[j].z > 0.f ? 1.f : q.slope; v[4 * j + 3] *= av[j].w > 0.f ? 1.f : q.slope;
                    }
                    if (chunk + 2 < bn / 32) load_aux(chunk + 2);       // prefetch for the next iteration
                }
                if (nvalid <= 0) continue;                              // warp-uniform
                if (P.debug & 1) continue;
                if (q.epi == kTcAtomic && q.out_tma) {
                    stage(v);
                    if (lane == 0) { tma_reduce_add_2d(tmOh, sbuf_u32, col0, m0 + wq * 32); bulk_commit(); }
                    continue;
                }
                if (q.epi == kTcAtomic) {
                    if (!row_ok) continue;
                    float* op = q.out + (size_t)row * q.ldo + col0;
                    if (nvalid == 32 && (q.ldo & 3) == 0 && ((reinterpret_cast<uintptr_t>(q.out) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) red_add_v4(op + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) atomicAdd(op + j, v[j]);
                    }
                    continue;
                }
                if (q.out != nullptr && row_ok) {
                    float* op = q.out + (size_t)row * q.ldo + col0;
                    if (nvalid == 32 && (q.ldo & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) op[j] = v[j];
                    }
                }
                const bool want_sum = (q.colsum != nullptr) && !(P.debug & 2);
                if (q.out_hi == nullptr && !want_sum) continue;
                // rows past M hold zero accumulators (TMA zero-fill) but a bias may have been added: keep them out of
                // the column sums (the tensor map keeps them out of the stores)
                if (want_sum && !row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                float csum = 0.f;
                const int wrow0 = m0 + wq * 32;
                if (q.out_hi != nullptr && q.out_lo != nullptr) {
                    float h[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) { h[j] = round_to_tf32(v[j]); v[j] = round_to_tf32(v[j] - h[j]); }
                    stage(h);
                    if (lane == 0) { tma_store_2d(tmOh, sbuf_u32, col0, wrow0); bulk_commit(); }
                    if (want_sum) csum = staged_colsum();
                    stage(v);
                    if (lane == 0) { tma_store_2d(tmOl, sbuf_u32, col0, wrow0); bulk_commit(); }
                    if (want_sum) csum += staged_colsum();
                } else {
                    stage(v);
                    if (q.out_hi != nullptr && lane == 0) { tma_store_2d(tmOh, sbuf_u32, col0, wrow0); bulk_commit(); }
                    if (want_sum) csum = staged_colsum();
                }
                if (want_sum && lane < nvalid) atomicAdd(q.colsum + col0 + lane, csum);
            }
            // this warp has read its share of the accumulator: hand it back to the MMA warp (K-chunked items did so per chunk)
            if (!multi) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CTAS == 1) mbar_arrive(&tmem_empty_bar[as]);
                    else mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[as]), 0u));
                }
                ++acc_iter;
            }
            if (publish) {
                // a later GEMM of the chain reads these rows: once every epilogue warp's stores are complete (not merely
                // read out of shared memory), count this CTA's tile in the arrival counter of its 256-row block
                if (lane == 0) bulk_wait_all0();
                __syncwarp();
                epi_bar_sync();
                if (etid == 0) {
                    fence_proxy_async_all();       // async-proxy (TMA) writes before the generic-proxy release below
                    __threadfence();
                    atomicAdd(P.flags + (size_t)gi * P.flag_stride + (m0 / 256), 1u);
                }
            }
        }
        if (lane == 0) bulk_wait_all0();      // the staging buffer must outlive the last bulk store
        __syncwarp();
    }
    tcgen05_fence_before();
    __syncthreads();
    if constexpr (CTAS == 2) cluster_sync_all();    // the peer's shared memory / TMEM stay alive until both are done
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<CTAS>(tmem_base, kTmemCols);
    }
}

// split a plain matrix into planes; columns [cols, ld_dst) of the planes are zero-filled
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src, int ld_src, int rows, int cols,
                                                            float* __restrict__ hi, float* __restrict__ lo, int ld_dst) {
    pdl_enter();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)rows * ld_dst) return;
    const int r = (int)(idx / ld_dst), c = (int)(idx - (long long)r * ld_dst);
    const float v = (c < cols) ? __ldg(src + (size_t)r * ld_src + c) : 0.f;
    if (lo != nullptr) {
        const float h = round_to_tf32(v);
        hi[idx] = h;
        lo[idx] = round_to_tf32(v - h);   // rounded (not truncated by the MMA): keeps the split unbiased
    } else {
        hi[idx] = v;
    }
}

constexpr int kSplitMaxJobs = 8;
struct SplitMultiArgs {
    SplitJob job[kSplitMaxJobs];
    int block_start[kSplitMaxJobs + 1];     // prefix sums of ceil(rows * ld_dst / 256)
    int n;
};
__global__ void __launch_bounds__(256) split_planes_multi_kernel(const SplitMultiArgs a) {
    pdl_enter();
    int t = 0;
    while (t + 1 < a.n && a.block_start[t + 1] <= (int)blockIdx.x) ++t;
    const SplitJob& j = a.job[t];
    const long long idx = (long long)((int)blockIdx.x - a.block_start[t]) * 256 + threadIdx.x;
    if (idx >= (long long)j.rows * j.ld_dst) return;
    const int r = (int)(idx / j.ld_dst), c = (int)(idx - (long long)r * j.ld_dst);
    const float v = (c < j.cols) ? __ldg(j.src + (size_t)r * j.ld_src + c) : 0.f;
    if (j.lo != nullptr) {
        const float h = round_to_tf32(v);
        j.hi[idx] = h;
        j.lo[idx] = round_to_tf32(v - h);
    } else {
        j.hi[idx] = v;
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

// Tile width and split-K factor.  Measured in round 1 (clock64 phase stamps): once the ring is flowing, a CTA's main loop
// is paced by the operand bytes it pulls through the TMA (~95-105 GB/s per SM whether 96 or 148 CTAs run), not by
// the tensor pipe, so an item costs  kb * (BM + BN) * BK * 4 * planes  bytes of loads plus an epilogue that moves
// BM * BN * 4 * planes output bytes (weighted x4: write / reduce traffic drains slower than loads stream), and
// a persistent CTA works through ceil(items / #SMs) items.  Pick the (BN, splits) with the shortest makespan;
// ties go to the wider tile (less operand traffic per flop).
// ctas == 2: CTA pairs -- a unit of work is a 256 x bn tile on one of sm_count / 2 pairs, and each CTA of the pair
// pulls (BM + bn / 2) operand rows per k-block.
struct TilePlan { int bn; int splits; };
TilePlan plan_tiles(int Mo, int No, int kb_total, int sm_count, bool split_k, int ctas) {
    const int forced = env_int("CLICA_TC_BN", 0);
    const int widths[3] = {256, 192, 128};
    TilePlan best = {256, 1};
    double best_cost = 1e300;
    const int units = sm_count / ctas > 0 ? sm_count / ctas : 1;
    for (int i = 0; i < 3; ++i) {
        const int bn = widths[i];
        if (forced == 128 || forced == 192 || forced == 256) { if (bn != forced) continue; }
        else if (bn > 128 && No <= bn - 64) continue;           // a narrower tile already covers the whole output
        const long long tiles = (long long)ceil_div(Mo, BM * ctas) * ceil_div(No, bn);
        const int max_splits = split_k ? (kb_total / 4 > 0 ? kb_total / 4 : 1) : 1;
        for (int sp = 1; sp <= max_splits; ++sp) {
            const int kb_per = ceil_div(kb_total, sp);
            const int sp_eff = ceil_div(kb_total, kb_per);
            if (sp_eff != sp) continue;                         // same work division as a smaller split count
            const long long items = tiles * sp;
            const long long waves = (items + units - 1) / units;
            const double item_cost = (double)kb_per * (BM + bn / ctas) + 4.0 * bn * (split_k ? 2.0 : 1.0);
            const double cost = (double)waves * item_cost;
            if (cost < best_cost * 0.999) { best_cost = cost; best.bn = bn; best.splits = sp; }
            if (items > 8LL * units) break;
        }
    }
    return best;
}

bool tc_shape_ok(int M_out, int N_out, int K_red) { return M_out >= 32 && N_out >= 32 && K_red >= 32; }

int tc_split_planes(const float* src, int ld_src, int rows, int cols, float* hi, float* lo, int ld_dst, cudaStream_t st) {
    const long long n = (long long)rows * ld_dst;
    { LaunchScope ls(st, kFamMisc); launch_k(split_planes_kernel, (unsigned)((n + 255) / 256), 256, 0, st, src, ld_src, rows, cols, hi, lo, ld_dst); }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}

int tc_split_planes_multi(const SplitJob* jobs, int n, cudaStream_t st) {
    for (int first = 0; first < n; first += kSplitMaxJobs) {
        SplitMultiArgs a;
        a.n = (n - first < kSplitMaxJobs) ? (n - first) : kSplitMaxJobs;
        int blocks = 0;
        for (int t = 0; t < a.n; ++t) {
            a.job[t] = jobs[first + t];
            a.block_start[t] = blocks;
            blocks += (int)(((long long)a.job[t].rows * a.job[t].ld_dst + 255) / 256);
        }
        a.block_start[a.n] = blocks;
        if (blocks == 0) continue;
        { LaunchScope ls(st, kFamMisc); launch_k(split_planes_multi_kernel, blocks, 256, 0, st, a); }
        CLICA_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

namespace {
// everything tc_gemm_launch decides before the launch: tile plan, tensor maps, kernel parameters of ONE GEMM
struct TcPrepared {
    ChainGemm cg;
    int ctas; bool conv;
    size_t stage_bytes;
    int items;
};
int tc_prepare(const TcGemm& g, int sm_count, TcPrepared* out, bool in_chain = false) {
    CLICA_REQUIRE(g.Mo >= 1 && g.No >= 1 && g.Kr >= 1, CLICA_E_BADARG, "tc_gemm: empty problem");
    CLICA_REQUIRE(g.A.hi && g.B.hi, CLICA_E_BADARG, "tc_gemm: null operand");
    CLICA_REQUIRE(g.epi != kTcAtomic || g.out != nullptr, CLICA_E_BADARG, "tc_gemm: atomic epilogue needs a plain output");
    // 3xTF32 when asked for (g.nterms == 3) or when any operand comes as a (hi, lo) pair; an operand WITHOUT a lo
    // plane is then split inside the kernel by the converter warps
    const int nterms = (g.nterms == 3 || g.A.lo || g.B.lo) ? 3 : 1;
    const int conv_a = (nterms == 3 && !g.A.lo) ? 1 : 0, conv_b = (nterms == 3 && !g.B.lo) ? 1 : 0;
    const int nplanes = (nterms == 3) ? 2 : 1;
    // CTA pairs (tcgen05 cta_group::2) whenever the output has more than one 128-row tile; CLICA_TC_PAIR=0 disables
    //   1: always;  2: only where the isolated per-shape timings favoured pairs (tools/gemm_bench.py): weight-gradient
    //   GEMMs always, forward GEMMs wider than one tile, masked backward-data GEMMs with >= 32 k-blocks
    const int pair_mode = env_int("CLICA_TC_PAIR", kPairDefault);
    // (a GEMM with a single 128-row tile -- e.g. the weight gradient of a 100-wide layer -- still runs as a pair when it is
    // part of a chain: the peer CTA multiplies TMA zero-fill, which costs less than breaking the chain into three launches)
    bool pair_ok = pair_mode != 0 && (g.Mo > BM || in_chain) && sm_count >= 2;
    if (pair_ok && pair_mode == 2) {
        if (g.epi == kTcBiasAct) pair_ok = g.No > 128;
        else if (g.epi == kTcMask) pair_ok = ceil_div(g.Kr, BK) >= 32;
    }
    const int ctas = pair_ok ? 2 : 1;
    const TilePlan tp = plan_tiles(g.Mo, g.No, ceil_div(g.Kr, BK), sm_count, g.epi == kTcAtomic && g.allow_split_k, ctas);
    const int bn = tp.bn;
    const int bnl = bn / ctas;                                  // B rows one CTA loads per k-block
    ChainGemm& cg = out->cg;
    memset(&cg, 0, sizeof(cg));
    int rc;
    // storage shape of each operand: K-major [MN rows][Kr cols] (box = tile rows); MN-major [Kr rows][MN cols] (box 32 rows)
    const int a_rows = g.a_mn_major ? g.Kr : g.Mo, a_cols = g.a_mn_major ? g.Mo : g.Kr, a_box = g.a_mn_major ? BK : BM;
    const int b_rows = g.b_mn_major ? g.Kr : g.No, b_cols = g.b_mn_major ? g.No : g.Kr, b_box = g.b_mn_major ? BK : bnl;
    if ((rc = get_tensor_map(g.A.hi, a_rows, a_cols, g.A.ld, a_box, g.a_mn_major != 0, &cg.tmAh))) return rc;
    if ((rc = get_tensor_map(g.B.hi, b_rows, b_cols, g.B.ld, b_box, g.b_mn_major != 0, &cg.tmBh))) return rc;
    cg.tmAl = cg.tmAh; cg.tmBl = cg.tmBh;
    if (g.A.lo && (rc = get_tensor_map(g.A.lo, a_rows, a_cols, g.A.ld, a_box, g.a_mn_major != 0, &cg.tmAl))) return rc;
    if (g.B.lo && (rc = get_tensor_map(g.B.lo, b_rows, b_cols, g.B.ld, b_box, g.b_mn_major != 0, &cg.tmBl))) return rc;
    // planar outputs are written by TMA stores of 32 x 32 blocks (clipped to [Mo x No] by the map)
    cg.tmOh = cg.tmAh; cg.tmOl = cg.tmAh;
    if (g.outp.hi && (rc = get_tensor_map(g.outp.hi, g.Mo, g.No, g.outp.ld, 32, false, &cg.tmOh))) return rc;
    if (g.outp.lo && (rc = get_tensor_map(g.outp.lo, g.Mo, g.No, g.outp.ld, 32, false, &cg.tmOl))) return rc;
    TcKernelParams& q = cg.q;
    q.out_tma = 0;
    if (g.epi == kTcAtomic && (g.ldo % 4) == 0 && (((uintptr_t)g.out) & 15u) == 0 && env_int("CLICA_TC_RED_TMA", 1)) {
        if ((rc = get_tensor_map(g.out, g.Mo, g.No, g.ldo, 32, false, &cg.tmOh))) return rc;
        q.out_tma = 1;
    }
    q.Mo = g.Mo; q.No = g.No;
    q.kb_total = ceil_div(g.Kr, BK);
    q.a_mn = g.a_mn_major; q.b_mn = g.b_mn_major; q.nterms = nterms; q.conv_a = conv_a; q.conv_b = conv_b;
    q.conv_trunc = env_int("CLICA_TC_CONV_TRUNC", 1);
    q.epi = g.epi; q.bias = g.bias; q.slope = g.slope; q.aux = g.aux; q.ldaux = g.ldaux;
    q.out = g.out; q.ldo = g.ldo; q.out_hi = g.outp.hi; q.out_lo = g.outp.lo; q.ldp = g.outp.ld;
    q.mn_lbo = (uint32_t)env_int("CLICA_TC_MN_LBO", BK * 128);   // debug overrides of the MN-major descriptor
    q.mn_sbo = (uint32_t)env_int("CLICA_TC_MN_SBO", 512);
    q.mn_lt = (uint32_t)env_int("CLICA_TC_MN_LT", 1);
    q.colsum = g.colsum;
    int splits = tp.splits;
    q.kb_per_split = ceil_div(q.kb_total, splits);
    splits = ceil_div(q.kb_total, q.kb_per_split);
    q.splits = splits;
    // K-chunked accumulation (3xTF32 only, reductions of >= 3 chunks only): CLICA_TC_KCHUNK k-blocks per TMEM accumulator
    // (0: off).  Measured at n = 40 (tools/n40_accuracy.py; error through the 7-layer stack relative to the tensor max,
    // output / input gradient / worst dW): off 1.1e-5 / 4.8e-5 / 4.8e-5, 16 k-blocks 3.7e-6 / 1.4e-5 / 1.6e-5, 8: 2.6e-6 /
    // 8e-6 / 1.4e-5, 4: 1.1e-6 / 4e-6 / 6e-6; cost at config 3 (every drain reads TMEM while the tensor core writes the
    // other buffer): 16 -> +6 %, 8 -> +13 % per step.  n = 10 (K <= 500: 16 k-blocks) is not chunked.
    {
        const int kc = env_int("CLICA_TC_KCHUNK", 16);
        q.kchunk = (nterms == 3 && kc > 0 && q.kb_per_split >= 3 * kc) ? kc : 0;
    }
    cg.bn = bn;
    cg.num_m = ceil_div(g.Mo, BM * ctas);
    cg.num_n = ceil_div(g.No, bn);
    cg.dep = -1;
    out->ctas = ctas;
    out->conv = (conv_a || conv_b);
    out->stage_bytes = (size_t)nplanes * (BM + bnl) * BK * 4;
    out->items = cg.num_m * cg.num_n * splits;
    return 0;
}

// one launch for the GEMMs in P (P.g[i], item_begin / dep / publish already set); flags: [n][flag_stride], zeroed here
int tc_launch(ChainParams& P, int ctas, bool conv, size_t max_stage_bytes, int sm_count, cudaStream_t st) {
    // SMs left free for a concurrent kernel (the NCCL all-reduce of the previous gradient bucket in the multi-GPU
    // step): a persistent grid that does not fit entirely would serialise its last CTAs behind the first ones
    { const int r = tc_sm_reserve(); if (r > 0 && sm_count - r >= 8) sm_count -= r; }
    const size_t smem_cap = 227 * 1024 - 2048;                  // static shared memory (barriers, bias slice) + slack
    const size_t smem_fixed = 1024 + kEpiStageBytes;            // alignment slack + 8 x 4 KB epilogue staging
    P.stages = (int)((smem_cap - smem_fixed) / max_stage_bytes);
    if (P.stages > kMaxStages) P.stages = kMaxStages;
    { const int so = env_int("CLICA_TC_STAGES", 0); if (so >= 1 && so < P.stages) P.stages = so; }   // debug override
    P.stage_stride = (uint32_t)max_stage_bytes;
    P.debug = env_int("CLICA_TC_DEBUG", 0);
    const size_t smem = (size_t)P.stages * max_stage_bytes + smem_fixed;
    {   // the opt-in is per device (context)
        static PerDeviceOnce attr;
        if (first_on_this_device(attr)) {
            const int max_dyn = (int)smem_cap;
#define CLICA_TC_ATTR(C_, V_) CLICA_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<C_, V_>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn))
            CLICA_TC_ATTR(1, false); CLICA_TC_ATTR(2, false); CLICA_TC_ATTR(1, true); CLICA_TC_ATTR(2, true);
#undef CLICA_TC_ATTR
        }
    }
    const int units = sm_count / ctas;
    const int grid = (P.total_items < units ? P.total_items : units) * ctas;
    const int threads = conv ? kTcThreads : kTcThreads - 32 * kConvWarps;
    {
        LaunchScope ls(st, kFamGemmTc);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see launch_k (common.cuh)
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = env_flag("CLICA_PDL", 1) != 0 ? 2 : 1;
        cudaError_t e;
        if (ctas == 1 && !conv) e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<1, false>, P);
        else if (ctas == 1) e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<1, true>, P);
        else if (!conv) e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<2, false>, P);
        else e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<2, true>, P);
        CLICA_CUDA_OK(e);
    }
    CLICA_CUDA_OK(cudaGetLastError());
    return 0;
}
}  // namespace

int tc_gemm_launch(const TcGemm& g, int sm_count, cudaStream_t st) {
    TcPrepared pr;
    int rc = tc_prepare(g, sm_count, &pr);
    if (rc) return rc;
    static thread_local ChainParams P;          // ~12 KB: kept off the stack
    P.g[0] = pr.cg;
    P.g[0].item_begin = 0; P.g[0].n_fastest = 0; P.g[0].dep = -1; P.g[0].publish = 0;
    P.n = 1; P.total_items = pr.items; P.flags = nullptr; P.flag_stride = 0; P.err = nullptr;
    return tc_launch(P, pr.ctas, pr.conv, pr.stage_bytes, sm_count, st);
}

size_t tc_chain_flag_bytes(int n_links, int max_rows) {
    return ((size_t)n_links * (size_t)ceil_div(max_rows, 256) + 1) * sizeof(unsigned);
}

// Runs links[0..n) in order.  Maximal runs of consecutive links that all take the CTA-pair kernel without converter warps
// are issued as ONE launch each (see ChainParams); every other link is its own launch (stream order then provides the
// dependencies).  CLICA_TC_CHAIN=0 always launches per GEMM.
int tc_chain_launch(const TcChainLink* links, int n, int sm_count, void* flag_ws, size_t flag_ws_bytes, cudaStream_t st) {
    if (n <= 0) return 0;
    CLICA_REQUIRE(n <= kTcMaxLinks, CLICA_E_BADARG, "tc_chain: %d links > %d", n, kTcMaxLinks);
    static thread_local TcPrepared pr[kTcMaxLinks];
    static thread_local ChainParams P;
    const bool enabled = env_int("CLICA_TC_CHAIN", 1) != 0 && flag_ws != nullptr;
    for (int i = 0; i < n; ++i) {
        int rc = tc_prepare(links[i].g, sm_count, &pr[i], enabled && n > 1);
        if (rc) return rc;
    }
    auto chainable = [&](int i) { return enabled && pr[i].ctas == 2 && !pr[i].conv; };
    int i = 0;
    while (i < n) {
        int j = i + 1;
        if (chainable(i)) {
            int rows = pr[i].cg.q.Mo;
            while (j < n && j - i < kMaxChain && chainable(j) && pr[j].cg.q.nterms == pr[i].cg.q.nterms) {
                const int r2 = pr[j].cg.q.Mo > rows ? pr[j].cg.q.Mo : rows;
                if (tc_chain_flag_bytes(j - i + 1, r2) > flag_ws_bytes) break;
                rows = r2;
                ++j;
            }
        }
        const int m = j - i;
        int max_rows = 1;
        size_t max_stage = 0;
        int items = 0;
        for (int k = 0; k < m; ++k) {
            P.g[k] = pr[i + k].cg;
            ChainGemm& G = P.g[k];
            G.item_begin = items;
            items += pr[i + k].items;
            G.n_fastest = (m > 1 && G.q.splits == 1 && G.q.epi != kTcAtomic) ? 1 : 0;
            const int dep = links[i + k].dep;
            CLICA_REQUIRE(dep < i + k, CLICA_E_BADARG, "tc_chain: link %d depends on a later link %d", i + k, dep);
            G.dep = (dep >= i) ? dep - i : -1;       // a dependency outside this launch is ordered by the stream
            G.dep_rows = links[i + k].dep_rows;
            G.publish = 0;
            if (G.q.Mo > max_rows) max_rows = G.q.Mo;
            if (pr[i + k].stage_bytes > max_stage) max_stage = pr[i + k].stage_bytes;
        }
        bool any_dep = false;
        for (int k = 0; k < m; ++k) {
            ChainGemm& G = P.g[k];
            if (G.dep < 0) continue;
            ChainGemm& D = P.g[G.dep];
            CLICA_REQUIRE(D.q.splits == 1 && D.q.epi != kTcAtomic, CLICA_E_BADARG, "tc_chain: a split-K GEMM cannot be a dependency");
            D.publish = 1;
            G.dep_count = D.num_n * 2;              // both CTAs of a pair count every tile of the 256-row block
            any_dep = true;
        }
        P.n = m; P.total_items = items;
        P.flag_stride = ceil_div(max_rows, 256);
        P.flags = any_dep ? (unsigned*)flag_ws : nullptr;
        P.err = any_dep ? P.flags + (size_t)m * P.flag_stride : nullptr;
        if (any_dep) CLICA_CUDA_OK(cudaMemsetAsync(flag_ws, 0, tc_chain_flag_bytes(m, max_rows), st));
        int rc = tc_launch(P, pr[i].ctas, pr[i].conv, max_stage, sm_count, st);
        if (rc) return rc;
        i = j;
    }
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
