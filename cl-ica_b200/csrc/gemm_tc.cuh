// gemm_tc.cuh -- interface of the tcgen05 (5th-gen tensor core) TF32 / 3xTF32 GEMM path (gemm_tc.cu).
//
// Operands are "planar": a matrix is one fp32 plane (TF32 mode: the hardware drops the low 13 mantissa bits)
// or a (hi, lo) pair of planes with  value = hi + lo,  hi exactly representable in tf32 (3xTF32 mode:
// D += A_hi B_hi + A_lo B_hi + A_hi B_lo  -- three tensor-core MMAs per product, ~2^-21 relative error).
// In 3xTF32 mode an operand may also be given as ONE fp32 plane (lo == nullptr with TcGemm::nterms == 3): the kernel's
// converter warps then produce the (hi, lo) pair in shared memory, tile by tile, after the TMA delivered the fp32 data
// -- activations and gradients are stored that way (half the HBM / L2 bytes), weights are pre-split once per step.
#pragma once
#include "common.cuh"

namespace clica {

struct PlanesIn { const float* hi; const float* lo; int ld; };    // lo == nullptr: single plane
struct PlanesOut { float* hi; float* lo; int ld; };               // hi == nullptr: not written

enum TcEpilogue : int {
    kTcBiasAct = 0,   // v = leaky(acc + bias[n], slope)
    kTcMask = 1,      // v = acc * (aux[m,n] > 0 ? 1 : slope)      (aux nullable: no mask)
    kTcAtomic = 2,    // atomicAdd(out[m,n], acc)                  (split-K; out pre-zeroed by the caller)
};

// D[Mo x No] = sum_{r < Kr} A(mo, r) * B(no, r)
struct TcGemm {
    PlanesIn A; int a_mn_major;   // 0: stored [Mo rows][Kr cols] (K-major); 1: stored [Kr rows][Mo cols] (MN-major)
    PlanesIn B; int b_mn_major;   // 0: stored [No rows][Kr cols];           1: stored [Kr rows][No cols]
    int Mo, No, Kr;
    int epi;
    const float* bias; float slope;
    const float* aux; int ldaux;
    float* out; int ldo;          // plain fp32 output (nullable unless epi == kTcAtomic)
    PlanesOut outp;               // planar output (hi = round-to-tf32(v), lo = v - hi), nullable
    int allow_split_k;            // epi == kTcAtomic only
    int nterms;                   // 3: 3xTF32 even when an operand has no lo plane (it is split inside the kernel); 0: infer
    float* colsum;                // optional [No]: += column sums of the stored values (caller zeroes it first)
};

int tc_gemm_launch(const TcGemm& g, int sm_count, cudaStream_t st);

// A chain of GEMMs issued as ONE persistent launch (gemm_tc.cu, ChainParams): link i may read -- as its A operand
// (dep_rows = 0: the rows of each output tile) or as reduction rows (dep_rows = 1: dW = g^T x) -- the planar OUTPUT of an
// earlier link `dep`; tiles wait per 256-row block for the producing tiles instead of for a kernel boundary.
constexpr int kTcMaxLinks = 160;
struct TcChainLink { TcGemm g; int dep; int dep_rows; };
size_t tc_chain_flag_bytes(int n_links, int max_rows);      // arrival counters the chain needs (caller-provided workspace)
int tc_chain_launch(const TcChainLink* links, int n, int sm_count, void* flag_ws, size_t flag_ws_bytes, cudaStream_t st);

// ld of a plane holding `cols` columns: TMA needs 16-byte row multiples; rows padded to whole 128-byte lines make
// every 128-byte box row exactly one L2 line (an unaligned 2000-byte pitch splits each into two requests)
inline int plane_ld(int cols) { return (cols + 31) / 32 * 32; }

// split a plain matrix into (hi, lo) planes (lo == nullptr: plain copy with the plane's ld)
int tc_split_planes(const float* src, int ld_src, int rows, int cols, float* hi, float* lo, int ld_dst,
                    cudaStream_t st);

// the (hi, lo) planes of up to 8 matrices in ONE launch (default; CLICA_PACK_FUSED=0: one launch per matrix) -- the
// encoder re-packs its five hidden weight matrices after every optimizer step, and five 3 us launches cost more than
// the 5 MB they move
struct SplitJob { const float* src; int ld_src; int rows; int cols; float* hi; float* lo; int ld_dst; };
int tc_split_planes_multi(const SplitJob* jobs, int n, cudaStream_t st);

// true when the tensor-core kernel handles an [M_out x N_out] GEMM with reduction length K_red
bool tc_shape_ok(int M_out, int N_out, int K_red);
size_t tc_workspace_bytes(int M, int N, int K, int mode);

int tc_linear_fwd(const float* x, int ldx, const float* W, int ldw, const float* b, float* y, int ldy,
                  int M, int K, int N, float slope, int mode, void* ws, size_t ws_bytes, int sm_count,
                  cudaStream_t st);
int tc_linear_bwd_data(const float* dy, int lddy, const float* W, int ldw, const float* x_act, int ldxa,
                       float slope_prev, float* dx, int lddx, int M, int K, int N, int mode, void* ws,
                       size_t ws_bytes, int sm_count, cudaStream_t st);
int tc_linear_bwd_weight(const float* dy, int lddy, const float* x, int ldx, float* dW, int lddw, float* db,
                         int M, int K, int N, int mode, void* ws, size_t ws_bytes, int sm_count,
                         cudaStream_t st);

}  // namespace clica
