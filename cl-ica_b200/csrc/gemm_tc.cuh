// gemm_tc.cuh -- interface of the tcgen05 (5th-gen tensor core) TF32 / 3xTF32 GEMM path (gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace clica {

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
