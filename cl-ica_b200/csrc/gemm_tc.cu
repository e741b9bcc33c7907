// gemm_tc.cu -- tcgen05 TF32 / 3xTF32 GEMM path (under construction: reports no supported shape yet).
#include "gemm_tc.cuh"

namespace clica {
bool tc_shape_ok(int, int, int) { return false; }
size_t tc_workspace_bytes(int, int, int, int) { return 0; }
int tc_linear_fwd(const float*, int, const float*, int, const float*, float*, int, int, int, int, float, int, void*,
                  size_t, int, cudaStream_t) { return fail(CLICA_E_UNSUPPORTED, "tensor-core path not built"); }
int tc_linear_bwd_data(const float*, int, const float*, int, const float*, int, float, float*, int, int, int, int,
                       int, void*, size_t, int, cudaStream_t) { return fail(CLICA_E_UNSUPPORTED, "tensor-core path not built"); }
int tc_linear_bwd_weight(const float*, int, const float*, int, float*, int, float*, int, int, int, int, void*, size_t,
                         int, cudaStream_t) { return fail(CLICA_E_UNSUPPORTED, "tensor-core path not built"); }
}  // namespace clica
