// lpnce_inst.cuh -- included by lpnce_p{0..4}.cu with CLICA_P defined: instantiates the fused loss
// kernels of one exponent for every (feature width, feature split) combination.
#include "lpnce_kernels.cuh"

#define CLICA_CAT_(a, b) a##b
#define CLICA_CAT(a, b) CLICA_CAT_(a, b)

namespace clica {
int CLICA_CAT(launch_fwd_p, CLICA_P)(int DP, int F, const FwdParams& q, dim3 g, cudaStream_t s, int r4) {
    CLICA_DISPATCH_DPF(CLICA_P, DP, F, launch_fwd_pd, q, g, s, r4)
}
int CLICA_CAT(launch_bwd_p, CLICA_P)(int DP, int F, const BwdParams& q, dim3 g, cudaStream_t s) {
    CLICA_DISPATCH_DPF(CLICA_P, DP, F, launch_bwd_pd, q, g, s)
}
int CLICA_CAT(occ_fwd_p, CLICA_P)(int DP, int F, int r4) {
    CLICA_DISPATCH_DPF(CLICA_P, DP, F, occ_fwd_pd, r4)
}
int CLICA_CAT(occ_bwd_p, CLICA_P)(int DP, int F) {
    CLICA_DISPATCH_DPF(CLICA_P, DP, F, occ_bwd_pd)
}
}  // namespace clica
