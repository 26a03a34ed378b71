// Production build, whole-episode rollout kernels (their own translation unit: build time).
#include "pd_kernels.cuh"
#ifndef PD_FP32_RBF_T
#define PD_FP32_RBF_T double
#endif
namespace pd {
int rollout_fp32(const LaunchCtx &lc, int policy, int phase, int rtd, int wind, const RolloutIO &io,
                 const WindCtx &wc, const double *sig, int *status, cudaStream_t st) {
    return Launch<float, PD_FP32_RBF_T>::rollout(lc, policy, phase, rtd, wind, io, wc, sig, status, st);
}
}  // namespace pd
