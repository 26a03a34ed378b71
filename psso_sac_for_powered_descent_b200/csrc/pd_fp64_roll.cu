// Parity build, whole-episode rollout kernels (their own translation unit: build time).
#include "pd_kernels.cuh"
namespace pd {
int rollout_fp64(const LaunchCtx &lc, int policy, int phase, int rtd, int wind, const RolloutIO &io,
                 const WindCtx &wc, const double *sig, int *status, cudaStream_t st) {
    return Launch<double, double>::rollout(lc, policy, phase, rtd, wind, io, wc, sig, status, st);
}
}  // namespace pd
