// Production build: forces, atmosphere, trigonometry and rtd in float; the 11-state, the
// angle differences and the RBF neighbour search stay in double.
#include "pd_kernels.cuh"
#ifndef PD_FP32_RBF_T
#define PD_FP32_RBF_T double
#endif
namespace pd {
static const Impl k_impl = {
    impl_reset<float, PD_FP32_RBF_T>, Launch<float, PD_FP32_RBF_T>::step,
    rollout_fp32, impl_get_state, impl_set_state, impl_transpose, impl_observe<float>};
const Impl *impl_fp32() { return &k_impl; }
}  // namespace pd
