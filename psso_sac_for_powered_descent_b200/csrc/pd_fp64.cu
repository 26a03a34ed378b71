// Parity build: every arithmetic step in double.
#include "pd_kernels.cuh"
namespace pd {
static const Impl k_impl = {
    impl_reset<double, double>, Launch<double, double>::step,
    rollout_fp64, impl_get_state, impl_set_state, impl_transpose, impl_observe<double>};
const Impl *impl_fp64() { return &k_impl; }
}  // namespace pd
