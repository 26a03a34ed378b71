#pragma once
namespace pd {
// TFLOP/s of a pure FMA kernel (2 FLOP per FMA) on the current device; ms = its duration
int measure_fma_peak(int fp64, int n_sm, double *tflops, double *ms);
}  // namespace pd
