// pd_peak.cu - measured FFMA / DFMA pipe peaks of the device the bench runs on.
//
// bench.py's fp_roofline needs a denominator that is a measurement, not a data-sheet product
// ("the build must measure them").  One kernel per precision: every thread runs 8
// independent fused-multiply-add chains (enough instruction-level parallelism to cover the pipe
// latency at 64 resident warps per SM), 2 FLOP per FMA, timed with CUDA events on the device.
#include <cuda_runtime.h>

#include "pd_peak.h"

namespace pd {

template <typename T>
__global__ void __launch_bounds__(1024)
fma_peak_kernel(T *out, int iters, T b, T c) {
    T a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (T)(threadIdx.x + j) * (T)1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = a[j] * b + c;      // contracts to one FFMA / DFMA
        }
    }
    T s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T>
static int measure(int n_sm, int iters, double *tflops, double *ms_out) {
    const int blocks = n_sm * 2, threads = 1024;       // 2 x 1024 threads = 64 warps per SM
    T *out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(T)) != cudaSuccess) return 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {                // first launch = warm-up
        cudaEventRecord(e0, 0);
        fma_peak_kernel<T><<<blocks, threads>>>(out, iters, (T)0.999999, (T)1e-6);
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return 1; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    const double flop = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = flop / ((double)best * 1e-3) / 1e12;
    *ms_out = best;
    return cudaGetLastError() != cudaSuccess;
}

int measure_fma_peak(int fp64, int n_sm, double *tflops, double *ms) {
    return fp64 ? measure<double>(n_sm, 2048, tflops, ms) : measure<float>(n_sm, 4096, tflops, ms);
}

}  // namespace pd
