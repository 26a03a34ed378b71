#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace pd {
struct PsoUpdateArgs {
    double *x, *v, *best, *best_fit;      // [n*P], [n*P], [n*P], [n]   (this rank's particles)
    const double *fitness;                // [n] fitness of x
    const int *swarm_of;                  // [n] sub-swarm id
    const double *swarm_best;             // [S*P] sub-swarm best positions (already updated)
    float *weights_out;                   // [n*P] or null: float32 copy of the new positions
    int n, P;
    long long index0;                     // global index of this rank's first particle
    double w, c1, c2, lo, hi;
    unsigned long long seed;
    int generation;
};
int pso_update_launch(const PsoUpdateArgs &a, cudaStream_t st);
}  // namespace pd
