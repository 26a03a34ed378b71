// pd_pso.cu - device-resident particle-swarm update.
//
// Reference: update_velocity_with_local_best + update_position + the personal-best bookkeeping
// of ParticleSubswarmOptimisation.run (particle_swarm_optimisation.py:425-490, 517-521, 112-118):
//     if fitness < best_fitness: best_fitness, best_position = fitness, position
//     v = w v + c1 r1 (best_position - x) + c2 r2 (subswarm_best - x)   (ONE r1, r2 per particle)
//     x = clip(x + v, lo, hi)
// One warp per particle row, lanes stride the parameter vector (coalesced fp64 rows); the fp32
// weight matrix the next rollout consumes is written in the same pass.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pd_pso.h"

namespace pd {

__device__ __forceinline__ void philox_pso(unsigned int c0, unsigned int c1, unsigned int k0,
                                           unsigned int k1, unsigned int out[4]) {
    unsigned int c2 = 0x50534F55u, c3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned int n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(256)
pso_update_kernel(PsoUpdateArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.n) return;
    const size_t row = (size_t)warp * a.P;
    const double f = a.fitness[warp];
    const bool improved = f < a.best_fit[warp];
    unsigned int r[4];
    // the stream is keyed by the GLOBAL particle index so that the result does not depend on
    // how the swarm is sharded over ranks
    philox_pso((unsigned int)(a.index0 + warp), (unsigned int)a.generation, (unsigned int)a.seed,
               (unsigned int)(a.seed >> 32), r);
    const double r1 = ((((unsigned long long)r[0] << 32) | r[1]) >> 11) * (1.0 / 9007199254740992.0);
    const double r2 = ((((unsigned long long)r[2] << 32) | r[3]) >> 11) * (1.0 / 9007199254740992.0);
    const double *lbest = a.swarm_best + (size_t)a.swarm_of[warp] * a.P;
    for (int j = lane; j < a.P; j += 32) {
        const double x = a.x[row + j];
        double pb = a.best[row + j];
        if (improved) { pb = x; a.best[row + j] = x; }
        double v = a.w * a.v[row + j] + a.c1 * r1 * (pb - x) + a.c2 * r2 * (lbest[j] - x);
        double xn = x + v;
        xn = xn < a.lo ? a.lo : (xn > a.hi ? a.hi : xn);
        a.v[row + j] = v;
        a.x[row + j] = xn;
        if (a.weights_out) a.weights_out[row + j] = (float)xn;
    }
    if (improved && lane == 0) a.best_fit[warp] = f;
}

int pso_update_launch(const PsoUpdateArgs &a, cudaStream_t st) {
    const int threads = 256;
    const long long lanes = (long long)a.n * 32;
    pso_update_kernel<<<(int)((lanes + threads - 1) / threads), threads, 0, st>>>(a);
    return cudaGetLastError() != cudaSuccess;
}

}  // namespace pd
