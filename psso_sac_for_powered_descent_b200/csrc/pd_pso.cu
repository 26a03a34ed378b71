// pd_pso.cu - device-resident particle-swarm bookkeeping.
//
// Reference: ParticleSubswarmOptimisation.run (particle_swarm_optimisation.py:413-515):
//     per sub-swarm, in particle order:
//         if fitness < particle.best_fitness:  best_fitness, best_position = fitness, position
//         if fitness < subswarm_best_fitness:  subswarm_best_fitness, subswarm_best_position = ...
//     per-generation metrics (best / avg / min / max / std / count per sub-swarm, :455-470)
//     global best = sequential scan of the sub-swarm bests (:474-477)
//     v = w v + c1 r1 (best_position - x) + c2 r2 (subswarm_best - x)   (ONE r1, r2 per particle, :517-521)
//     x = clip(x + v, lo, hi)                                            (:112-118)
// A generation runs without a single host synchronisation: seed-mean -> (all-gather) -> select
// (arg-min + metrics) -> gather candidate rows -> (all-reduce = broadcast from the unknown owner)
// -> apply -> update.  The arithmetic of the update is written with explicit round-to-nearest
// intrinsics in NumPy's evaluation order, so that the host drop-in (pso.ParticleSubswarmOptimisation
// with rng='philox') and the device swarm follow bit-identical trajectories.
#include <cuda_runtime.h>
#include <math.h>
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

// One warp per particle row, lanes stride the parameter vector (coalesced fp64 rows); the fp32
// weight matrix the next rollout consumes is written in the same pass.
__global__ void __launch_bounds__(256)
pso_update_kernel(PsoUpdateArgs a) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= a.n) return;
    const int sw = a.swarm_of[warp];
    if (sw < 0) return;                       // slot not in use (after re-initialisation)
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
    const double c1r1 = __dmul_rn(a.c1, r1), c2r2 = __dmul_rn(a.c2, r2);
    const double *lbest = a.swarm_best + (size_t)sw * a.P;
    for (int j = lane; j < a.P; j += 32) {
        const double x = a.x[row + j];
        double pb = a.best[row + j];
        if (improved) { pb = x; a.best[row + j] = x; }
        // (w v + (c1 r1) (pb - x)) + (c2 r2) (lb - x): no fused multiply-add, NumPy's order
        const double t1 = __dmul_rn(a.w, a.v[row + j]);
        const double t2 = __dmul_rn(c1r1, __dsub_rn(pb, x));
        const double t3 = __dmul_rn(c2r2, __dsub_rn(lbest[j], x));
        const double v = __dadd_rn(__dadd_rn(t1, t2), t3);
        double xn = __dadd_rn(x, v);
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

__global__ void pso_seed_mean_kernel(const double *__restrict__ fit, int n, int n_seeds, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < n_seeds; ++k) s = __dadd_rn(s, fit[(size_t)i * n_seeds + k]);
    out[i] = n_seeds == 1 ? s : s / (double)n_seeds;
}

int pso_seed_mean_launch(const double *fit, int n, int n_seeds, double *out, cudaStream_t st) {
    pso_seed_mean_kernel<<<(n + 255) / 256, 256, 0, st>>>(fit, n, n_seeds, out);
    return cudaGetLastError() != cudaSuccess;
}

// block k < S: sub-swarm k; block S: the whole swarm (metrics only)
__global__ void __launch_bounds__(1024)
pso_select_kernel(PsoSelectArgs a) {
    const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const bool all = k == a.S;
    __shared__ double s_min[1024], s_max[1024], s_sum[1024];
    __shared__ int s_idx[1024], s_cnt[1024];
    double mn = INFINITY, mx = -INFINITY, sum = 0.0;
    int idx = 0x7fffffff, cnt = 0;
    for (int i = tid; i < a.N; i += nt) {
        const int sw = a.swarm_of_all[i];
        if (sw < 0 || (!all && sw != k)) continue;
        const double f = a.allfit[i];
        if (f < mn) { mn = f; idx = i; }      // ascending i per thread: keeps the first occurrence
        mx = fmax(mx, f);
        sum += f;
        ++cnt;
    }
    s_min[tid] = mn; s_idx[tid] = idx; s_max[tid] = mx; s_sum[tid] = sum; s_cnt[tid] = cnt;
    __syncthreads();
    for (int off = nt >> 1; off > 0; off >>= 1) {
        if (tid < off) {
            const double m2 = s_min[tid + off];
            const int i2 = s_idx[tid + off];
            if (m2 < s_min[tid] || (m2 == s_min[tid] && i2 < s_idx[tid])) { s_min[tid] = m2; s_idx[tid] = i2; }
            s_max[tid] = fmax(s_max[tid], s_max[tid + off]);
            s_sum[tid] += s_sum[tid + off];
            s_cnt[tid] += s_cnt[tid + off];
        }
        __syncthreads();
    }
    const int count = s_cnt[0];
    const double mean = count ? s_sum[0] / count : 0.0;
    const double bmin = s_min[0], bmax = s_max[0];
    const int bidx = s_idx[0];
    __syncthreads();
    // population standard deviation (np.std), second pass about the mean
    double ss = 0.0;
    for (int i = tid; i < a.N; i += nt) {
        const int sw = a.swarm_of_all[i];
        if (sw < 0 || (!all && sw != k)) continue;
        const double d = a.allfit[i] - mean;
        ss += d * d;
    }
    s_sum[tid] = ss;
    __syncthreads();
    for (int off = nt >> 1; off > 0; off >>= 1) {
        if (tid < off) s_sum[tid] += s_sum[tid + off];
        __syncthreads();
    }
    if (tid == 0) {
        double best = all ? 0.0 : a.swarm_best_fit[k];
        if (!all) {
            const int imp = count > 0 && bmin < best;
            if (imp) { best = bmin; a.swarm_best_fit[k] = bmin; }
            a.improved[k] = imp;
            a.sel_idx[k] = count > 0 ? bidx : -1;
        }
        double *row = a.stats + (size_t)k * 6;
        row[0] = best; row[1] = mean; row[2] = bmin; row[3] = bmax;
        row[4] = count ? sqrt(s_sum[0] / count) : 0.0;
        row[5] = (double)count;
    }
}

int pso_select_launch(const PsoSelectArgs &a, cudaStream_t st) {
    pso_select_kernel<<<a.S + 1, 1024, 0, st>>>(a);
    return cudaGetLastError() != cudaSuccess;
}

__global__ void pso_gather_kernel(const double *__restrict__ x, long long lo, int n_local, int P,
                                  const int *__restrict__ sel_idx, const int *__restrict__ improved,
                                  double *__restrict__ cand) {
    const int k = blockIdx.x;
    const long long j = (long long)sel_idx[k] - lo;
    const bool mine = improved[k] && sel_idx[k] >= 0 && j >= 0 && j < n_local;
    for (int p = threadIdx.x; p < P; p += blockDim.x)
        cand[(size_t)k * P + p] = mine ? x[(size_t)j * P + p] : 0.0;
}

int pso_gather_launch(const double *x, long long lo, int n_local, int P, const int *sel_idx, const int *improved,
                      int S, double *cand, cudaStream_t st) {
    pso_gather_kernel<<<S, 128, 0, st>>>(x, lo, n_local, P, sel_idx, improved, cand);
    return cudaGetLastError() != cudaSuccess;
}

__global__ void pso_apply_kernel(const double *__restrict__ cand, const int *__restrict__ improved, int S, int P,
                                 double *__restrict__ swarm_best, const double *__restrict__ swarm_best_fit,
                                 double *__restrict__ gbest_pos, double *__restrict__ gbest_fit,
                                 double *__restrict__ hist_row) {
    for (int k = 0; k < S; ++k) {
        if (!improved[k]) continue;
        for (int p = threadIdx.x; p < P; p += blockDim.x) swarm_best[(size_t)k * P + p] = cand[(size_t)k * P + p];
    }
    __syncthreads();
    // for i, fitness in enumerate(subswarm_best_fitnesses): if fitness < global_best: take it (:474-477)
    double g = *gbest_fit;
    __syncthreads();
    for (int k = 0; k < S; ++k) {
        const double f = swarm_best_fit[k];
        if (f < g) {
            g = f;
            for (int p = threadIdx.x; p < P; p += blockDim.x) gbest_pos[p] = swarm_best[(size_t)k * P + p];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *gbest_fit = g;
        if (hist_row) hist_row[0] = g;
    }
}

int pso_apply_launch(const double *cand, const int *improved, int S, int P, double *swarm_best,
                     const double *swarm_best_fit, double *gbest_pos, double *gbest_fit, double *hist_row,
                     cudaStream_t st) {
    pso_apply_kernel<<<1, 128, 0, st>>>(cand, improved, S, P, swarm_best, swarm_best_fit, gbest_pos, gbest_fit, hist_row);
    return cudaGetLastError() != cudaSuccess;
}

}  // namespace pd
