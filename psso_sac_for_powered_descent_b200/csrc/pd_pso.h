#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace pd {
struct PsoUpdateArgs {
    double *x, *v, *best, *best_fit;      // [n*P], [n*P], [n*P], [n]   (this rank's particles)
    const double *fitness;                // [n] fitness of x
    const int *swarm_of;                  // [n] sub-swarm id (< 0: slot not in use)
    const double *swarm_best;             // [S*P] sub-swarm best positions (already updated)
    float *weights_out;                   // [n*P] or null: float32 copy of the new positions
    int n, P;
    long long index0;                     // global index of this rank's first particle
    double w, c1, c2, lo, hi;
    unsigned long long seed;
    int generation;
};
int pso_update_launch(const PsoUpdateArgs &a, cudaStream_t st);

// fitness[e] of n * n_seeds episodes -> mean over the seeds of each particle
int pso_seed_mean_launch(const double *fit, int n, int n_seeds, double *out, cudaStream_t st);

// Per sub-swarm: arg-min (first occurrence) of this generation's fitness, the reference's
// per-generation metrics, and the sub-swarm best bookkeeping.
struct PsoSelectArgs {
    const double *allfit;                 // [N] fitness of every particle (all ranks' slices)
    const int *swarm_of_all;              // [N] sub-swarm id, < 0 = slot not in use
    int N, S;
    double *swarm_best_fit;               // [S] in/out
    int *sel_idx;                         // [S] global index of the sub-swarm's best particle of this generation
    int *improved;                        // [S] 1 if it beat swarm_best_fit
    double *stats;                        // [(S + 1) * 6] per sub-swarm: best so far, avg, min, max, std, count;
                                          // row S: avg / min / max / std / count over the whole swarm
};
int pso_select_launch(const PsoSelectArgs &a, cudaStream_t st);

// cand[k] = x[sel_idx[k] - lo] if this rank owns that particle and improved[k], else 0
int pso_gather_launch(const double *x, long long lo, int n_local, int P, const int *sel_idx, const int *improved,
                      int S, double *cand, cudaStream_t st);

// swarm_best[k] <- cand[k] where improved[k]; then the global best (sequential scan over the
// sub-swarms as upstream); hist_row = {global best fitness after this generation}
int pso_apply_launch(const double *cand, const int *improved, int S, int P, double *swarm_best,
                     const double *swarm_best_fit, double *gbest_pos, double *gbest_fit, double *hist_row,
                     cudaStream_t st);
}  // namespace pd
