// pd_impl.h - plain structs shared by the precision TUs (pd_fp64.cu / pd_fp32.cu) and the
// C-ABI front end (pd_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pd {

template <typename R> struct Scalars;
struct Tables;
struct WindCtx;
struct KParams;

// what a launch needs besides its data: the handle's constant block (passed to every kernel as a
// __grid_constant__ parameter), the SM count of the handle's device and the device ordinal
struct LaunchCtx {
    const KParams *kp;
    int n_sm;
    int device;
};

// Per-env persistent data, field-major (SoA) so that a warp touches 32 consecutive words.
struct EnvSoA {
    int n;
    double *st;            // [11][n] state
    double *gwin;          // [10][n] g-load FIFO (oldest first)
    int *gwin_n;           // [n]
    double *aprev;         // [3][n]  gimbal_deg_prev, delta_left_prev, delta_right_prev
    double *wst;           // [6][n]  gust filter states + sigma_u, sigma_v
    unsigned int *wctr;    // [n]     noise draws consumed this episode
    unsigned int *episode; // [n]     episode counter (Philox stream id for sigma draws)
    int *trunc_id;         // [n]
    int *ep_steps;         // [n]
    int *status;           // [1] sticky error bits (PD_RBF_MISS ...)
};

struct StepIO {
    const void *actions;
    int action_dtype;
    int raw_actions;       // skip the RL wrapper's augment_action
    int dbg_full;          // dbg rows are PD_INFO_DIM wide (fp64 build, no wind)
    int supervisory;       // type = 'supervisory': replace the rl closures' verdict (rtd_supervisory_mock.py)
    void *obs, *reward, *next_obs;
    uint8_t *done, *truncated;
    int32_t *trunc_id;
    double *dbg;
};

struct RolloutIO {
    int n_episodes, n_seeds, max_steps;
    unsigned int generation;   // Philox episode word of the gust noise (fresh draws every generation)
    const float *wT;       // [n_params][w_stride]
    size_t w_stride;
    const void *actions;   // tape [max_steps][n_episodes][A]
    int action_dtype;
    double *ret;
    int32_t *steps, *trunc_id;
    double *terminal, *traj, *rewards;
    float *act_out;        // [max_steps][n_episodes][A] actions applied (MLP policy)
    int *queue;            // work-queue head (next episode index to hand out)
    // Straggler hand-off (per-particle MLP policy): an episode still running after handoff_steps
    // (then handoff2_steps) is appended with its complete state to continuation records and
    // finished by a later stage with more lanes per episode (1 -> 8 -> 32): a generation is
    // otherwise bounded by the sequential latency of its longest episode.
    int handoff_steps;     // 0 = off
    int handoff2_steps;    // second boundary (<= handoff_steps: 2 x handoff_steps); later ones double
    int lanes8_below, lanes32_below;   // survivor counts at or below which a stage uses 8 / 32 lanes (0 = default)
    int run_if_gt, run_if_le;   // record-fed stages run only if the input record count is in (gt, le]
    double *cont_d;        // input records [PD_CONT_D][cont_cap]
    int *cont_i;           // [PD_CONT_I][cont_cap]: episode, t, g-window count, wind draw counter
    int *cont_count;       // [1]
    int cont_cap;
    double *out_d;         // output records of a hand-off stage
    int *out_i;
    int *out_count;
    int out_cap;
};
#define PD_CONT_D 32
#define PD_CONT_I 4

struct Impl {
    void (*reset)(const LaunchCtx &, const EnvSoA &, const uint8_t *, const WindCtx &, const double *, cudaStream_t);
    void (*step)(const LaunchCtx &, int phase, int rtd, int wind, const EnvSoA &, const StepIO &, const WindCtx &,
                 const double *, int auto_reset, cudaStream_t);
    int (*rollout)(const LaunchCtx &, int policy, int phase, int rtd, int wind, const RolloutIO &, const WindCtx &,
                   const double *, int *status, cudaStream_t);
    void (*get_state)(const EnvSoA &, double *, double *, int *, double *, cudaStream_t);
    void (*set_state)(const EnvSoA &, const double *, const double *, const int *, const double *,
                      cudaStream_t);
    void (*transpose)(const float *, float *, int, int, cudaStream_t);
    void (*observe)(const LaunchCtx &, int phase, int rtd, const EnvSoA &, void *obs, cudaStream_t);
};

const Impl *impl_fp64();
const Impl *impl_fp32();

// the rollout kernels of each precision are compiled in their own translation unit (build time)
int rollout_fp64(const LaunchCtx &, int policy, int phase, int rtd, int wind, const RolloutIO &, const WindCtx &,
                 const double *, int *status, cudaStream_t);
int rollout_fp32(const LaunchCtx &, int policy, int phase, int rtd, int wind, const RolloutIO &, const WindCtx &,
                 const double *, int *status, cudaStream_t);

}  // namespace pd
