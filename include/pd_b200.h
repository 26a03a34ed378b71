/*
 * pd_b200.h - C ABI of the B200-native batched powered-descent hot path.
 *
 * The reference (JvanZyl1/PSSO-SAC-for-powered-descent) is pure Python and has no FFI;
 * its seam for this path is duck-typed Python.  Each entry point
 * below names the reference interface it replaces (paths relative to the reference
 * root).  The Python host layer (psso_sac_for_powered_descent_b200/envs.py, pso.py)
 * binds these with ctypes and presents the reference's own class / method names.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; pd_last_error() gives text;
 *   - pointers marked "dev" are caller-owned CUDA device memory (e.g. torch tensors);
 *     pointers marked "host" are read during the call only;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work
 *     is asynchronous on that stream;
 *   - one PdEnv per GPU / per batch, not thread-safe per handle;
 *   - there is no CPU fallback: without a CUDA device pd_create fails.
 *
 * State layout: 11 doubles per env, [x, y, vx, vy, theta, theta_dot, gamma, alpha, mass,
 * mass_propellant, time] (src/envs/rockets_physics.py:475,646).  Inside the handle the
 * batch is stored SoA (field-major) in fp64 for both precisions; the AoS views of
 * pd_get_state/pd_set_state exist for parity tests and checkpointing.
 */
#ifndef PD_B200_H
#define PD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PD_STATE_DIM 11
#define PD_MAX_LEVELS 5
#define PD_RBF_NEIGHBOURS 50
#define PD_RBF_ROW_BYTES 512
#define PD_DBG_DIM 16
#define PD_INFO_DIM 48     /* PD_DBG_DIM + 32 (pd_set_info_mode) */

/* flight_phase strings (src/envs/base_environment.py:21).  0/1 work with type 'pso' and 'rl';
 * 2..5 with type 'rl' only, as upstream (their pso closures have the wrong arity,
 * src/envs/pso/rtd_pso.py:38-157).  'flip_over_boostbackburn' (TypeError in rtd_rl.py:132) and
 * 'landing_burn_ACS' (broken in compile_physics, rockets_physics.py:867-889) do not run upstream
 * and are not offered. */
enum { PD_PHASE_PURE_THROTTLE = 0, PD_PHASE_GIMBALLED = 1, PD_PHASE_SUBSONIC = 2,
       PD_PHASE_SUPERSONIC = 3, PD_PHASE_BALLISTIC_ARC = 4, PD_PHASE_PCONTROL = 5,
       PD_PHASE_FLIP_OVER = 6,   /* flip_over_boostbackburn: type 'supervisory' only, as upstream */
       PD_N_PHASES = 7 };
/* type = 'pso' | 'rl' | 'supervisory' (src/envs/supervisory/rtd_supervisory_mock.py: done /
 * truncation per phase, reward 0; pd_step only, actions as the base env takes them) */
enum { PD_RTD_PSO = 0, PD_RTD_RL = 1, PD_RTD_SUPERVISORY = 2 };
enum { PD_FP64 = 0, PD_FP32 = 1 };                             /* compute precision build */
enum { PD_ACT_F64 = 0, PD_ACT_F32 = 1 };                       /* dtype of the action     */
enum { PD_POLICY_MLP = 0, PD_POLICY_TAPE = 1, PD_POLICY_CLASSICAL = 2 };

/* Uniform (Mach, AoA) lookup grid over one query region of a local RBF table.
 * cells[ia*nm+im] >= 0 : set id valid for the whole grid cell;
 * cells < 0            : -(k+1) indexes (imp_hint, imp_id).  imp_hint bit 63 set: the cell is cut
 *                        by exactly one Voronoi edge; imp_id = first set, bits 16..31 = second set,
 *                        bits 8..15 / 0..7 = point slots p / q; the query belongs to the first set iff
 *                        it is not farther from p than from q.  Bit 63 clear: imp_hint = packed
 *                        per-level [lo, hi) of the candidate set imp_id to walk from. */
typedef struct {
    double m0, dm, a0, da;
    int32_t nm, na;
    int32_t n_impure, _pad;
    const int32_t *cells;                 /* host [nm*na] */
    const uint64_t *imp_hint;             /* host [n_impure] */
    const int32_t *imp_id;                /* host [n_impure] */
} PdRbfGrid;

/* Local thin-plate-spline table (host pointers, copied by pd_create);
 * built by psso_sac_for_powered_descent_b200/rbf_sets.py.  Replaces the per-call
 * scipy RBFInterpolator(neighbors=50) of src/envs/utils/aerodynamic_coefficients.py:57-66. */
typedef struct {
    int32_t n_levels;
    int32_t n_points;
    int32_t n_sets;
    int32_t hash_size;                    /* power of two */
    double levels[PD_MAX_LEVELS];         /* AoA coordinate of each level */
    int32_t level_off[PD_MAX_LEVELS + 1];
    int32_t n_grids;
    const double *mach_sorted;            /* host [n_points] level-major, Mach ascending */
    const double *points;                 /* host [n_points*2] (Mach, AoA) per slot */
    const uint8_t *rows;                  /* host [n_sets*512]: 57 doubles + 50 index bytes */
    const uint64_t *hash_keys;            /* host [hash_size] */
    const int32_t *hash_vals;             /* host [hash_size] */
    PdRbfGrid grids[2];
} PdRbfTable;

/* Constants of the flight phases outside the two landing burns (host pointers, copied by
 * pd_create): force_moment_decomposer_ascent (rockets_physics.py:17-56, 727-752), RCS
 * (:149-166, 782-801), full_rocket_inertia closure cells
 * (src/RocketSizing/functions/rocket_dimensions.py:199-241), per-phase initial states
 * (src/envs/load_initial_states.py:5-54) and normalisation vectors
 * (src/envs/utils/input_normalisation.py:5-71), the ascent reference trajectory
 * (src/envs/utils/reference_trajectory_interpolation.py:5-35). */
typedef struct {
    int32_t n_engines_stage1;             /* 'Number of engines stage 1' */
    int32_t n_ref;                        /* reference-trajectory rows */
    double max_rcs_force_per_thruster, d_base_rcs_bottom, d_base_rcs_top;
    /* m_s_1, x_dry_1, I_dry_1, m_2, m_pay, x_wet_2_initial, I_wet_2_initial, h_1, h_1_ox, h_1_f,
     * m_1_ox, m_1_f, h_lower_1 */
    double inertia_full[13];
    double engine_height_full, cop_full;
    double initial_state[4][PD_STATE_DIM]; /* subsonic, supersonic, ballistic_arc_descent, flip_over_boostbackburn */
    double norm_vals[4][8];               /* same order; ballistic uses the first 4, flip-over the first 2 */
    const double *ref_y, *ref_x, *ref_vx, *ref_vy;   /* host [n_ref], raw csv order */
    double ref_terminal[5];               /* last row: x, y, vx, vy, mass */
} PdOtherPhases;

/* Constants the reference loads in compile_physics (src/envs/rockets_physics.py:707-957),
 * module import side effects (rockets_physics.py:12-14, acs_model.py:10-11),
 * load_landing_burn_initial_state (src/envs/load_initial_states.py:56-62) and
 * landing_burn_input_normalisation (src/envs/utils/input_normalisation.py:73-89). */
typedef struct {
    double thrust_per_engine, nozzle_exit_pressure, nozzle_exit_area, v_exhaust;
    int32_t n_engines_gimballed;
    int32_t _pad0;
    double grid_fin_area, d_base_grid_fin, rocket_radius, frontal_area;
    double propellant_mass_stage1;        /* kg */
    double c_gust_x, c_gust_y;
    /* stage_inertia closure cells: I_dry, h_f, h_lower, h_ox, m_dry, m_f, m_ox, x_dry */
    double inertia[8];
    double engine_height, cop;
    double initial_state[PD_STATE_DIM];
    double norm_vals[7];
    double v_opt_a, v_opt_b;              /* classical controller velocity profile */
    /* grid-fin tables sorted ascending in Mach (scipy interp1d sorts them) */
    int32_t n_gf_ca, n_gf_cn;
    const double *gf_ca_mach, *gf_ca_val; /* host */
    const double *gf_cn_mach, *gf_cn_val; /* host */
    /* wind: altitude profile (km, m/s; sorted) and the unit-sigma discretised gust filters */
    int32_t n_wind, _pad1;
    const double *wind_alt_km, *wind_speed;   /* host */
    double vk_Adu[4], vk_Bdu[2], vk_Adv[4], vk_Bdv[2];
    PdRbfTable cd, cl;
    const PdOtherPhases *other;           /* host; required for PD_PHASE_SUBSONIC .. PD_PHASE_BALLISTIC_ARC and PD_PHASE_FLIP_OVER */
} PdParams;

/* rocket_environment_pre_wrap.__init__ kwargs (src/envs/base_environment.py:12-20) + batch */
typedef struct {
    int32_t phase;            /* PD_PHASE_* */
    int32_t rtd;              /* PD_RTD_*   */
    int32_t precision;        /* PD_FP64 | PD_FP32 */
    int32_t enable_wind;
    int32_t stochastic_wind;
    int32_t auto_reset;       /* reset an env in the step that ends its episode */
    int32_t n_envs;
    int32_t device;
    uint64_t seed;            /* Philox key for gust noise / sigma draws */
    double rl_reward_scale;   /* (1-g)/(1-g^L) of rtd_rl.py:266 (phase G, rl only) */
    double discount_factor;   /* rtd_rl.py:496 ALIVE_BONUS = 0.01 (1-g) (P-control phase, rl only) */
    int32_t raw_actions;      /* type 'rl' only.  0: pd_step takes the policy's action and applies
                                 rl_wrapped_env_pytorch.augment_action (log-compression for
                                 landing_burn, reference-speed scaling for P-control) in the kernel;
                                 1: actions are already what rocket_environment_pre_wrap.step expects */
    int32_t exact_aero;       /* PD_FP32 only.  0: C_L / C_D from the bicubic patches of the thin-plate sums
                                 (1e-8 absolute, csrc/pd_patch.h; queries on a rejected patch or a walk cell
                                 take the exact sum); 1: the exact 50-term sums everywhere, as PD_FP64 */
} PdConfig;

typedef struct PdEnv PdEnv;

const char *pd_last_error(void);
int pd_version(void);

/* rocket_environment_pre_wrap(...)  (base_environment.py:12-78) for a batch of n_envs */
int pd_create(const PdConfig *cfg, const PdParams *params, PdEnv **out);
int pd_destroy(PdEnv *env);

/* .reset()  (base_environment.py:80-97).  mask: dev uint8[n_envs] or NULL = all. */
int pd_reset(PdEnv *env, const uint8_t *mask, void *stream);

/* .step(actions)  (base_environment.py:99-154) for every env of the batch.
 *   actions   dev [n_envs * A], double or float per action_dtype (A = 1, 4, 2, 2, 1, 1 for
 *             phases 0..5).  A float32
 *             action reproduces NumPy's NEP-50 float32 contamination of throttle, thrust and
 *             mass flow in the fp64 build (SURVEY.md 8a "dtype rule").
 *   obs       dev [n_envs * O] pso: pso_wrapper.augment_state (env_wrapped_ea.py:97-123);
 *             rl: rl_wrapped_env_pytorch._process_state/augment_state
 *             (env_wrapped_rl_pytorch.py:42-47,167-202).  Element type = double (PD_FP64) or
 *             float (PD_FP32).  Observation of the post-step, pre-reset state.
 *   reward    dev [n_envs] same element type as obs
 *   done, truncated  dev uint8[n_envs];  trunc_id dev int32[n_envs] (env.truncation_id)
 *   next_obs  dev or NULL: observation after the auto-reset (== obs where no reset happened)
 *   dbg       dev double[n_envs * PD_DBG_DIM] or NULL: last sub-step's info values
 *             (mach, q, CL, CD, rho, p_atm, a, x_cog, inertia, mass_flow, throttle,
 *              alpha_eff, g_load_1_sec_window, ug, vg, rbf_status)
 * "dev" pointers may also be mapped pinned host memory (cudaHostAlloc / torch pin_memory, same
 * address under UVA): each action is read once and each result stored once, coalesced, so a
 * host caller needs no staging copies (BatchedRocketEnv.step_host does exactly that). */
int pd_step(PdEnv *env, const void *actions, int action_dtype, void *obs, void *reward,
            uint8_t *done, uint8_t *truncated, int32_t *trunc_id, void *next_obs, double *dbg,
            void *stream);

/* Width of pd_step's `dbg` rows.  full = 0 (default): PD_DBG_DIM values.  full = 1 (PD_FP64 build,
 * wind disabled): PD_INFO_DIM values - the 16 above followed by the remaining primitives of the
 * reference's `info` dict (rockets_physics.py:649-702) of the last sub-step: drag, lift, d_cp_cg,
 * d_thrust_cg, fuel_percentage_consumed, control_force_parallel, control_force_perpendicular,
 * control_force_x, control_force_y, aero_force_x, aero_force_y, g, control_moment_z,
 * aero_moment_z, moments_z, theta_dot_dot, vx_dot, vy_dot, F_wind_x, gimbal_angle_deg,
 * delta_command_left_rad, delta_command_right_rad, mach_number_max, pitch angle at the start of
 * the sub-step, grid-fin C_a(M), C_n_alpha(M), filtered left / right fin deflection (the inputs of
 * acs_info, src/envs/utils/acs_model.py:62-84), 4 reserved zeros.  A separate
 * diagnostic instantiation of the step kernel; the throughput kernels are unaffected. */
int pd_set_info_mode(PdEnv *env, int full);

/* AoS state access (dev double[n_envs*11]); g_window dev double[n_envs*10] + n_window dev
 * int32[n_envs]; act_prev dev double[n_envs*3] = gimbal_angle_deg_prev,
 * delta_command_left_rad_prev, delta_command_right_rad_prev.  NULL = leave untouched /
 * do not return. */
int pd_get_state(PdEnv *env, double *state, double *g_window, int32_t *n_window,
                 double *act_prev, void *stream);
int pd_set_state(PdEnv *env, const double *state, const double *g_window,
                 const int32_t *n_window, const double *act_prev, void *stream);

/* Parity hook for the stochastic wind: replace Philox by an explicit N(0,1) tape
 * (dev double[n_envs * tape_len], consumed u-then-v per sub-step below 15 km exactly like
 * src/envs/wind/vonkarman.py:33-36) and explicit (sigma_u, sigma_v) per env
 * (dev double[n_envs*2]).  tape = NULL returns to Philox. */
int pd_set_wind_tape(PdEnv *env, const double *tape, int tape_len, const double *sigma_uv);

/* pso_wrapped_env.objective_function for a whole swarm
 * (src/envs/pso/env_wrapped_ea.py:200-222; replaces parallel_evaluate,
 * src/particle_swarm_optimisation/particle_swarm_optimisation.py:334-350).
 * One persistent kernel: reset -> [obs -> per-particle MLP -> 4 sub-steps -> rtd]* .
 *   weights   dev float[n_particles * n_params], named_parameters() order
 *             (weight row-major then bias per layer, env_wrapped_ea.py:46-59)
 *   n_seeds   episodes per particle (wind seeds); episode e = particle * n_seeds + seed
 *   max_steps step cap (the reference has none).  A capped episode reports trunc_id = -1 and is
 *             scored as a truncation at its final state (the closures' truncated-branch reward:
 *             rtd_pso.py:222-229 / 300-316), so that a stalling policy never outranks a crash
 *   fitness   dev double[n_particles*n_seeds] = -sum(reward)
 *   steps, trunc_id  dev int32[...] or NULL;  terminal_state dev double[... * 11] or NULL
 *   traj / actions_out / rewards: optional per-step trace, step-major
 *   (dev double[max_steps*E*11], float[max_steps*E*A], double[max_steps*E]) - what
 *   collect_trajectory_data (particle_swarm_optimisation.py:759-785) records. */
int pd_rollout_pso(PdEnv *env, const float *weights, int n_particles, int n_params, int n_seeds,
                   int max_steps, double *fitness, int32_t *steps, int32_t *trunc_id,
                   double *terminal_state, double *traj, float *actions_out, double *rewards,
                   void *stream);

/* PD_FP32 handles without exact_aero: counts[0..3] = bicubic patches built / rejected by the 1e-8
 * validation for C_D, then for C_L (all zero otherwise); *max_err (may be NULL) = largest validation
 * error among the patches in use. */
int pd_aero_patch_stats(PdEnv *env, int64_t *counts, double *max_err);
/* The patch sets are shared by all handles of a process (one per GPU and table, 0.34 GB) and stay
 * cached after the last handle is destroyed, so that the next pd_create does not rebuild them
 * (0.2 s).  This frees the sets no live handle uses; returns the bytes of patches released. */
int64_t pd_release_aero_patches(void);

/* Straggler hand-off of pd_rollout_pso.  Episode lengths are ragged (a random
 * landing_burn_pure_throttle swarm has a median of 130 steps, 2 % above 512 and a few episodes at
 * the cap; after 30 generations of the optimiser 20-30 % are above 512), and a generation cannot
 * end before its longest episode has, so the rollout runs as a chain of stages
 *     reset -> steps -> steps2 -> 2 steps2 -> 4 steps2 -> ... -> end:
 * an episode still running at a boundary is appended, with its complete state, to continuation
 * records, and the next stage serves the records with 1, 8 or 32 lanes per episode depending on how
 * many there are (per-step latency of a lone episode 31 / 10 / 7.1 us, instructions per
 * episode-step 625 / 2 000 / 3 000).  Defaults: 128 / 256 for landing_burn_pure_throttle; one
 * hand-off after 16 steps for landing_burn while a call has at most 1.5 x the GPU's lane count of
 * episodes.  An episode that ends inside the first stage is bit-identical to the one-pass rollout;
 * one that is handed on resumes from its exact state in another instantiation of the kernel, so
 * rounding-level differences (summation order of the cooperative RBF sums, FMA contraction; in the
 * fp32 build atan2 against its incremental form) appear from there on: 99 % of the fitness values
 * within 1e-9 relative in the fp64 build, 1e-4 in the fp32 build
 * (tests/test_gpu_parity.py::test_rollout_stage_chain_every_lane_choice).
 * steps = 0: off.  steps2 <= steps: 2 x steps. */
int pd_set_rollout_handoff(PdEnv *env, int steps);
int pd_set_rollout_stages(PdEnv *env, int steps, int steps2);
/* Survivor counts at or below which a record-fed stage uses 8 / 32 lanes per episode
 * (0 = default: lanes / 4 and lanes / 32 with lanes = SMs x 448, the measured cross-over points). */
int pd_set_rollout_lanes(PdEnv *env, int lanes8_below, int lanes32_below);

/* Whole-episode rollouts with a scripted policy, one launch:
 *   PD_POLICY_TAPE       actions dev [max_steps * n_episodes * A] (step-major), dtype per
 *                        action_dtype; env.step loop (base_environment.py:99-154)
 *   PD_POLICY_CLASSICAL  LandingBurn(test_case='control').run_closed_loop
 *                        (src/classical_controls/landing_burn_pure_throttle.py:261-339);
 *                        physics only, its own stop condition, no rtd
 * traj: dev double[max_steps * n_episodes * 11] or NULL (state after every step).
 * rewards: dev double[max_steps * n_episodes] or NULL. */
int pd_rollout_policy(PdEnv *env, int policy, const void *actions, int action_dtype,
                      int n_episodes, int max_steps, double *ret, int32_t *steps,
                      int32_t *trunc_id, double *terminal_state, double *traj, double *rewards,
                      void *stream);

/* SAC data collection with one shared actor over the env batch
 * (sac_pytorch_powered_descent.py:160-183 loop body; Actor.forward/sample,
 * src/agents/sac_pytorch.py:129-179: state -> [Linear+ReLU] x 2 -> (mean, clamp(log_std,-20,2)),
 * action = tanh(mean + std * eps) * max_action).  Runs n_steps x [actor inference -> fused env
 * step with auto-reset] on the handle's batch (type 'rl', PD_FP32 build) and writes the
 * transitions step-major.  hidden == 256 and fp32_path == 0: the 256x256 layer runs on the
 * tcgen05 tensor cores with IEEE-half operands and fp32 accumulation (~3e-4 of the activation
 * scale against torch fp32); otherwise an exact fp32 CUDA-core kernel.
 *   w1 [H*O], b1 [H], w2 [H*H], b2 [H], wm [A*H], bm [A], ws [A*H], bs [A]  dev float
 *   (torch nn.Linear layouts: weight[out][in])
 *   obs_out      dev float[n_steps*n_envs*O]  observation the action was computed from
 *   act_out      dev float[n_steps*n_envs*A]  (required: the step kernel reads its action here)
 *   rew_out      dev float[n_steps*n_envs]
 *   done_out, trunc_out  dev uint8[n_steps*n_envs]
 *   next_obs_out dev float[n_steps*n_envs*O]  observation of the post-step, pre-reset state
 *   (obs_out, rew_out, done_out, trunc_out, next_obs_out may be NULL). */
typedef struct {
    int32_t hidden;           /* H */
    int32_t deterministic;    /* != 0 -> tanh(mean) */
    float max_action;
    int32_t fp32_path;        /* != 0 -> force the fp32 CUDA-core actor */
    const float *w1, *b1, *w2, *b2, *wm, *bm, *ws, *bs;
    uint64_t seed;            /* Philox key of the action noise */
} PdSharedActor;
int pd_collect_shared_actor(PdEnv *env, const PdSharedActor *actor, int n_steps, float *obs_out,
                            float *act_out, float *rew_out, uint8_t *done_out,
                            uint8_t *trunc_out, float *next_obs_out, void *stream);

/* Actor inference alone on a caller-supplied observation batch (dev float[n*O] -> dev
 * float[n*A]; mean_out optional): the numerics hook for the tensor-core path. */
int pd_actor_forward(PdEnv *env, const PdSharedActor *actor, const float *obs, int n, float *act,
                     float *mean_out, void *stream);

/* Device-resident swarm update of one PSO generation for this rank's n particles
 * (particle_swarm_optimisation.py:431-436 personal bests, :517-521 velocity with the sub-swarm
 * best and ONE scalar r1, r2 per particle, :112-118 position clamp).  All pointers dev.
 *   x, v, best [n*P] fp64 (updated in place), best_fit [n], fitness [n] (of x), swarm_of int32[n],
 *   swarm_best [S*P] (sub-swarm bests, already including this generation),
 *   weights_out float[n*P] or NULL (fp32 copy of the new x for the next pd_rollout_pso).
 *   r1, r2 = Philox(seed, generation, index0 + i): independent of the sharding.  Rows with
 *   swarm_of < 0 are skipped.  The update is evaluated in NumPy's order without fused
 *   multiply-adds, so the host drop-in (rng='philox') follows the identical trajectory. */
int pd_pso_update(double *x, double *v, double *best, double *best_fit, const double *fitness,
                  const int32_t *swarm_of, const double *swarm_best, float *weights_out, int n, int P,
                  int64_t index0, double w, double c1, double c2, double lo, double hi, uint64_t seed,
                  int generation, void *stream);

/* The rest of one PSO generation on the device, no host synchronisation anywhere
 * (ParticleSubswarmOptimisation.run, particle_swarm_optimisation.py:425-477).  All pointers dev.
 *   pd_pso_seed_mean  fitness[n*n_seeds] -> out[n], mean over a particle's wind seeds (sequential sum)
 *   pd_pso_select     allfit[N] = this generation's fitness of EVERY particle (the all-gathered
 *                     slices), swarm_of_all int32[N] (sub-swarm id, < 0 = slot not in use).  Per
 *                     sub-swarm k: sel_idx[k] = global index of its best particle (first occurrence),
 *                     improved[k] = it beat swarm_best_fit[k] (which is updated in place), and the
 *                     reference's per-generation metrics stats[(S+1)*6] = best so far, avg, min, max,
 *                     std (population), count per sub-swarm (:455-470); row S = the whole swarm.
 *   pd_pso_gather     cand[S*P]: row k = x[sel_idx[k] - lo] if this rank owns that particle and
 *                     improved[k], else zeros - so that ONE all-reduce(sum) of cand over the ranks
 *                     is the broadcast of every improved sub-swarm best from its (unknown) owner.
 *   pd_pso_apply      swarm_best[k] <- cand[k] where improved[k]; global best = sequential scan of
 *                     the sub-swarm bests (:474-477); hist_row[0] = global best fitness (or NULL). */
int pd_pso_seed_mean(const double *fitness, int n, int n_seeds, double *out, void *stream);
int pd_pso_select(const double *allfit, const int32_t *swarm_of_all, int N, int S, double *swarm_best_fit,
                  int32_t *sel_idx, int32_t *improved, double *stats, void *stream);
int pd_pso_gather(const double *x, int64_t lo, int n_local, int P, const int32_t *sel_idx,
                  const int32_t *improved, int S, double *cand, void *stream);
int pd_pso_apply(const double *cand, const int32_t *improved, int S, int P, double *swarm_best,
                 const double *swarm_best_fit, double *gbest_pos, double *gbest_fit, double *hist_row,
                 void *stream);

/* Gust-noise stream of the next pd_rollout_pso / pd_rollout_policy calls.  The Philox counter of
 * an episode's noise is (index0 * n_seeds + local episode, draw, episode word = generation + 1):
 * index0 = GLOBAL index of this rank's first particle (block-sharded swarms: the windy fitness of a
 * particle then does not depend on the rank layout), generation = PSO generation (fresh gusts every
 * generation, as upstream draws fresh noise on every reset: src/envs/wind/vonkarman.py:86-96).
 * Default (0, 0).  The physics constants of a handle travel as a __grid_constant__ kernel parameter
 * of every launch: there is no process-global constant state, handles on different streams or
 * devices are independent, and a captured CUDA graph stays valid whatever other handles do. */
int pd_set_rollout_stream(PdEnv *env, int64_t index0, uint32_t generation);

/* Sticky device status (synchronises): 0 = ok; bit 0 = an aero-table query fell outside the
 * enumerated neighbour-set table, bit 1 = neighbour search did not converge.  Cannot happen for
 * finite states (the enumeration covers the whole clamped query box); a lane that hits it has its
 * C_L / C_D set to NaN, so the env turns NaN instead of continuing on a wrong interpolant. */
int pd_check_status(PdEnv *env, int32_t *status);

/* Measured peak of the FP32 (fp64 = 0) or FP64 (fp64 = 1) FMA pipe of `device`, TFLOP/s at 2 FLOP
 * per FMA: 8 independent FMA chains per thread, 64 warps per SM, best of 3 timed launches (CUDA
 * events).  Synchronous.  The denominator of bench.py's fp_roofline. */
int pd_measure_fma_peak(int device, int fp64, double *tflops, double *kernel_ms);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t pd_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* PD_B200_H */
