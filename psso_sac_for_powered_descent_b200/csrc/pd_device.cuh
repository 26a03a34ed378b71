// pd_device.cuh - device-side physics of the powered-descent hot path (sm_100a).
//
// One environment per thread, state in registers.  Everything here is templated on the
// compute type R (double = parity build, float = production build); the 11-state itself
// is always carried in double (HBM traffic is <1% of the roofline for this path, see
// DESIGN.md), and quantities that are differences of large state values (alpha_effective,
// gamma) are formed in double in both builds.
//
// Reference algorithm (file:line relative to the reference root):
//   isa()            src/envs/utils/atmosphere_dynamics.py:5-27 (+ ambiance ISA)
//   cog_inertia()    src/RocketSizing/functions/rocket_dimensions.py:167-196
//   rbf_*()          src/envs/utils/aerodynamic_coefficients.py:57-66 (scipy local TPS RBF)
//   coef_cd/cl()     aerodynamic_coefficients.py:105-132 + rockets_physics.py:711-712
//   gridfin_*()      src/envs/utils/grid_fin_aerodynamics.py:7-46
//   acs()            src/envs/utils/acs_model.py:13-86
//   control_*()      src/envs/rockets_physics.py:340-400 (P), 168-269 (G)
//   substep()        src/envs/rockets_physics.py:455-646
//   wind             src/envs/wind/full_wind_model.py:35-43, vonkarman.py:33-36
//   rtd_*()          src/envs/pso/rtd_pso.py:172-317, src/envs/rl/rtd_rl.py:194-336
// Phases outside the landing burns (RL mode only, as upstream):
//   control_ascent() rockets_physics.py:17-56 ; control_rcs() :149-166 ; control_C() :402-451
//   cog_inertia_full()  src/RocketSizing/functions/rocket_dimensions.py:199-241
//   rtd ascent / ballistic / P-control   src/envs/rl/rtd_rl.py:11-114, 147-188, 353-534
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PD_PI 3.141592653589793
#define PD_TWO_PI 6.283185307179586

namespace pd {

// ------------------------------------------------------------------ flight phases
// 0 landing_burn_pure_throttle, 1 landing_burn, 2 subsonic, 3 supersonic,
// 4 ballistic_arc_descent, 5 landing_burn_pure_throttle_Pcontrol, 6 flip_over_boostbackburn
__host__ __device__ constexpr int phase_adim(int ph) { return ph == 1 ? 4 : ((ph == 2 || ph == 3) ? 2 : 1); }
__host__ __device__ constexpr int phase_odim(int ph) {
    return ph == 0 ? 2 : ph == 1 ? 5 : (ph == 2 || ph == 3) ? 8 : ph == 4 ? 4 : ph == 6 ? 2 : 1;
}
// phases that carry the gimbal (and, landing_burn, the fin commands) from one env step to the next
__host__ __device__ constexpr bool phase_has_actuator_memory(int ph) { return ph == 1 || ph == 6; }
__host__ __device__ constexpr int phase_nsub(int ph) { return ph <= 1 ? 4 : 1; }
__host__ __device__ constexpr bool phase_ascent(int ph) { return ph == 2 || ph == 3; }

// ------------------------------------------------------------------ constants
template <typename R>
struct Scalars {
    R T_e, p_e, A_e, te_over_vex, n_eng, te_neng_dummy;
    R nominal, one_minus_nominal;
    R S_gf, d_gf, R_rocket, S_ref, m_prop0, c_gust_x;
    R I_dry, h_f, h_lower, h_ox, m_dry, m_f, m_ox, x_dry, engine_height, cop;
    R dt_phys, dt_act;
    R max_gimbal_rad, max_gimbal_deg, max_defl_rad;
    R norm_y, norm_vy, norm_x, norm_vx;
    R y0, mass0;
    R k_theta_pso, k_theta_rl, k_thetadot_rl;
    R rl_reward_scale;
    R v_opt_a, v_opt_b;
    // ISA layers 0..7
    R isa_Hb[8], isa_Tb[8], isa_beta[8], isa_pb[8], isa_boT[8], isa_expo[8], isa_iso[8];
    // wind profile
    R wind_x[16], wind_y[16], wind_slope[16];
    R Adu[4], Bdu[2], Adv[4], Bdv[2];
    R cd_levels[5], cl_levels[5];
    // fast_log constants, read as constant-bank operands (64-bit immediates would cost a pair
    // of UMOVs per use): 1/5, -1/4, 1/3, -1/2, ln 2, 2^52 + 1023, -1
    R log_c[8];
    // ---- phases 2..5
    R n_eng_ng;                     // non-gimballed engines (ascent)
    R fi[13];                       // full_rocket_inertia cells: m_s_1, x_dry_1, I_dry_1, m_2, m_pay,
                                    // x_wet_2_initial, I_wet_2_initial, h_1, h_1_ox, h_1_f, m_1_ox,
                                    // m_1_f, h_lower_1
    R rcs_force, d_rcs_bottom, d_rcs_top;
    R norm8[8];                     // obs normalisation of phases 2..4
    R speed0, terminal_mach, alive_bonus;
    R sup_terminal_alt;             // supervisory closures: last altitude of the supersonic recording
    R flip_max_gimbal_deg;          // flip_over_boostbackburn: 10 (rockets_physics.py:759)
    // Mach-scheduled ascent thresholds (rtd_rl.py:544-575): grid, 4 value rows (max_x, max_vy,
    // max_vx, max_alpha_deg) and their segment slopes; the reward weights are 100 everywhere
    R hyp_m[12], hyp_v[4][12], hyp_s[4][12];
};

// Uniform (Mach, AoA) lookup grid over one query box.  cells >= 0: the grid cell lies
// wholly inside one order-50 Voronoi cell, value = set id (no search).  cells < 0:
// -(k+1) -> (imp_hint[k], imp_id[k]) = the set at the cell centre; walk from there.
struct RbfGrid {
    double m0, inv_dm, a0, inv_da;
    int nm, na;
    const int *cells;
    const unsigned long long *imp_hint;
    const int *imp_id;
    // fp32 build: bicubic patches per sub-cell (pd_patch.h); pbase == nullptr: none
    const int2 *pbase;       // [nm*na]: x = (first patch << 1) | two_sets, -1 = no patch (walk cell);
                             //          y = point slots p << 8 | q of the bisector (two_sets cells)
    const float *patch;      // [n][16] C[q][p] of u^p v^q (constant = [0] + [15]), [0] = NaN: rejected
    int sub_x, sub_y;
};

struct RbfDev {
    const double *mach;      // [n_points] level-major, Mach ascending (used by the walk)
    const double2 *points;   // [n_points] (Mach, AoA) of every slot; staged to shared memory
    const double *rows;      // [n_sets][64]: 50 coeffs, 3 poly, shift(2), scale(2), 50 index bytes
    const unsigned long long *hkeys;
    const int *hvals;
    int hash_mask;
    int n_levels;
    int n_points;
    int off[6];
    RbfGrid grid[2];
};

struct Tables {
    RbfDev cd, cl;
    const double2 *logtab;              // [256] (1/c_j rounded, -log of that), see fast_log
    const void *sh_image;               // SharedTables image (host-replicated), source of the TMA bulk copy
    // ascent reference trajectory, sorted by altitude: x(y), vx(y), vy(y) values and slopes
    const double *ref_y, *ref_v[3], *ref_s[3];
    int n_ref;
    const double *ca_x, *ca_y, *ca_s;   // grid fin C_a segments (x_lo, y_lo, slope)
    const double *cn_x, *cn_y, *cn_s;
    int n_ca, n_cn;
    // 256-bin index tables of the two abscissa vectors: lut[b] = number of entries below the start of
    // bin b (seg_index_lut)
    const unsigned char *ca_lut, *cn_lut;
    double ca_x0, ca_inv_w, cn_x0, cn_inv_w;
    int n_wind;
    double init[11];
};

// Block-shared staging of the small hot tables.  Every entry is replicated PD_REP = 8 times,
// entry i of copy c at [i * 8 + c]; lane l uses copy l & 7.  A 128-bit shared load is served a
// quarter-warp (8 lanes) at a time, and with this layout the 8 lanes always hit 8 different
// 16-byte bank groups, whatever (random) entries they gather: 4 wavefronts per request instead
// of ~11 (the LSU data pipe was the busiest unit of the step kernel at 62 %,
// profiles/r1_step_kernel_final_steady.txt).  74 KB of dynamic shared memory, one block per SM.
// The replicated image is built once on the host (pd_create) and pulled into shared memory by
// ONE TMA bulk copy per block (cp.async.bulk + mbarrier) while the threads load their env state:
// the per-launch staging loop it replaces cost ~5 % of the step kernel (STS stalls in the
// prologue, gpurun_out/prof_step_r1f).
#define PD_REP 8
// Each table starts on a 32 KB boundary of the shared window (the kernels round the dynamic
// shared base up), so a gather address is base | (field & 0x7F80): one shift + one LOP3 per
// lookup instead of shift + mask + scaled add.
// PD_LOG_BITS = mantissa bits that index the log table.  8 (256 entries, degree 5 / 4) is the
// default.  9 (512 entries, |r| < 2^-10, one polynomial degree and one DFMA per term less) was
// measured 1.5 % SLOWER on B200 (80.0 vs 78.8 us per 65 536-env step): the larger image costs more
// to stage per launch than the saved FP64 instruction buys - the kernel is not FP64-bound.
#ifndef PD_LOG_BITS
#define PD_LOG_BITS 8
#endif
#define PD_LOG_N (1 << PD_LOG_BITS)
struct SharedTables {
    double2 logtab[PD_LOG_N * PD_REP];  // 32 KB (256 entries)
    double2 cd_pts[256 * PD_REP];       // 32 KB slot, 192 used
    double2 cl_pts[144 * PD_REP];
    unsigned long long bar;             // mbarrier of the bulk copy (not part of the image)
};
constexpr unsigned PD_SH_ALIGN = PD_LOG_N * PD_REP * 16;                  // the largest table's span
constexpr unsigned PD_SH_BYTES = sizeof(SharedTables) + PD_SH_ALIGN;      // dynamic shared memory per block
constexpr unsigned PD_SH_IMAGE_BYTES = (PD_LOG_N + 256 + 144) * PD_REP * 16;
constexpr unsigned PD_LOG_MASK = (PD_LOG_N - 1) << 7;                     // byte-offset field of a table index

// Per-handle constants: the whole block travels as ONE __grid_constant__ kernel parameter (4.4 KB of
// the 32 KB parameter space), so it sits in constant bank 0 of every launch, two handles can be
// stepped on two streams at the same time, and a captured CUDA graph carries its own copy - there
// is no process-global __constant__ state and no "active handle".
struct KParams {
    Scalars<double> sd;
    Scalars<float> sf;
    Tables tb;
#ifdef PD_EXP_PARAM_PAD
    char pad[PD_EXP_PARAM_PAD];      // experiment: does the launch cost depend on the block's size?
#endif
};

// ------------------------------------------------------------------ small math shims
__device__ __forceinline__ double m_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float m_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double m_exp(double x) { return exp(x); }
__device__ __forceinline__ float m_exp(float x) { return expf(x); }
__device__ __forceinline__ double m_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ float m_pow(float x, float y) { return powf(x, y); }
// fp32 production build: the ISA pressure
// ratio (1 + z)^y, z = beta/T_b (H - H_b) in (-0.3, 0.3), y = 5.26 or -34.2 / 12.2, as
// exp(y log1p(z)).  Forming 1 + z in float first would already cost 6e-8 * |y| = 2e-6 relative on
// p in the stratosphere layers - as much as the MUFU-only __powf this replaces (4e-6), which showed
// up as a systematic drift of q and 62 flag disagreements with the fp64 build per 65.5 M env-steps,
// all of them q within 5e-6 of the 65 kPa threshold (tools/fp32_flag_probe.py).  log1pf / expf
// are 1-2 ulp: ~3e-7 relative on p and rho for ~40 instructions, a third of powf's cost and more
// accurate than powf(1 + z, y), so every phase and the rtd closures' q use it.
__device__ __forceinline__ double m_pow1p(double z, double y) { return pow(1.0 + z, y); }
__device__ __forceinline__ float m_pow1p(float z, float y) { return expf(y * log1pf(z)); }
__device__ __forceinline__ double m_log(double x) { return log(x); }
__device__ __forceinline__ float m_log(float x) { return logf(x); }
__device__ __forceinline__ double m_tanh(double x) { return tanh(x); }
__device__ __forceinline__ float m_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ void m_sincos(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __forceinline__ void m_sincos(float x, float *s, float *c) { sincosf(x, s, c); }
// body -> inertial rotation by the pitch angle (O(1) rad, |x| <= 2 pi): MUFU sin / cos are good to
// ~1e-6 ABSOLUTE there, a 1e-9 relative effect on the state per sub-step.  Not for the small
// angles (alpha_eff, gimbal, fins), whose sines need relative accuracy.
__device__ __forceinline__ void m_sincos_pitch(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __forceinline__ void m_sincos_pitch(float x, float *s, float *c) { __sincosf(x, s, c); }
__device__ __forceinline__ double m_abs(double x) { return fabs(x); }
__device__ __forceinline__ float m_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double m_min(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float m_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double m_hypot(double a, double b) { return hypot(a, b); }
__device__ __forceinline__ float m_hypot(float a, float b) { return hypotf(a, b); }
// division: IEEE in the parity build; reciprocal-multiply (2 ulp, no slow-path branch) in the
// fp32 production build, whose tolerance is 1e-5
__device__ __forceinline__ double m_div(double a, double b) { return a / b; }
__device__ __forceinline__ float m_div(float a, float b) { return __fdividef(a, b); }

// ------------------------------------------------------------------ per-env registers
struct State {
    double x, y, vx, vy, theta, theta_dot, gamma, alpha, mass, m_prop, time;
};

struct ActPrev {   // landing_burn only: gimbal_angle_deg_prev, delta_command_{left,right}_rad_prev
    double gimbal_deg, dl, dr;
};

struct WindState {
    double xu0, xu1, xv0, xv1;     // gust filter states
    double sigma_u, sigma_v;
    unsigned int ctr;              // draws consumed this episode (tape position / Philox counter)
    unsigned int episode;          // episode number of this env / rollout generation (Philox counter word)
    // cooperative rollouts: Gaussian pair of draw actr + 2 * (lane mod COOP), made by the first
    // lanes of the group for all sub-steps of an env step at once (gust_ahead)
    double an0, an1;
    unsigned int actr;
};

// values of the last sub-step (reference `info` dict, rockets_physics.py:649-702).  FULL adds
// the remaining primitives of that dict (forces, moments, accelerations, actuator outputs); it is
// only instantiated for the fp64 diagnostic kernels so the production kernels keep their
// register budget.
#define PD_INFO_X 32
template <typename R, bool FULL = false>
struct Info {
    R mach, q, CL, CD, rho, p_atm, a, x_cog, inertia, mass_flow, throttle, alpha_eff, ug, vg;
    int rbf_status;
};
template <typename R>
struct Info<R, true> {
    R mach, q, CL, CD, rho, p_atm, a, x_cog, inertia, mass_flow, throttle, alpha_eff, ug, vg;
    int rbf_status;
    // 0 drag, 1 lift, 2 d_cp_cg, 3 d_thrust_cg, 4 fuel_percentage_consumed, 5 control_force_parallel,
    // 6 control_force_perpendicular, 7 control_force_x, 8 control_force_y, 9 aero_force_x,
    // 10 aero_force_y, 11 g, 12 control_moment_z, 13 aero_moment_z, 14 moments_z, 15 theta_dot_dot,
    // 16 vx_dot, 17 vy_dot, 18 F_wind_x, 19 gimbal_angle_deg, 20 delta_command_left_rad,
    // 21 delta_command_right_rad, 22 mach_number_max (Qmax = 65 kPa landing phases, 30 kPa others),
    // 23 pitch angle at the start of the sub-step, 24 C_a(M), 25 C_n_alpha(M) [per degree],
    // 26 / 27 filtered left / right fin deflection [rad] (acs_info, acs_model.py:62-84)
    R x[PD_INFO_X];
};

// ------------------------------------------------------------------ plain per-env types
struct WindCtx {
    const double *tape;     // N(0,1) tape or nullptr -> Philox
    int tape_len;
    unsigned long long seed;
    int stochastic;
    unsigned int id_offset; // added to the env / episode index for the Philox stream id only (global
                            // particle index of a sharded swarm); the tape is indexed by the local id
};

// The action as the reference sees it: either float64 (pure fp64 step) or float32
// (NumPy NEP-50: throttle / thrust / mass-flow become float32; SURVEY 8a dtype rule).
template <int A>
struct Action {
    double u[A];
    bool f32;
};

template <typename R>
struct Control {
    R par, perp, mz, mass_flow_dt;   // mass_flow * dt_phys, rounded as the reference rounds it
    R mass_flow, throttle;
    double gimbal_deg, dl_cmd, dr_cmd;
    bool f32_forces;                 // par / perp are np.float32 upstream (ascent, float32 action):
                                     // the body->inertial rotation then runs in float32 too
};

template <typename R>
struct GWindow {
    R w[10];
    int n;
};

template <typename R>
struct Rtd {
    R reward;
    int done, truncated, trunc_id;
};

// ------------------------------------------------------------------ device physics
// All functions that read the per-handle constants are members of Dev, which only holds three
// references into the kernel's __grid_constant__ parameter block; everything is force-inlined, so
// the references fold into direct constant-bank operands.
struct Dev {
    const Scalars<double> &sd;
    const Scalars<float> &sf;
    const Tables &tb;
    __device__ __forceinline__ explicit Dev(const KParams &k) : sd(k.sd), sf(k.sf), tb(k.tb) {}
    template <typename R>
    __device__ __forceinline__ const Scalars<R> &SC() const {
        if constexpr (sizeof(R) == 8) return sd; else return sf;
    }

// ------------------------------------------------------------------ ISA
template <typename R, bool FAST = false>
__device__ __forceinline__ void isa(R alt, R &rho, R &p, R &a) {
    const Scalars<R> &c = SC<R>();
    if (alt < R(0)) alt = R(0);
    if (!(alt < R(81020))) {
        rho = p = a = R(0);
        return;
    }
    const R RE = R(6356766.0);
    R H = m_div(RE * alt, RE + alt);
    int k = 1;
#pragma unroll
    for (int j = 2; j < 8; ++j)
        if (H >= c.isa_Hb[j]) k = j;
    R dH = H - c.isa_Hb[k];
    R beta = c.isa_beta[k];
    R T = c.isa_Tb[k] + beta * dH;
    if (beta == R(0))
        p = c.isa_pb[k] * m_exp(c.isa_iso[k] * dH);
    else
        p = c.isa_pb[k] * m_pow1p(c.isa_boT[k] * dH, c.isa_expo[k]);
    const R Rgas = R(287.05287);
    rho = m_div(p, Rgas * T);
    a = m_sqrt(R(1.4 * 287.05287) * T);
}

// density only (the rtd closures need q at the new state)
template <typename R>
__device__ __forceinline__ R isa_rho(R alt) {
    R rho, p, a;
    isa<R>(alt, rho, p, a);
    return rho;
}

// ------------------------------------------------------------------ inertia
template <typename R>
__device__ __forceinline__ void cog_inertia(R fill, R &x_cog, R &inertia) {
    const Scalars<R> &c = SC<R>();
    R h_ox_t = c.h_ox * fill;
    R h_f_t = c.h_f * fill;
    R m_ox_t = c.m_ox * fill;
    R m_f_t = c.m_f * fill;
    R a_ox = c.h_lower + h_ox_t / R(2);
    R a_f = c.h_lower + c.h_ox + h_f_t / R(2);
    R m_p = m_ox_t + m_f_t;
    R x_prop = m_div(m_ox_t * a_ox + m_f_t * a_f, m_p);
    R d_ox = a_ox - x_prop;
    R d_f = a_f - x_prop;
    const R twelfth = R(1.0 / 12);
    R I_ox = twelfth * m_ox_t * (h_ox_t * h_ox_t) + m_ox_t * (d_ox * d_ox);
    R I_f = twelfth * m_f_t * (h_f_t * h_f_t) + m_f_t * (d_f * d_f);
    R I_prop = I_ox + I_f;
    R x_wet = m_div(c.m_dry * c.x_dry + m_p * x_prop, c.m_dry + m_ox_t + m_f_t);
    R dd = c.x_dry - x_wet;
    R dp = x_prop - x_wet;
    R I_dry_hat = c.I_dry + c.m_dry * (dd * dd);
    R I_prop_hat = I_prop + m_p * (dp * dp);
    x_cog = x_wet;
    inertia = I_dry_hat + I_prop_hat;
}

// full_rocket_inertia (rocket_dimensions.py:199-241): the ascent phases' closure.  Note the
// upstream asymmetries kept as they are: the fuel column sits on the *current* oxidiser height,
// and x_prop weights it with the full fuel mass m_1_f.
template <typename R>
__device__ __forceinline__ void cog_inertia_full(R fill, R &x_cog, R &inertia) {
    const R *f = SC<R>().fi;
    const R m_s_1 = f[0], x_dry_1 = f[1], I_dry_1 = f[2], m_2 = f[3], m_pay = f[4], x_wet_2 = f[5],
            I_wet_2 = f[6], h_1 = f[7], h_1_ox = f[8], h_1_f = f[9], m_1_ox = f[10], m_1_f = f[11],
            h_lower_1 = f[12];
    R h_ox = h_1_ox * fill, h_f = h_1_f * fill, m_ox = m_1_ox * fill, m_f = m_1_f * fill;
    R m_prop = m_ox + m_f;
    R a_ox = h_lower_1 + h_ox / R(2);
    R a_f = h_lower_1 + h_ox + h_f / R(2);
    R x_prop = m_div(m_ox * a_ox + m_1_f * a_f, m_ox + m_f);
    const R twelfth = R(1.0 / 12);
    R d_ox = a_ox - x_prop, d_f = a_f - x_prop;
    R I_ox = twelfth * m_ox * (h_ox * h_ox) + m_ox * (d_ox * d_ox);
    R I_f = twelfth * m_f * (h_f * h_f) + m_f * (d_f * d_f);
    R I_prop = I_ox + I_f;
    R x_r = m_div(m_s_1 * x_dry_1 + (m_2 + m_pay) * (x_wet_2 + h_1) + m_prop * x_prop,
                  m_s_1 + m_2 + m_pay + m_prop);
    R d1 = x_dry_1 - x_r, d2 = x_wet_2 - x_r, d3 = x_prop - x_r;
    x_cog = x_r;
    inertia = I_dry_1 + m_s_1 * (d1 * d1) + I_wet_2 + m_2 * (d2 * d2) + I_prop + m_prop * (d3 * d3);
}

// ------------------------------------------------------------------ local TPS RBF
static __device__ __forceinline__ unsigned long long hash_u64(unsigned long long k) {
    k ^= k >> 30;
    k *= 0xBF58476D1CE4E5B9ULL;
    k ^= k >> 27;
    k *= 0x94D049BB133111EBULL;
    k ^= k >> 31;
    return k;
}

#define PD_RBF_OK 0
#define PD_RBF_MISS 1
#define PD_RBF_ITER 2

// Slow path (impure grid cells only): walk from the candidate set `hint` to the exact 50-NN set
// of (M, a).  Sets are one contiguous [lo,hi) interval per level; the check compares the
// farthest interval end against the nearest point just outside any interval.
template <int NL>
static __device__ __noinline__ int rbf_walk(const RbfDev &T, const double *levels_d, double M, double a,
                                     unsigned long long hint, int &sid) {
    int lo[NL], hi[NL];
    double dl2[NL];
    const unsigned long long hint_in = hint;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        lo[l] = (int)((hint >> (12 * l)) & 63);
        hi[l] = (int)((hint >> (12 * l + 6)) & 63);
        double d = levels_d[l] - a;
        dl2[l] = d * d;
    }
    int status = PD_RBF_OK;
    int it = 0;
    for (; it < 400; ++it) {
        double max_in = -1.0, min_out = 1e300;
        int in_l = 0, in_side = 0, out_l = 0, out_side = 0;
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const double *m = T.mach + T.off[l];
            const int n = T.off[l + 1] - T.off[l];
            if (lo[l] == hi[l]) {     // empty level: park at the insertion point of M
                int p = lo[l];
                while (p > 0 && __ldg(m + p - 1) > M) --p;
                while (p < n && __ldg(m + p) < M) ++p;
                lo[l] = hi[l] = p;
            } else {
                double d0 = __ldg(m + lo[l]) - M;
                d0 = d0 * d0 + dl2[l];
                double d1 = __ldg(m + hi[l] - 1) - M;
                d1 = d1 * d1 + dl2[l];
                if (d0 > max_in) { max_in = d0; in_l = l; in_side = 0; }
                if (d1 > max_in) { max_in = d1; in_l = l; in_side = 1; }
            }
            if (lo[l] > 0) {
                double d = __ldg(m + lo[l] - 1) - M;
                d = d * d + dl2[l];
                if (d < min_out) { min_out = d; out_l = l; out_side = 0; }
            }
            if (hi[l] < n) {
                double d = __ldg(m + hi[l]) - M;
                d = d * d + dl2[l];
                if (d < min_out) { min_out = d; out_l = l; out_side = 1; }
            }
        }
        if (max_in <= min_out) break;
        // swap: drop the farthest member, take the nearest outsider.  A one-point interval
        // that trades its point with a neighbour on the same level must give up the end
        // opposite to the one it grows at (otherwise the two updates cancel).
        if (in_l == out_l) {
#pragma unroll
            for (int l = 0; l < NL; ++l)
                if (l == in_l && hi[l] - lo[l] == 1) in_side = 1 - out_side;
        }
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            if (l == in_l) { if (in_side == 0) ++lo[l]; else --hi[l]; }
        }
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            if (l == out_l) { if (out_side == 0) --lo[l]; else ++hi[l]; }
        }
    }
    if (it >= 400) status = PD_RBF_ITER;
    unsigned long long h = 0, key = 1ULL << 63;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        h |= ((unsigned long long)lo[l] << (12 * l)) | ((unsigned long long)hi[l] << (12 * l + 6));
        if (lo[l] != hi[l])
            key |= ((unsigned long long)lo[l] << (12 * l)) | ((unsigned long long)hi[l] << (12 * l + 6));
    }
    if (h != hint_in) {
        unsigned int slot = (unsigned int)hash_u64(key) & T.hash_mask;
        int found = -1;
        for (int probe = 0; probe < 64; ++probe) {
            unsigned long long k = __ldg(T.hkeys + slot);
            if (k == key) { found = __ldg(T.hvals + slot); break; }
            if (k == 0ULL) break;
            slot = (slot + 1) & T.hash_mask;
        }
        if (found < 0) status |= PD_RBF_MISS; else sid = found;
    }
    return status;
}

// Stateless set lookup: grid cell -> set id; only cells cut by a Voronoi edge take the walk.
// Split in two so that a caller can put the cell loads of several lookups in flight before it
// consumes the first one (each is an L2 round trip on the critical path of a sub-step).
__device__ __forceinline__ const int *rbf_cell_ptr(const RbfGrid &G, double M, double a) {
    int im = (int)((M - G.m0) * G.inv_dm);
    int ia = (int)((a - G.a0) * G.inv_da);
    im = max(0, min(im, G.nm - 1));
    ia = max(0, min(ia, G.na - 1));
    return G.cells + ia * G.nm + im;
}
// pts: this lane's replica of the table's (Mach, AoA) points in shared memory.
// Impure cells come in two kinds (rbf_sets._resolve_single_edge_cells): cut by exactly one
// order-50 Voronoi edge - bit 63 of the hint; the two sets differ by one swap p <-> q and the
// query belongs to the first iff it is not farther from p than from q: two gathers and six FP64
// operations, no divergence worth the name - or anything else (< 1 % of the cells): the walk.
template <int NL>
__device__ __forceinline__ int rbf_resolve(const RbfDev &T, const RbfGrid &G, const double *levels_d,
                                           const double2 *__restrict__ pts, double M, double a, int cell,
                                           int &status) {
    if (cell >= 0) return cell;
    const int k = -cell - 1;
    int sid = __ldg(G.imp_id + k);
    const unsigned long long hint = __ldg(G.imp_hint + k);
    if (hint >> 63) {
        const double2 P = pts[(int)((hint >> 8) & 255) * PD_REP], Q = pts[(int)(hint & 255) * PD_REP];
        const double px = M - P.x, py = a - P.y, qx = M - Q.x, qy = a - Q.y;
        const double dp = fma(px, px, py * py), dq = fma(qx, qx, qy * qy);
        return dp <= dq ? sid : (int)((hint >> 16) & 0xFFFF);
    }
    status |= rbf_walk<NL>(T, levels_d, M, a, hint, sid);
    return sid;
}

// Natural log of a positive normal double, ~1 ulp, ~18 instructions (CUDA's log() costs ~130
// here and was 83% of the first step kernel, profiles/r1_step_kernel_baseline.txt).
//   x = 2^e * m, m in [1,2); j = top 8 mantissa bits; tab[j] = (u_j, -log(u_j)) with
//   u_j = double(1 / (1 + (j + 0.5)/256)); r = m*u_j - 1 (exact in one fma, |r| < 2^-9);
//   log x = e ln2 - log u_j + log1p(r); log1p by a Taylor polynomial of degree PD_LOG_DEG
//   (degree 5: |r|^6/6 < 1e-17, full double; the fp32 production build uses degree 4,
//   |r|^5/5 < 6e-15, still 1e3 x below what its 1e-5 state tolerance needs after the
//   thin-plate-spline cancellation factor of ~5e4).
// Integer work stays on the high 32-bit word; e is converted with the 2^52 magic-number trick
// (one DADD) instead of an I2F on the XU pipe.  x = 0 yields a finite value (-709), so
// phi(0) = 0 * finite = 0 needs no special case.
template <int DEG>
__device__ __forceinline__ double fast_log(double x, const double2 *__restrict__ tab) {
    const double *K = sd.log_c;
    const int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    const double2 t = tab[((hi >> (20 - PD_LOG_BITS)) & (PD_LOG_N - 1)) * PD_REP];
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    // (double)((hi >> 20) - 1023): biased exponent in the low word of 2^52, minus (2^52 + 1023);
    // x > 0 here, so the sign bit needs no masking
    const double ed = __hiloint2double(0x43300000, hi >> 20) - K[5];
    const double r = fma(m, t.x, K[6]);
    double p;
    if (DEG >= 5) {
        p = fma(r, K[0], K[1]);
        p = fma(p, r, K[2]);
        p = fma(p, r, K[3]);
    } else if (DEG == 4) {
        p = fma(r, K[1], K[2]);
        p = fma(p, r, K[3]);
    } else {
        p = fma(r, K[2], K[3]);
    }
    p = fma(p * r, r, r);
    return fma(ed, K[4], t.y + p);
}

// one thin-plate-spline term accumulated into acc: c2 * r^2 * log r^2 with c2 = c/2 folded on
// the host (exact).
template <int DEG>
__device__ __forceinline__ double tps_acc(double acc, double c2, double M, double a, double2 pt,
                                          const double2 *__restrict__ logtab) {
    const double dm = M - pt.x, da = a - pt.y;
    const double r2 = fma(dm, dm, da * da);
    return fma(c2 * r2, fast_log<DEG>(r2, logtab), acc);
}

// Row of one neighbour set: 64 doubles = 50 coefficients (already halved), 3 polynomial
// coefficients, shift(2), 1/scale(2), then 50 point-index bytes.
struct RbfRow {
    const double2 *c2;
    const unsigned int *ib;
    const double *base;
};
__device__ __forceinline__ RbfRow rbf_row(const double *__restrict__ rows, int sid) {
    RbfRow r;
    r.base = rows + (size_t)sid * 64;
    r.c2 = reinterpret_cast<const double2 *>(r.base);
    r.ib = reinterpret_cast<const unsigned int *>(r.base + 57);
    return r;
}
__device__ __forceinline__ double rbf_poly(const RbfRow &r, double acc, double M, double a) {
    const double2 p0 = __ldg(r.c2 + 25);     // c50, c51
    const double2 p1 = __ldg(r.c2 + 26);     // c52, shift_m
    const double2 p2 = __ldg(r.c2 + 27);     // shift_a, scale_m
    const double isa = __ldg(r.base + 56);
    // the row stores 1/scale (host-side division): two multiplies instead of two fp64 divisions
    // on the serial tail of every evaluation; moves the result by < 1 ulp of the polynomial part
    const double xh = (M - p1.y) * p2.y;
    const double yh = (a - p2.x) * isa;
    return acc + p0.x + p0.y * xh + p1.x * yh;
}

__device__ __forceinline__ unsigned sh_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double2 lds_d2(unsigned addr) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// log with the table entry already fetched (lets the caller issue all table loads of a batch
// of terms before the first dependent fma)
template <int DEG>
__device__ __forceinline__ double fast_log_t(double x, double2 t) {
    const double *K = sd.log_c;
    const int hi = __double2hiint(x);
    // r = m u_j - 1 with m = x 2^-e: the power of two goes into u_j (integer subtract on its high
    // word, exact), so m is never assembled: r = fma(x, u_j 2^-e, -1)
    const int eb = hi & 0x7FF00000;
    const double us = __hiloint2double(__double2hiint(t.x) - eb + 0x3FF00000, __double2loint(t.x));
    const double ed = __hiloint2double(0x43300000, hi >> 20) - K[5];
    const double s = fma(ed, K[4], t.y);            // e ln2 - log u_j : independent of the polynomial
    const double r = fma(x, us, K[6]);
    // log1p(r) = r + r^2 q(r), Estrin form: dependent depth 3 after r instead of 4-5
    const double r2 = r * r;
    double q = fma(r, K[2], K[3]);                  // -1/2 + r/3
    if (DEG >= 5) q = fma(r2, fma(r, K[0], K[1]), q);   // + r^2 (-1/4 + r/5)
    else if (DEG == 4) q = fma(r2, K[1], q);            // - r^2/4
    return s + fma(r2, q, r);
}

// Values of two interpolants (C_L at (M, aL) from set sidL, C_D at (M, aD) from set sidD) in one
// flat, uniform loop: 12 trips x (4 + 4) terms + a 2 + 2 tail, no per-lane trip counts.
// The loop is software-pipelined by hand - profiles/r1_step_kernel_final_steady.txt showed 60 %
// of its stall samples on the two shared-memory gathers (data point, log table) and 14 % on the
// global load of the index word: (1) index words and coefficients of trip w+1 are fetched
// during trip w, (2) all 8 point gathers, then all 8 r^2, then all 8 table gathers are issued
// before the first polynomial.
template <int DEG>
__device__ __forceinline__ void rbf_eval2(const double *__restrict__ rowsL, int sidL,
                                          const double2 *__restrict__ ptsL, double aL,
                                          const double *__restrict__ rowsD, int sidD,
                                          const double2 *__restrict__ ptsD, double aD, double M,
                                          const double2 *__restrict__ logtab, double &vL, double &vD) {
    const RbfRow rl = rbf_row(rowsL, sidL), rd = rbf_row(rowsD, sidD);
    double l0 = 0.0, l1 = 0.0, d0 = 0.0, d1 = 0.0;
    unsigned int wl = __ldg(rl.ib), wd = __ldg(rd.ib);
    double2 cl0 = __ldg(rl.c2), cl1 = __ldg(rl.c2 + 1), cd0 = __ldg(rd.c2), cd1 = __ldg(rd.c2 + 1);
    // 32-bit shared-window addresses of this lane's replica; the tables are 32 KB aligned, so an
    // entry address is base | (index << 7)
    const unsigned bL = sh_addr(ptsL), bD = sh_addr(ptsD), bT = sh_addr(logtab);
#ifndef PD_EXP_SUM_TRIPS
#define PD_EXP_SUM_TRIPS 12      // experiments only: fewer trips = wrong sums, measures the loop's share
#endif
#pragma unroll 1
    for (int w = 0; w < PD_EXP_SUM_TRIPS; ++w) {
        double2 pt[8];
        pt[0] = lds_d2(bL | ((wl << 7) & 0x7F80u)); pt[1] = lds_d2(bD | ((wd << 7) & 0x7F80u));
        pt[2] = lds_d2(bL | ((wl >> 1) & 0x7F80u)); pt[3] = lds_d2(bD | ((wd >> 1) & 0x7F80u));
        pt[4] = lds_d2(bL | ((wl >> 9) & 0x7F80u)); pt[5] = lds_d2(bD | ((wd >> 9) & 0x7F80u));
        pt[6] = lds_d2(bL | ((wl >> 17) & 0x7F80u)); pt[7] = lds_d2(bD | ((wd >> 17) & 0x7F80u));
        const double c[8] = {cl0.x, cd0.x, cl0.y, cd0.y, cl1.x, cd1.x, cl1.y, cd1.y};
        // prefetch trip w + 1 (trip 12 = the tail: words ib[12], coefficient pair c2[24])
        wl = __ldg(rl.ib + w + 1); wd = __ldg(rd.ib + w + 1);
        cl0 = __ldg(rl.c2 + 2 * w + 2); cl1 = __ldg(rl.c2 + 2 * w + 3);
        cd0 = __ldg(rd.c2 + 2 * w + 2); cd1 = __ldg(rd.c2 + 2 * w + 3);
        double r2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double dm = M - pt[i].x, da = ((i & 1) ? aD : aL) - pt[i].y;
            r2[i] = fma(dm, dm, da * da);
        }
        double2 t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = lds_d2(bT | (((unsigned)__double2hiint(r2[i]) >> (13 - PD_LOG_BITS)) & PD_LOG_MASK));
        double lg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) lg[i] = fast_log_t<DEG>(r2[i], t[i]);
        l0 = fma(c[0] * r2[0], lg[0], l0); d0 = fma(c[1] * r2[1], lg[1], d0);
        l1 = fma(c[2] * r2[2], lg[2], l1); d1 = fma(c[3] * r2[3], lg[3], d1);
        l0 = fma(c[4] * r2[4], lg[4], l0); d0 = fma(c[5] * r2[5], lg[5], d0);
        l1 = fma(c[6] * r2[6], lg[6], l1); d1 = fma(c[7] * r2[7], lg[7], d1);
    }
    {
        l0 = tps_acc<DEG>(l0, cl0.x, M, aL, ptsL[(wl & 255) * PD_REP], logtab);
        d0 = tps_acc<DEG>(d0, cd0.x, M, aD, ptsD[(wd & 255) * PD_REP], logtab);
        l1 = tps_acc<DEG>(l1, cl0.y, M, aL, ptsL[((wl >> 8) & 255) * PD_REP], logtab);
        d1 = tps_acc<DEG>(d1, cd0.y, M, aD, ptsD[((wd >> 8) & 255) * PD_REP], logtab);
    }
    vL = rbf_poly(rl, l0 + l1, M, aL);
    vD = rbf_poly(rd, d0 + d1, M, aD);
}

// Cooperative variant: the COOP lanes of a group all hold the same env; lane `sub` takes the
// terms k = sub, sub + COOP, ... of both sums and a butterfly all-reduce (identical result in
// every lane) replaces 50 sequential terms by 50/COOP + log2(COOP) shuffle rounds.  Used by the
// rollout kernel when there are fewer episodes than lanes on the GPU (latency, not throughput).
template <int DEG, int COOP>
__device__ __forceinline__ void rbf_eval2_coop(const double *__restrict__ rowsL, int sidL,
                                               const double2 *__restrict__ ptsL, double aL,
                                               const double *__restrict__ rowsD, int sidD,
                                               const double2 *__restrict__ ptsD, double aD, double M,
                                               const double2 *__restrict__ logtab, double &vL, double &vD) {
    const RbfRow rl = rbf_row(rowsL, sidL), rd = rbf_row(rowsD, sidD);
    const int sub = threadIdx.x & (COOP - 1);
    const unsigned char *ibl = reinterpret_cast<const unsigned char *>(rl.ib);
    const unsigned char *ibd = reinterpret_cast<const unsigned char *>(rd.ib);
    double l = 0.0, d = 0.0;
#pragma unroll
    for (int k = sub; k < 50; k += COOP) {
        l = tps_acc<DEG>(l, __ldg(rl.base + k), M, aL, ptsL[(int)__ldg(ibl + k) * PD_REP], logtab);
        d = tps_acc<DEG>(d, __ldg(rd.base + k), M, aD, ptsD[(int)__ldg(ibd + k) * PD_REP], logtab);
    }
    // only the lanes of this group are guaranteed to be here (other groups of the warp may be
    // between episodes), so the shuffles name exactly the group
    const unsigned gmask = (COOP == 32 ? 0xffffffffu : ((1u << COOP) - 1u)) << ((threadIdx.x & 31) & ~(COOP - 1));
#pragma unroll
    for (int off = 1; off < COOP; off <<= 1) {
        l += __shfl_xor_sync(gmask, l, off);
        d += __shfl_xor_sync(gmask, d, off);
    }
    vL = rbf_poly(rl, l, M, aL);
    vD = rbf_poly(rd, d, M, aD);
}

// ---- fp32 production build: bicubic patches of the thin-plate sums (pd_patch.h)
// A lookup runs in three steps so that the loads of BOTH tables of a sub-step are in flight together
// (the kernel is bound by the latency of these dependent loads, not by their bytes):
//   patch_locate : cell index, local coordinates, one 8-byte load of the cell's patch base and
//                  bisector slots, address of the sub-cell's patch (side of the bisector for a cell
//                  cut by one Voronoi edge); a cell without patches points at patch 0
//   patch_fetch  : the 4 x 16-byte loads of the patch
//   patch_value  : false if there is no usable patch (a cell that needs the walk, or a fit the builder
//                  rejected: first coefficient NaN) - the caller evaluates the exact sum
struct PatchRef {
    float u, v;
    const float4 *c;      // the sub-cell's patch (patch 0 for a cell without patches)
    bool have;
};
struct PatchCoef {
    float4 c[4];
};
__device__ __forceinline__ PatchRef patch_locate(const RbfGrid &G, const double2 *__restrict__ pts, double M, double a) {
    PatchRef r;
    const double x = (M - G.m0) * G.inv_dm, y = (a - G.a0) * G.inv_da;
    int im = (int)x, ia = (int)y;
    im = max(0, min(im, G.nm - 1));
    ia = max(0, min(ia, G.na - 1));
    const int2 e = __ldg(G.pbase + ia * G.nm + im);
    const double fx = (x - (double)im) * (double)G.sub_x, fy = (y - (double)ia) * (double)G.sub_y;
    int sx = (int)fx, sy = (int)fy;
    sx = max(0, min(sx, G.sub_x - 1));
    sy = max(0, min(sy, G.sub_y - 1));
    r.u = (float)(2.0 * (fx - (double)sx) - 1.0);
    r.v = (float)(2.0 * (fy - (double)sy) - 1.0);
    r.have = e.x >= 0;
    const int pb = r.have ? e.x : 0;
    const int two = pb & 1;
    int w = 0;
    if (two) {      // cut by one order-50 Voronoi edge: the side of the bisector of p and q (rbf_resolve)
        const double2 P = pts[((e.y >> 8) & 255) * PD_REP], Q = pts[(e.y & 255) * PD_REP];
        const double px = M - P.x, py = a - P.y, qx = M - Q.x, qy = a - Q.y;
        w = fma(px, px, py * py) <= fma(qx, qx, qy * qy) ? 0 : 1;
    }
    r.c = reinterpret_cast<const float4 *>(G.patch) +
          ((size_t)(pb >> 1) + (size_t)(sy * G.sub_x + sx) * (size_t)(1 + two) + (size_t)w) * 4;
    return r;
}
__device__ __forceinline__ void patch_fetch(const PatchRef &r, PatchCoef &k) {
#pragma unroll
    for (int i = 0; i < 4; ++i) k.c[i] = __ldg(r.c + i);
}
// everything but the constant term is the variation of the coefficient over the sub-cell (2e-3 of its
// value): fp32 Horner, then the constant (a float pair) is added in double
__device__ __forceinline__ bool patch_value(const PatchRef &r, const PatchCoef &k, double &val) {
    const float u = r.u, v = r.v;
    const float r0 = fmaf(fmaf(fmaf(k.c[0].w, u, k.c[0].z), u, k.c[0].y), u, 0.0f);
    const float r1 = fmaf(fmaf(fmaf(k.c[1].w, u, k.c[1].z), u, k.c[1].y), u, k.c[1].x);
    const float r2 = fmaf(fmaf(fmaf(k.c[2].w, u, k.c[2].z), u, k.c[2].y), u, k.c[2].x);
    const float r3 = fmaf(fmaf(k.c[3].z, u, k.c[3].y), u, k.c[3].x);
    const float var = fmaf(fmaf(fmaf(r3, v, r2), v, r1), v, r0);
    val = (double)k.c[0].x + ((double)k.c[3].w + (double)var);
    return r.have && k.c[0].x == k.c[0].x;
}

// exact sum of one table, one lane on its own (partial warps only)
template <int DEG>
__device__ __noinline__ double rbf_eval1(const double *__restrict__ rows, int sid, const double2 *__restrict__ pts,
                                         double a, double M, const double2 *__restrict__ logtab) {
    const RbfRow r = rbf_row(rows, sid);
    const unsigned char *ib = reinterpret_cast<const unsigned char *>(r.ib);
    double acc = 0.0;
#pragma unroll 2
    for (int k = 0; k < 50; ++k)
        acc = tps_acc<DEG>(acc, __ldg(r.base + k), M, a, pts[(int)__ldg(ib + k) * PD_REP], logtab);
    return rbf_poly(r, acc, M, a);
}

// exact sums for the lanes in `need`, by the whole warp: the query of each such lane is broadcast,
// every lane takes terms lane and lane + 32, a butterfly adds them up - ~100 instructions per
// query instead of the 3 500 of a serial sum that the other 31 lanes would wait for
template <int DEG>
__device__ __forceinline__ double rbf_eval1_warp(unsigned need, const double *__restrict__ rows, int sid,
                                                 const double2 *__restrict__ pts, double a, double M,
                                                 const double2 *__restrict__ logtab, double mine) {
    const int lane = threadIdx.x & 31;
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const double Ms = __shfl_sync(0xffffffffu, M, src), as = __shfl_sync(0xffffffffu, a, src);
        const RbfRow r = rbf_row(rows, __shfl_sync(0xffffffffu, sid, src));
        const unsigned char *ib = reinterpret_cast<const unsigned char *>(r.ib);
        double acc = tps_acc<DEG>(0.0, __ldg(r.base + lane), Ms, as, pts[(int)__ldg(ib + lane) * PD_REP], logtab);
        if (lane < 18)
            acc = tps_acc<DEG>(acc, __ldg(r.base + lane + 32), Ms, as, pts[(int)__ldg(ib + lane + 32) * PD_REP], logtab);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == src) mine = rbf_poly(r, acc, Ms, as);
    }
    return mine;
}

// C_L and C_D of one sub-step.
//  C_D: CD_func passes degrees into a clamp written for radians
//       (rockets_physics.py:712, aerodynamic_coefficients.py:108-114)
//  C_L: degrees applied twice (rockets_physics.py:711, aerodynamic_coefficients.py:120-131);
//       the aoa < -10 branch evaluates (Mach, -10) and is not negated upstream.
template <typename R, int COOP>
__device__ __forceinline__ void aero_coefficients(R mach, R alpha_eff, R &C_L, R &C_D, int &status,
                                                  const SharedTables *sh) {
    const double M = (double)mach;
    const double deg = (double)alpha_eff * (180.0 / PD_PI);
    const double lim = 10.0 * (PD_PI / 180.0);
    double aD = fmin(fmax(deg, -lim), lim);
    const double aoa = deg * (180.0 / PD_PI);
    const bool zero = fabs(aoa) < 1e-6;
    const bool neg_line = aoa < -10.0;
    double aL = neg_line ? -10.0 : fmin(fmax(fabs(aoa), 1e-6), 10.0);
    const bool flip = !neg_line && aoa < 0.0;
    if constexpr (sizeof(R) == 8) {
        // fp64 build: pin the two clamped angles in registers.  Under the 128-register cap ptxas
        // otherwise re-derives them from alpha_eff (two multiplies, two fmin / fmax chains: 42 of the
        // 274 instructions of EVERY trip of the summation loop).  An empty asm does not survive to
        // ptxas; an identity shuffle is the one copy it cannot see through.  Measured: fp64 build
        // 81.2 -> 78.5 us per 65 536-env step; the fp32 build loses 2 % (66.1 -> 67.8 us: it is
        // latency-bound, the re-derived instructions ride in its stall slots and the two extra live
        // registers cost more), so it keeps ptxas' choice.
        const unsigned m = __activemask();
        const int me = threadIdx.x & 31;
        aL = __shfl_sync(m, aL, me);
        aD = __shfl_sync(m, aD, me);
    }
    const RbfGrid &GL = neg_line ? tb.cl.grid[1] : tb.cl.grid[0];
    if constexpr (sizeof(R) == 4 && COOP == 1) {
        if (tb.cd.grid[0].pbase) {
            const unsigned lanes = __activemask();
            const int copy = threadIdx.x & (PD_REP - 1);
            double vL = 0.0, vD = 0.0;
            const PatchRef rD = patch_locate(tb.cd.grid[0], sh->cd_pts + copy, M, aD);
            const PatchRef rL = patch_locate(GL, sh->cl_pts + copy, M, aL);
            PatchCoef kD, kL;
            patch_fetch(rD, kD);
            patch_fetch(rL, kL);
            const bool okD = patch_value(rD, kD, vD);
            const bool okL = patch_value(rL, kL, vL);
            const unsigned needD = __ballot_sync(lanes, !okD), needL = __ballot_sync(lanes, !okL);
            if (needD | needL) {
                constexpr int DEG = 4 - (PD_LOG_BITS >= 9 ? 1 : 0);
                int sidD = 0, sidL = 0;
                if (!okD) sidD = rbf_resolve<5>(tb.cd, tb.cd.grid[0], sd.cd_levels, sh->cd_pts + copy, M, aD,
                                                __ldg(rbf_cell_ptr(tb.cd.grid[0], M, aD)), status);
                if (!okL) sidL = rbf_resolve<5>(tb.cl, GL, sd.cl_levels, sh->cl_pts + copy, M, aL,
                                                __ldg(rbf_cell_ptr(GL, M, aL)), status);
                if (lanes == 0xffffffffu) {
                    __syncwarp();
                    vD = rbf_eval1_warp<DEG>(needD, tb.cd.rows, sidD, sh->cd_pts + copy, aD, M, sh->logtab + copy, vD);
                    vL = rbf_eval1_warp<DEG>(needL, tb.cl.rows, sidL, sh->cl_pts + copy, aL, M, sh->logtab + copy, vL);
                } else {
                    if (!okD) vD = rbf_eval1<DEG>(tb.cd.rows, sidD, sh->cd_pts + copy, aD, M, sh->logtab + copy);
                    if (!okL) vL = rbf_eval1<DEG>(tb.cl.rows, sidL, sh->cl_pts + copy, aL, M, sh->logtab + copy);
                }
            }
            C_L = zero ? R(0) : (R)(flip ? -vL : vL);
            C_D = (R)vD;
            if (status) { C_L = R(NAN); C_D = R(NAN); }
            return;
        }
    }
    // both grid-cell loads in flight before either is consumed
    const int cellD = __ldg(rbf_cell_ptr(tb.cd.grid[0], M, aD));
    const int cellL = __ldg(rbf_cell_ptr(GL, M, aL));
    const int copy = threadIdx.x & (PD_REP - 1);          // this lane's replica of the tables
    const int sidD = rbf_resolve<5>(tb.cd, tb.cd.grid[0], sd.cd_levels, sh->cd_pts + copy, M, aD, cellD, status);
    const int sidL = rbf_resolve<5>(tb.cl, GL, sd.cl_levels, sh->cl_pts + copy, M, aL, cellL, status);
    double vL, vD;
    constexpr int DEG = (sizeof(R) == 8 ? 5 : 4) - (PD_LOG_BITS >= 9 ? 1 : 0);
    if constexpr (COOP == 1)
        rbf_eval2<DEG>(tb.cl.rows, sidL, sh->cl_pts + copy, aL, tb.cd.rows, sidD, sh->cd_pts + copy, aD, M,
                       sh->logtab + copy, vL, vD);
    else
        rbf_eval2_coop<DEG, COOP>(tb.cl.rows, sidL, sh->cl_pts + copy, aL, tb.cd.rows, sidD,
                                  sh->cd_pts + copy, aD, M, sh->logtab + copy, vL, vD);
    C_L = zero ? R(0) : (R)(flip ? -vL : vL);
    C_D = (R)vD;
    // A neighbourhood outside the enumerated set table cannot happen for finite states (the
    // enumeration covers the whole clamped query box, rbf_sets.py); if it ever did, the lane's
    // coefficients are poisoned so that the env turns NaN instead of carrying on with the wrong
    // interpolant, and the sticky status bit makes pd_check_status fail.
    if (status) { C_L = R(NAN); C_D = R(NAN); }
}

// ------------------------------------------------------------------ grid fins
// scipy interp1d(kind='linear'): idx = clip(searchsorted(x, v, 'left'), 1, n-1); lo = idx-1;
// y = slope[lo]*(v - x[lo]) + y[lo]
__device__ __forceinline__ int seg_index(const double *x, int n, double v) {
    int a = 0, b = n;            // number of elements < v
    while (a < b) {
        int mid = (a + b) >> 1;
        if (__ldg(x + mid) < v) a = mid + 1; else b = mid;
    }
    int idx = a < 1 ? 1 : (a > n - 1 ? n - 1 : a);
    return idx - 1;
}

// Same index as seg_index, found from a 256-bin table instead of a six-step binary search (six
// dependent loads on the chain of every sub-step): start one bin early - every entry counted there
// is certainly below v whatever the rounding of the bin number - and count on.
__device__ __forceinline__ int seg_index_lut(const double *x, int n, double v, const unsigned char *lut,
                                             double x0, double inv_w) {
    int b = (int)((v - x0) * inv_w) - 1;
    b = max(0, min(b, 255));
    int a = (int)__ldg(lut + b);
    while (a < n && __ldg(x + a) < v) ++a;
    int idx = a < 1 ? 1 : (a > n - 1 ? n - 1 : a);
    return idx - 1;
}

template <typename R>
__device__ __forceinline__ R gridfin_ca(R mach) {
    const double M = (double)mach;
    if (M < __ldg(tb.ca_x)) return (R)__ldg(tb.ca_y);
    int lo = seg_index_lut(tb.ca_x, tb.n_ca, M, tb.ca_lut, tb.ca_x0, tb.ca_inv_w);
    return (R)(__ldg(tb.ca_s + lo) * (M - __ldg(tb.ca_x + lo)) + __ldg(tb.ca_y + lo));
}

// returns C_n_alpha(M); caller multiplies by degrees(alpha_local)
template <typename R>
__device__ __forceinline__ R gridfin_cn_alpha(R mach) {
    const double M = (double)mach;
    const int n = tb.n_cn;
    if (M < __ldg(tb.cn_x)) return (R)__ldg(tb.cn_y);
    double xmax = __ldg(tb.cn_x + n - 1);
    if (M <= xmax) {
        int lo = seg_index_lut(tb.cn_x, n, M, tb.cn_lut, tb.cn_x0, tb.cn_inv_w);
        return (R)(__ldg(tb.cn_s + lo) * (M - __ldg(tb.cn_x + lo)) + __ldg(tb.cn_y + lo));
    }
    return (R)(__ldg(tb.cn_y + n - 1) + __ldg(tb.cn_s + n - 2) * (M - xmax));
}

// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ void philox4x32(unsigned int c0, unsigned int c1, unsigned int c2,
                                           unsigned int c3, unsigned int k0, unsigned int k1,
                                           unsigned int out[4]) {
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

__device__ __forceinline__ double u01(unsigned int a, unsigned int b) {
    // 53-bit uniform in (0,1)
    unsigned long long v = (((unsigned long long)a << 32) | b) >> 11;
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}


// ------------------------------------------------------------------ wind
// One Box-Muller pair of the gust-noise stream.  counter = (stream id, draw, tag, episode): fresh
// noise every episode, as upstream's reset() re-seeds and rebuilds the filters (vonkarman.py:86-96)
__device__ __forceinline__ void gauss_pair(unsigned int env_id, const WindCtx &wc, unsigned int ctr,
                                           unsigned int episode, double &n0, double &n1) {
    unsigned int r[4];
    philox4x32(env_id + wc.id_offset, ctr, 0x57494E44u, episode, (unsigned int)wc.seed,
               (unsigned int)(wc.seed >> 32), r);
    double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
    double rad = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    n0 = rad * co;
    n1 = rad * s;
}

// Cooperative groups: the noise of a sub-step does not depend on the state, so lanes 0..3 of the
// group draw the pairs of the 4 sub-steps of this env step side by side (Philox + log + sqrt +
// sincospi are a quarter of the dependent chain of a windy sub-step) and wind_sample fetches the
// pair of the draw it has reached with a shuffle.  Same counters, same values as the serial path.
template <int NSUB, int COOP>
__device__ __forceinline__ void gust_ahead(WindState &w, const WindCtx &wc, unsigned int env_id) {
    if constexpr (COOP > 1) {
        if (wc.stochastic && !wc.tape) {
            const unsigned int j = threadIdx.x & (COOP - 1);
            w.actr = w.ctr;
            if (j < (unsigned int)NSUB) gauss_pair(env_id, wc, w.ctr + 2u * j, w.episode, w.an0, w.an1);
        }
    }
}

template <typename R, int COOP = 1>
__device__ __forceinline__ void wind_sample(double y, WindState &w, const WindCtx &wc,
                                            unsigned int env_id, R &ug, R &vg) {
    const Scalars<double> &c = sd;
    const int n = tb.n_wind;
    double akm = y / 1000.0;
    double fixed;
    if (akm < c.wind_x[0]) fixed = c.wind_y[0];
    else if (akm > c.wind_x[n - 1]) fixed = c.wind_y[n - 1];
    else {
        int a = 0;
        for (int j = 0; j < n; ++j) a += (c.wind_x[j] < akm) ? 1 : 0;
        int idx = a < 1 ? 1 : (a > n - 1 ? n - 1 : a);
        int lo = idx - 1;
        fixed = c.wind_slope[lo] * (akm - c.wind_x[lo]) + c.wind_y[lo];
    }
    double gu = 0.0, gv = 0.0;
    if (y < 15000.0 && wc.stochastic) {
        double n0, n1;
        if (wc.tape) {
            size_t base = (size_t)env_id * wc.tape_len;
            unsigned int p0 = w.ctr, p1 = w.ctr + 1;
            n0 = p0 < (unsigned)wc.tape_len ? wc.tape[base + p0] : 0.0;
            n1 = p1 < (unsigned)wc.tape_len ? wc.tape[base + p1] : 0.0;
        } else if constexpr (COOP > 1) {
            const int leader = (threadIdx.x & 31) & ~(COOP - 1);
            const unsigned gmask = (COOP == 32 ? 0xffffffffu : ((1u << COOP) - 1u)) << leader;
            const int src = leader + (int)((w.ctr - w.actr) >> 1);
            n0 = __shfl_sync(gmask, w.an0, src);
            n1 = __shfl_sync(gmask, w.an1, src);
        } else {
            gauss_pair(env_id, wc, w.ctr, w.episode, n0, n1);
        }
        w.ctr += 2;
        double bu0 = c.Bdu[0] * w.sigma_u, bu1 = c.Bdu[1] * w.sigma_u;
        double nu0 = (c.Adu[0] * w.xu0 + c.Adu[1] * w.xu1) + bu0 * n0;
        double nu1 = (c.Adu[2] * w.xu0 + c.Adu[3] * w.xu1) + bu1 * n0;
        w.xu0 = nu0; w.xu1 = nu1;
        double bv0 = c.Bdv[0] * w.sigma_v, bv1 = c.Bdv[1] * w.sigma_v;
        double nv0 = (c.Adv[0] * w.xv0 + c.Adv[1] * w.xv1) + bv0 * n1;
        double nv1 = (c.Adv[2] * w.xv0 + c.Adv[3] * w.xv1) + bv1 * n1;
        w.xv0 = nv0; w.xv1 = nv1;
        gu = nu1;      // C = [0, 1]
        gv = nv1;
    }
    ug = (R)(fixed + gu);
    vg = (R)gv;
}

// sigma_u ~ U(0.5, 2.25), sigma_v ~ U(1.25, 2.0) per reset (vonkarman.py:62-63)
__device__ __forceinline__ void wind_reset(WindState &w, const WindCtx &wc, unsigned int env_id,
                                           unsigned int episode, const double *sigma_uv) {
    w.xu0 = w.xu1 = w.xv0 = w.xv1 = 0.0;
    w.ctr = 0;
    w.episode = episode;
    if (sigma_uv) {
        w.sigma_u = sigma_uv[2 * (size_t)env_id];
        w.sigma_v = sigma_uv[2 * (size_t)env_id + 1];
    } else {
        unsigned int r[4];
        philox4x32(env_id + wc.id_offset, episode, 0x5349474Du, 1u, (unsigned int)wc.seed,
                   (unsigned int)(wc.seed >> 32), r);
        w.sigma_u = 0.5 + (2.25 - 0.5) * u01(r[0], r[1]);
        w.sigma_v = 1.25 + (2.0 - 1.25) * u01(r[2], r[3]);
    }
}

// ------------------------------------------------------------------ actions


// ACS (grid fins), acs_model.py:13-86.  d_cmd_* = delta_command_*_rad =
// radians(deflection_command_deg * 60), formed by the caller (its dtype depends on the action).
template <typename R>
__device__ __forceinline__ void acs(R alpha_eff, R q, R mach, R x_cog, double d_cmd_l, double d_cmd_r,
                                    double prev_l, double prev_r, R &f_perp, R &f_par, R &m_z) {
    const Scalars<R> &c = SC<R>();
    const double dt = sd.dt_act;
    R d_l = (R)(prev_l + dt * ((-prev_l + d_cmd_l) / 0.5));
    R d_r = (R)(prev_r + dt * ((-prev_r + d_cmd_r) / 0.5));
    R a_l = alpha_eff - d_l;
    R a_r = alpha_eff - d_r;
    R qS = q * c.S_gf;
    R Ca = gridfin_ca<R>(mach);
    R cna = gridfin_cn_alpha<R>(mach);
    const R r2d = R(180.0 / PD_PI);
    R Cn_L = cna * (a_l * r2d);
    R Cn_R = cna * (a_r * r2d);
    R sl, cl, sr, cr;
    m_sincos(d_l, &sl, &cl);
    m_sincos(d_r, &sr, &cr);
    f_perp = qS * (Cn_R * cr - Cn_L * cl - Ca * (sl - sr));
    f_par = qS * (Ca * (R(2) + cl + cr) - Cn_L * sl + Cn_R * sr);
    m_z = -(c.d_gf - x_cog) * f_perp + c.R_rocket * qS * (Ca * (sr - sl) - Cn_L * cl + Cn_R * cr);
}

// float32-contaminated thrust chain shared by P and G (fp64 build, float32 action):
//   throttle f32, thrust = f32(t_full*n_eng) * throttle, n_tot = thrust / f32(t_full),
//   mass_flow = f32(T_e/v_ex) * n_tot, mass_flow*dt in f32.
__device__ __forceinline__ float f32_throttle(float u) {
    float nn = __fdiv_rn(__fadd_rn(u, 1.0f), 2.0f);
    return __fadd_rn(__fmul_rn(nn, sf.one_minus_nominal), sf.nominal);
}

// landing_burn_pure_throttle: rockets_physics.py:340-400
template <typename R>
__device__ __forceinline__ void control_P(const Action<1> &act, R p_atm, R alpha_eff, R q, R x_cog,
                                          R mach, Control<R> &o) {
    const Scalars<R> &c = SC<R>();
    // ACS with zero deflection and zero memory: perpendicular force and moment cancel
    // exactly, axial force = 4 q S C_a(M) (acs_model.py:49-59 with delta = 0)
    R qS = q * c.S_gf;
    R Ca = gridfin_ca<R>(mach);
    R f_par = qS * (Ca * R(4));
    R t_full = c.T_e + (c.p_e - p_atm) * c.A_e;
    if (sizeof(R) == 8 && act.f32) {
        float thr = f32_throttle((float)act.u[0]);
        float thrust = __fmul_rn((float)((double)t_full * (double)c.n_eng), thr);
        float n_tot = __fdiv_rn(thrust, (float)t_full);
        float mf = __fmul_rn(sf.te_over_vex, n_tot);
        o.par = (R)thrust + f_par;
        o.mass_flow = (R)mf;
        o.mass_flow_dt = (R)__fmul_rn(mf, sf.dt_phys);
        o.throttle = (R)thr;
    } else {
        R u0 = (R)act.u[0];
        R throttle = (u0 + R(1)) / R(2) * c.one_minus_nominal + c.nominal;
        R thrust = t_full * c.n_eng * throttle;
        R n_tot = thrust / t_full;
        R mf = c.te_over_vex * n_tot;
        o.par = thrust + f_par;
        o.mass_flow = mf;
        o.mass_flow_dt = mf * c.dt_phys;
        o.throttle = throttle;
    }
    o.perp = R(0);
    o.mz = R(0);
}

// landing_burn (gimballed + grid fins): rockets_physics.py:168-269
template <typename R>
__device__ __forceinline__ void control_G(const Action<4> &act, const ActPrev &prev, R p_atm,
                                          R d_thrust_cg, R alpha_eff, R q, R x_cog, R mach,
                                          Control<R> &o) {
    const Scalars<R> &c = SC<R>();
    const bool f32 = (sizeof(R) == 8) && act.f32;
    R t_full = c.T_e + (c.p_e - p_atm) * c.A_e;
    // gimbal: first-order low-pass (tau 1.0, dt_act) on degrees, clipped
    double gimbal_cmd_deg;
    if (f32) gimbal_cmd_deg = (double)__fmul_rn((float)act.u[0], sf.max_gimbal_rad) * (180.0 / PD_PI);
    else gimbal_cmd_deg = (act.u[0] * sd.max_gimbal_rad) * (180.0 / PD_PI);
    double x = prev.gimbal_deg;
    double gdeg = x + sd.dt_act * ((-x + gimbal_cmd_deg) / 1.0);
    const double gmax = sd.max_gimbal_deg;
    gdeg = gdeg < -gmax ? -gmax : (gdeg > gmax ? gmax : gdeg);
    double grad = gdeg * (PD_PI / 180.0);
    R sg, cg;
    m_sincos((R)grad, &sg, &cg);
    R t_par, t_perp, m_z;
    if (f32) {
        float thr = f32_throttle((float)act.u[1]);
        float thrust = __fmul_rn((float)((double)t_full * (double)c.n_eng), thr);
        float fpar = __fmul_rn(thrust, (float)cg);
        float fperp = __fmul_rn(-thrust, (float)sg);
        m_z = (R)((double)__fmul_rn(-thrust, (float)sg) * (double)d_thrust_cg);
        float total = __fsqrt_rn(__fadd_rn(__fmul_rn(fpar, fpar), __fmul_rn(fperp, fperp)));
        float n_tot = __fdiv_rn(total, (float)t_full);
        float mf = __fmul_rn(sf.te_over_vex, n_tot);
        t_par = (R)fpar;
        t_perp = (R)fperp;
        o.mass_flow = (R)mf;
        o.mass_flow_dt = (R)__fmul_rn(mf, sf.dt_phys);
        o.throttle = (R)thr;
    } else {
        R u1 = (R)act.u[1];
        R throttle = (u1 + R(1)) / R(2) * c.one_minus_nominal + c.nominal;
        R thrust = t_full * c.n_eng * throttle;
        t_par = thrust * cg;
        t_perp = -thrust * sg;
        m_z = -thrust * sg * d_thrust_cg;
        R total = m_sqrt(t_par * t_par + t_perp * t_perp);
        R n_tot = total / t_full;
        R mf = c.te_over_vex * n_tot;
        o.mass_flow = mf;
        o.mass_flow_dt = mf * c.dt_phys;
        o.throttle = throttle;
    }
    o.gimbal_deg = grad * (180.0 / PD_PI);
    // fin commands: u * radians(20) is called "deg" upstream and multiplied by 60 in ACS
    if (f32) {
        float cl60 = __fmul_rn(__fmul_rn((float)act.u[2], sf.max_defl_rad), 60.0f);
        float cr60 = __fmul_rn(__fmul_rn((float)act.u[3], sf.max_defl_rad), 60.0f);
        o.dl_cmd = (double)cl60 * (PD_PI / 180.0);
        o.dr_cmd = (double)cr60 * (PD_PI / 180.0);
    } else {
        o.dl_cmd = ((act.u[2] * sd.max_defl_rad) * 60.0) * (PD_PI / 180.0);
        o.dr_cmd = ((act.u[3] * sd.max_defl_rad) * 60.0) * (PD_PI / 180.0);
    }
    R f_perp, f_par, a_mz;
    acs<R>(alpha_eff, q, mach, x_cog, o.dl_cmd, o.dr_cmd, prev.dl, prev.dr, f_perp, f_par, a_mz);
    o.par = t_par + f_par;
    o.perp = t_perp + f_perp;
    o.mz = m_z + a_mz;
}

// subsonic / supersonic: force_moment_decomposer_ascent, rockets_physics.py:17-56 with
// max_gimbal 7 deg, nominal throttle 0.5, 16 gimballed + 26 fixed engines (:727-739).
// float32 action (fp64 build): gimbal angle, throttle, both thrusts, the parallel /
// perpendicular sums, total thrust and mass flow are float32; math.cos / math.sin return
// Python floats that NumPy casts back to float32; the moment is float32 * np.float64.
template <typename R>
__device__ __forceinline__ void control_ascent(const Action<2> &act, R p_atm, R d_thrust_cg, Control<R> &o) {
    const Scalars<R> &c = SC<R>();
    R t_full = c.T_e + (c.p_e - p_atm) * c.A_e;
    if (sizeof(R) == 8 && act.f32) {
        float grad = __fmul_rn((float)act.u[0], sf.max_gimbal_rad);
        double sgd, cgd;
        sincos((double)grad, &sgd, &cgd);
        float sg = (float)sgd, cg = (float)cgd;
        float thr = f32_throttle((float)act.u[1]);
        float thrust_g = __fmul_rn((float)((double)t_full * (double)c.n_eng), thr);
        float thrust_ng = __fmul_rn((float)((double)t_full * (double)c.n_eng_ng), thr);
        float fpar = __fadd_rn(thrust_ng, __fmul_rn(thrust_g, cg));
        float fperp = __fmul_rn(-thrust_g, sg);
        o.mz = (R)((double)__fmul_rn(-thrust_g, sg) * (double)d_thrust_cg);
        float total = __fsqrt_rn(__fadd_rn(__fmul_rn(fpar, fpar), __fmul_rn(fperp, fperp)));
        float n_tot = __fdiv_rn(total, (float)t_full);
        float mf = __fmul_rn(sf.te_over_vex, n_tot);
        o.par = (R)fpar;
        o.perp = (R)fperp;
        o.mass_flow = (R)mf;
        o.mass_flow_dt = (R)__fmul_rn(mf, sf.dt_phys);
        o.throttle = (R)thr;
        o.f32_forces = true;         // no np.float64 ACS term is added here, unlike the landing burns
    } else {
        R grad = (R)act.u[0] * c.max_gimbal_rad;
        R sg, cg;
        m_sincos(grad, &sg, &cg);
        R u1 = (R)act.u[1];
        R throttle = (u1 + R(1)) / R(2) * c.one_minus_nominal + c.nominal;
        R thrust_g = t_full * c.n_eng * throttle;
        R thrust_ng = t_full * c.n_eng_ng * throttle;
        R t_par = thrust_ng + thrust_g * cg;
        R t_perp = -thrust_g * sg;
        o.mz = -thrust_g * sg * d_thrust_cg;
        R total = m_sqrt(t_par * t_par + t_perp * t_perp);
        R n_tot = total / t_full;
        R mf = c.te_over_vex * n_tot;
        o.par = t_par;
        o.perp = t_perp;
        o.mass_flow = mf;
        o.mass_flow_dt = mf * c.dt_phys;
        o.throttle = throttle;
    }
}

// ballistic_arc_descent: RCS, rockets_physics.py:149-166.  Pure couple, no force, no mass flow.
// float32 action: the thruster force is float32, the moment arms are np.float64.
template <typename R>
__device__ __forceinline__ void control_rcs(const Action<1> &act, R x_cog, Control<R> &o) {
    const Scalars<R> &c = SC<R>();
    R F;
    if (sizeof(R) == 8 && act.f32) F = (R)__fmul_rn(sf.rcs_force, (float)act.u[0]);
    else F = c.rcs_force * (R)act.u[0];
    o.mz = -F * (x_cog - c.d_rcs_bottom) + F * (c.d_rcs_top - x_cog);
    o.par = R(0); o.perp = R(0);
    o.mass_flow = R(0); o.mass_flow_dt = R(0); o.throttle = R(0);
}

// flip_over_boostbackburn: force_moment_decomposer_flipoverboostbackburn, rockets_physics.py:63-92
// (max gimbal 10 deg, :759).  Gimbal command through a first-order low-pass (tau 1.0) on the ENV dt,
// throttle 1, gimballed engines only.  float32 action (NEP 50): the command and the filter run in
// float32 (0.1 and 1.0 are weak Python floats) and the filtered angle stays a float32 array in the
// env's memory; math.radians() then promotes, so thrust, moment and mass flow are float64.
template <typename R>
__device__ __forceinline__ void control_flip(const Action<1> &act, const ActPrev &prev, R p_atm, R d_thrust_cg,
                                             Control<R> &o) {
    const Scalars<R> &c = SC<R>();
    double gdeg;
    if (sizeof(R) == 8 && act.f32) {
        const float u = __fmul_rn((float)act.u[0], sf.flip_max_gimbal_deg);
        const float x = (float)prev.gimbal_deg;
        gdeg = (double)__fadd_rn(x, __fmul_rn(sf.dt_act, __fdiv_rn(__fadd_rn(-x, u), 1.0f)));
    } else {
        const double u = act.u[0] * sd.flip_max_gimbal_deg;
        const double x = prev.gimbal_deg;
        gdeg = x + sd.dt_act * ((-x + u) / 1.0);
    }
    const double grad = gdeg * (PD_PI / 180.0);
    R sg, cg;
    m_sincos((R)grad, &sg, &cg);
    R t_full = c.T_e + (c.p_e - p_atm) * c.A_e;
    R thrust = t_full * c.n_eng * R(1);
    o.par = thrust * cg;
    o.perp = -thrust * sg;
    o.mz = -thrust * sg * d_thrust_cg;
    R total = m_sqrt(o.par * o.par + o.perp * o.perp);
    R n_tot = total / t_full;
    R mf = c.te_over_vex * n_tot;
    o.mass_flow = mf;
    o.mass_flow_dt = mf * c.dt_phys;
    o.throttle = R(1);
    o.gimbal_deg = gdeg;
    o.dl_cmd = 0.0;
    o.dr_cmd = 0.0;
}

// landing_burn_pure_throttle_Pcontrol: force_moment_decomposer_landing_burn_throttle_PID,
// rockets_physics.py:402-451.  action = reference speed; Kp = -0.08; the inner throttle
// command is handed on as a *list*, which upstream unpacks with float(): whatever the dtype of
// v_ref, the thrust chain after it runs in float64.
template <typename R>
__device__ __forceinline__ void control_C(const Action<1> &act, R speed, R p_atm, R alpha_eff, R q,
                                          R x_cog, R mach, Control<R> &o) {
    Action<1> inner;
    inner.f32 = false;
    if (sizeof(R) == 8 && act.f32) {
        float err = __fsub_rn((float)act.u[0], (float)speed);
        float nn = __fmul_rn(err, -0.08f);
        nn = nn < 0.f ? 0.f : (nn > 1.f ? 1.f : nn);
        inner.u[0] = (double)__fmul_rn(2.0f, __fsub_rn(nn, 0.5f));
    } else {
        R err = (R)act.u[0] - speed;
        R nn = err * R(-0.08);
        nn = nn < R(0) ? R(0) : (nn > R(1) ? R(1) : nn);
        inner.u[0] = (double)(R(2) * (nn - R(0.5)));
    }
    control_P<R>(inner, p_atm, alpha_eff, q, x_cog, mach, o);
}

// ------------------------------------------------------------------ one Euler sub-step
// RT = accumulation type of the RBF dot products.
template <typename R, typename RT, int PHASE, bool WIND, int COOP = 1, bool FULL = false>
__device__ __forceinline__ void substep(State &s, const Action<phase_adim(PHASE)> &act,
                                        const ActPrev &prev, WindState &w, const WindCtx &wc,
                                        unsigned int env_id, Info<R, FULL> &info, Control<R> &ctl,
                                        const SharedTables *sh) {
    const Scalars<R> &c = SC<R>();
    R y = (R)s.y, vx = (R)s.vx, vy = (R)s.vy;
    const double theta_pre = s.theta;
    R rho, p_atm, a_snd;
    isa<R, PHASE != 1>(y, rho, p_atm, a_snd);
    R speed = m_sqrt(vx * vx + vy * vy);
    R mach = a_snd != R(0) ? m_min(m_div(speed, a_snd), R(10)) : R(0);
    R q = R(0.5) * rho * (speed * speed);
    R fuel = m_div(c.m_prop0 - (R)s.m_prop, c.m_prop0);
    if (fuel == R(0)) fuel = R(1e-6);
    R x_cog, inertia;
    if constexpr (phase_ascent(PHASE)) cog_inertia_full<R>(R(1) - fuel, x_cog, inertia);
    else cog_inertia<R>(R(1) - fuel, x_cog, inertia);
    R d_thrust_cg = x_cog + c.engine_height;
    R alpha_eff = s.vy < 0.0 ? (R)(s.gamma - s.theta - PD_PI) : (R)s.alpha;
    R d_cp_cg = x_cog - c.cop;
    R ug = R(0), vg = R(0);
    R f_wind_x = R(0);
    if (WIND) {
        wind_sample<R, COOP>(s.y, w, wc, env_id, ug, vg);
        f_wind_x = R(0.5) * rho * (ug * ug) * c.S_ref * c.c_gust_x;
    }
    R C_L = R(0), C_D = R(0);
    int status = 0;
    if (a_snd != R(0)) {
        aero_coefficients<R, COOP>(mach, alpha_eff, C_L, C_D, status, sh);
    }
    R qdyn = R(0.5) * rho * (speed * speed);
    R drag = qdyn * C_D * c.S_ref;
    R lift = qdyn * C_L * c.S_ref;
    R sa, ca;
    m_sincos(alpha_eff, &sa, &ca);
    R a_par, a_perp;
    if (s.vy >= 0.0) {
        a_par = lift * sa - drag * ca;
        a_perp = -lift * ca - drag * sa;
    } else {
        a_par = drag * ca - lift * sa;
        a_perp = -drag * sa - lift * ca;
    }
    R st, ct;
    if constexpr (PHASE != 1) m_sincos_pitch((R)s.theta, &st, &ct);
    else m_sincos((R)s.theta, &st, &ct);
    R aero_x = a_par * ct + a_perp * st;
    R aero_y = a_par * st - a_perp * ct;
    R aero_mz = a_perp * d_cp_cg;
    ctl.f32_forces = false;
    if constexpr (PHASE == 0)
        control_P<R>(act, p_atm, alpha_eff, q, x_cog, mach, ctl);
    else if constexpr (PHASE == 1)
        control_G<R>(act, prev, p_atm, d_thrust_cg, alpha_eff, q, x_cog, mach, ctl);
    else if constexpr (phase_ascent(PHASE))
        control_ascent<R>(act, p_atm, d_thrust_cg, ctl);
    else if constexpr (PHASE == 4)
        control_rcs<R>(act, x_cog, ctl);
    else if constexpr (PHASE == 6) {
        control_flip<R>(act, prev, p_atm, d_thrust_cg, ctl);
        // "No aerodynamic forces in upper atmosphere, this is a redundancy." (rockets_physics.py:556-560;
        // C_L and C_D are still evaluated and reported)
        aero_x = R(0); aero_y = R(0); aero_mz = R(0);
    } else
        control_C<R>(act, speed, p_atm, alpha_eff, q, x_cog, mach, ctl);
    R c_par = ctl.par, c_perp = ctl.perp, c_mz = ctl.mz;
    // NaN guards are an if/elif chain upstream: only the first NaN is cleared
    if (c_par != c_par) c_par = R(0);
    else if (c_perp != c_perp) c_perp = R(0);
    else if (c_mz != c_mz) c_mz = R(0);
    R c_x = c_par * ct + c_perp * st;
    R c_y = c_par * st - c_perp * ct;
    R fx = aero_x + c_x + f_wind_x;
    R fy = aero_y + c_y;
    if (sizeof(R) == 8 && ctl.f32_forces) {
        // rockets_physics.py:608-616 with np.float32 control forces: math.cos/sin(theta) are cast
        // to float32, and so is the aerodynamic force - a Python float (weak) here, since the
        // RBF coefficients come back as Python floats - before the float32 sum; the wind force
        // (np.float64) then promotes the total
        const float pf = (float)c_par, qf = (float)c_perp, cf = (float)ct, stf = (float)st;
        const float cxf = __fadd_rn(__fmul_rn(pf, cf), __fmul_rn(qf, stf));
        const float cyf = __fsub_rn(__fmul_rn(pf, stf), __fmul_rn(qf, cf));
        fx = (R)__fadd_rn((float)aero_x, cxf) + f_wind_x;
        fy = (R)__fadd_rn((float)aero_y, cyf);
    }
    const R RE = R(6371000.0);
    R gr = m_div(RE, RE + y);
    R g = R(9.80665) * (gr * gr);
    R mass = (R)s.mass;
    R vx_dot = m_div(fx, mass);
    R vy_dot = m_div(fy, mass) - g;
    const double dt = sd.dt_phys;
    const double vx_before = s.vx, vy_before = s.vy;
    s.vx += (double)(vx_dot * c.dt_phys);
    s.vy += (double)(vy_dot * c.dt_phys);
    s.x += s.vx * dt;
    s.y += s.vy * dt;
    R mz = c_mz + aero_mz;
    R tdd = m_div(mz, inertia);
    s.theta_dot += (double)(tdd * c.dt_phys);
    s.theta += s.theta_dot * dt;
    double gam;
    bool turned = false;
    if constexpr (sizeof(R) == 4) {
        // fp32 build: atan2 is 11-12 % of the dependent instruction chain that bounds a lone episode
        // and, with the aero patches, the step kernel.  The velocity turns by a few mrad per 25 ms sub-step, so
        // gamma advances by atan(cross / dot) of the old and new velocity, |t| < 2^-5: a
        // float-seeded Newton reciprocal and five series terms, 1e-16 rad from atan2 per
        // sub-step against the 1e-7 relative rounding of the fp32 forces.
        const double cross = vx_before * s.vy - vy_before * s.vx;
        const double dot = vx_before * s.vx + vy_before * s.vy;
        if (dot > 1e-30 && dot < 1e30 && fabs(cross) < 0.03125 * dot) {
            double r = (double)__frcp_rn((float)dot);
            r = r * (2.0 - dot * r);
            r = r * (2.0 - dot * r);
            const double t = cross * r, t2 = t * t;
            double poly = fma(t2, 1.0 / 9.0, -1.0 / 7.0);
            poly = fma(poly, t2, 1.0 / 5.0);
            poly = fma(poly, t2, -1.0 / 3.0);
            gam = s.gamma + fma(t * t2, poly, t);
            if (gam >= PD_TWO_PI) gam -= PD_TWO_PI;
            turned = true;
        }
    }
    if (!turned) gam = atan2(s.vy, s.vx);
    if (s.theta > PD_TWO_PI) s.theta -= PD_TWO_PI;
    if (gam < 0.0) gam = PD_TWO_PI + gam;
    s.gamma = gam;
    s.alpha = s.theta - gam;
    s.m_prop -= (double)ctl.mass_flow_dt;
    s.mass -= (double)ctl.mass_flow_dt;
    s.time += dt;
    info.mach = mach; info.q = q; info.CL = C_L; info.CD = C_D; info.rho = rho;
    info.p_atm = p_atm; info.a = a_snd; info.x_cog = x_cog; info.inertia = inertia;
    info.mass_flow = ctl.mass_flow; info.throttle = ctl.throttle; info.alpha_eff = alpha_eff;
    info.ug = ug; info.vg = vg;
    info.rbf_status |= status;
    if constexpr (FULL) {
        R *x = info.x;
        x[0] = drag; x[1] = lift; x[2] = d_cp_cg; x[3] = d_thrust_cg; x[4] = fuel;
        x[5] = c_par; x[6] = c_perp; x[7] = c_x; x[8] = c_y; x[9] = aero_x; x[10] = aero_y;
        x[11] = g; x[12] = c_mz; x[13] = aero_mz; x[14] = mz; x[15] = tdd; x[16] = vx_dot;
        x[17] = vy_dot; x[18] = f_wind_x;
        x[19] = (PHASE == 1 || PHASE == 6) ? (R)ctl.gimbal_deg : (phase_ascent(PHASE) ? (R)(act.u[0] * sd.max_gimbal_rad * (180.0 / PD_PI)) : R(0));
        x[20] = PHASE == 1 ? (R)ctl.dl_cmd : R(0);
        x[21] = PHASE == 1 ? (R)ctl.dr_cmd : R(0);
        const R qmax = PHASE <= 1 || PHASE == 5 ? R(65000) : R(30000);
        x[22] = a_snd != R(0) ? m_sqrt(R(2) * qmax / rho) * R(1) / a_snd : R(200);
        x[23] = (R)theta_pre;
        x[24] = gridfin_ca<R>(mach);
        x[25] = gridfin_cn_alpha<R>(mach);
        x[26] = PHASE == 1 ? (R)(prev.dl + sd.dt_act * ((-prev.dl + ctl.dl_cmd) / 0.5)) : R(0);
        x[27] = PHASE == 1 ? (R)(prev.dr + sd.dt_act * ((-prev.dr + ctl.dr_cmd) / 0.5)) : R(0);
#pragma unroll
        for (int k = 28; k < PD_INFO_X; ++k) x[k] = R(0);
    }
}

// ------------------------------------------------------------------ g-load window

template <typename R>
__device__ __forceinline__ R gwindow_push(GWindow<R> &g, R g_load) {
    if (g.n < 10) {
#pragma unroll
        for (int i = 0; i < 10; ++i)
            if (i == g.n) g.w[i] = g_load;
        g.n += 1;
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) g.w[i] = g.w[i + 1];
        g.w[9] = g_load;
    }
    R sum = R(0);
#pragma unroll
    for (int i = 0; i < 10; ++i)
        if (i < g.n) sum += g.w[i];
    return sum / R(10);       // divided by the window length even while it is filling
}

// ------------------------------------------------------------------ reward / truncation / done

template <typename R>
__device__ __forceinline__ R overshoot(double x, double y) {
    if (x < 0.0 && y < 0.0) return m_sqrt((R)x * (R)x + (R)y * (R)y);
    if (x < 0.0) return (R)(-x);
    if (y < 0.0) return (R)(-y);
    return R(0);
}

// scipy interp1d(kind='linear', fill_value='extrapolate') on a sorted grid with precomputed
// segment slopes: idx = clip(searchsorted(x, v, 'left'), 1, n-1); y = s[idx-1] (v - x[idx-1]) + y[idx-1]
__device__ __forceinline__ double interp_global(const double *x, const double *y, const double *sl, int n, double v) {
    const int lo = seg_index(x, n, v);
    return __ldg(sl + lo) * (v - __ldg(x + lo)) + __ldg(y + lo);
}
template <typename R>
__device__ __forceinline__ R hyper(int row, R mach) {
    const Scalars<R> &c = SC<R>();
    int a = 0;
#pragma unroll
    for (int j = 0; j < 12; ++j) a += (c.hyp_m[j] < mach) ? 1 : 0;
    const int idx = a < 1 ? 1 : (a > 11 ? 11 : a);
    const int lo = idx - 1;
    return c.hyp_s[row][lo] * (mach - c.hyp_m[lo]) + c.hyp_v[row][lo];
}

// subsonic / supersonic closures, rtd_rl.py:11-114 (NaN states: truncated with id 0, reward 0)
template <typename R>
__device__ __forceinline__ void rtd_ascent(const State &s, Rtd<R> &o) {
    const Scalars<R> &c = SC<R>();
    const bool nan = (s.x != s.x) || (s.y != s.y) || (s.vx != s.vx) || (s.vy != s.vy) ||
                     (s.theta != s.theta) || (s.theta_dot != s.theta_dot) || (s.gamma != s.gamma) ||
                     (s.alpha != s.alpha) || (s.mass != s.mass) || (s.m_prop != s.m_prop) || (s.time != s.time);
    if (nan) { o.reward = R(0); o.done = 0; o.truncated = 1; o.trunc_id = 0; return; }
    R rho, p_atm, a_snd;
    isa<R>((R)s.y, rho, p_atm, a_snd);
    R vx = (R)s.vx, vy = (R)s.vy;
    R speed = m_sqrt(vx * vx + vy * vy);
    R mach = (speed != R(0) && a_snd != R(0)) ? speed / a_snd : R(0);
    const R xr = (R)interp_global(tb.ref_y, tb.ref_v[0], tb.ref_s[0], tb.n_ref, s.y);
    const R vxr = (R)interp_global(tb.ref_y, tb.ref_v[1], tb.ref_s[1], tb.n_ref, s.y);
    const R vyr = (R)interp_global(tb.ref_y, tb.ref_v[2], tb.ref_s[2], tb.n_ref, s.y);
    const R max_x = hyper<R>(0, mach), max_vy = hyper<R>(1, mach), max_vx = hyper<R>(2, mach),
            max_al = hyper<R>(3, mach);
    R ex = (R)s.x - xr, evx = vx - vxr, evy = vy - vyr;
    int tr = 0, id = 0;
    if (s.m_prop <= 0.0) { tr = 1; id = 1; }
    else if (mach > c.terminal_mach + R(0.09)) { tr = 1; id = 2; }
    else if (m_abs(ex) > max_x) { tr = 1; id = 3; }
    else if (s.y < 0.0) { tr = 1; id = 4; }
    else if ((R)fabs(s.alpha) > max_al * R(PD_PI / 180.0)) { tr = 1; id = 5; }
    else if (m_abs(evx) > max_vx) { tr = 1; id = 6; }
    else if (m_abs(evy) > max_vy) { tr = 1; id = 7; }
    const int dn = (s.m_prop >= 0.0 && mach > c.terminal_mach) ? 1 : 0;
    R reward = R(0);
    if (!(s.y < 0.0)) {
        const R adeg = (R)s.alpha * R(180.0 / PD_PI);
        reward += m_exp(R(-4) * (evx * evx) / (max_vx * max_vx)) * R(100);
        reward += m_exp(R(-4) * (evy * evy) / (max_vy * max_vy)) * R(100);
        reward += m_exp(R(-4) * (ex * ex) / (max_x * max_x)) * R(100);
        reward += m_exp(R(-4) * (adeg * adeg) / (max_al * max_al)) * R(100);
        if (dn) reward += R(2.5);
        reward /= R(10000);
    }
    o.reward = reward; o.done = dn; o.truncated = tr; o.trunc_id = id;
}

// ballistic_arc_descent closures, rtd_rl.py:147-188
template <typename R>
__device__ __forceinline__ void rtd_ballistic(const State &s, Rtd<R> &o) {
    R vx = (R)s.vx, vy = (R)s.vy;
    R speed = m_sqrt(vx * vx + vy * vy);
    R q = R(0.5) * isa_rho<R>((R)s.y) * (speed * speed);
    const double ae = fabs(s.gamma - s.theta - PD_PI);
    const int dn = (q > R(10000) && ae < 3.0 * (PD_PI / 180.0)) ? 1 : 0;
    const int tr = (q > R(10000 - 2000) && ae > 5.0 * (PD_PI / 180.0)) ? 1 : 0;
    R reward = (R)((PD_PI - ae) / PD_PI);
    if (dn) reward += R(3.5);
    o.reward = reward / R(100); o.done = dn; o.truncated = tr; o.trunc_id = tr;
}

// landing_burn_pure_throttle_Pcontrol closures, rtd_rl.py:353-401 and the second (winning)
// reward definition :478-531.  v_ref / vref_f32: the action as the base env received it.
template <typename R>
__device__ __forceinline__ void rtd_pcontrol(const State &s, R g1, double v_ref, bool vref_f32, Rtd<R> &o) {
    const Scalars<R> &c = SC<R>();
    R vx = (R)s.vx, vy = (R)s.vy;
    R speed = m_sqrt(vx * vx + vy * vy);
    R rho = isa_rho<R>((R)s.y);
    R q = R(0.5) * rho * (speed * speed);
    const double theta_lim = PD_PI + 2.0 * (PD_PI / 180.0);
    int tr = 0, id = 0;
    if (s.y < -10.0) { tr = 1; id = 1; }
    else if (s.m_prop <= 0.0) { tr = 1; id = 2; }
    else if (s.theta > theta_lim) { tr = 1; id = 3; }
    else if (q > R(65000)) { tr = 1; id = 4; }
    else if (g1 > R(6.0)) { tr = 1; id = 5; }
    else if (s.vy > 0.0) { tr = 1; id = 6; }
    const int dn = (s.y > 0.0 && s.y < 5.0 && speed < R(1)) ? 1 : 0;
    R sp = m_hypot(vx, vy);
    R qq = R(0.5) * rho * (sp * sp);
    R r = R(0);
    if (qq > R(60000)) {
        R e = (qq - R(60000)) / R(5000);
        r -= m_min(e * e, R(1));
    }
    if (g1 > R(5.5)) {
        R e = (g1 - R(5.5)) / R(0.5);
        r -= m_min(e * e, R(1));
    }
    R prog = (R)((sd.y0 - s.y) / sd.y0);
    R track;
    if (sizeof(R) == 8 && vref_f32) {
        // np.float32 arithmetic: abs(speed - v_ref) / 10.0, 1.0 - ..., then max(0.0, .)
        float t = __fsub_rn(1.0f, __fdiv_rn(fabsf(__fsub_rn((float)sp, (float)v_ref)), 10.0f));
        track = t > 0.f ? (R)t : R(0);
    } else {
        R t = R(1) - m_abs(sp - (R)v_ref) / R(10);
        track = t > R(0) ? t : R(0);
    }
    R wp = (qq <= R(60000) && g1 <= R(5.5)) ? R(0.5) : R(0.5 * 0.1);
    r += wp * prog * track;
    if (s.y < 100.0) {
        R sl = R(1) - m_abs(vy) / R(50);
        r += R(0.5) * (sl > R(0) ? sl : R(0));
    }
    r += c.alive_bonus;
    if (dn && !tr) {
        r += R(5);
        R used = (R)(sd.y0 * 0.0 + (sd.mass0 - s.mass));
        r -= m_min(R(0.1) * used, R(1));
    } else if (tr) {
        r -= m_min(R(4) * (R)(s.y / sd.y0) * (m_abs(vy) / R(100)), R(5));
    }
    r = r < R(-10) ? R(-10) : (r > R(10) ? R(10) : r);
    o.reward = r; o.done = dn; o.truncated = tr; o.trunc_id = id;
}

// type = 'supervisory' closures, src/envs/supervisory/rtd_supervisory_mock.py:6-104: done and
// truncation per phase, reward 0.  (Upstream's g-load branch raises NameError when it fires;
// here it truncates with id 5.)
template <typename R, int PHASE>
__device__ __forceinline__ void rtd_supervisory(const State &s, R g1, Rtd<R> &o) {
    const Scalars<R> &c = SC<R>();
    R rho, p_atm, a_snd;
    isa<R>((R)s.y, rho, p_atm, a_snd);
    R vx = (R)s.vx, vy = (R)s.vy;
    R speed = m_sqrt(vx * vx + vy * vy);
    R q = R(0.5) * rho * (speed * speed);
    const double ae = fabs(s.gamma - s.theta - PD_PI);
    int dn = 0, tr = 0, id = 0;
    if constexpr (PHASE == 2) {
        dn = speed / a_snd > R(1.1);
        if (s.m_prop <= 0.0) { tr = 1; id = 1; }
    } else if constexpr (PHASE == 3) {
        dn = s.y > (double)sd.sup_terminal_alt;
        if (s.m_prop <= 0.0) { tr = 1; id = 1; }
    } else if constexpr (PHASE == 4) {
        dn = q > R(65000) && ae < 3.0 * (PD_PI / 180.0);
        if (q > R(35000) && ae > 3.0 * (PD_PI / 180.0)) { tr = 1; id = 1; }
    } else if constexpr (PHASE == 6) {
        dn = s.vx < -60.0;          // flip_over_boostbackburn_terminal_vx, rtd_supervisory_mock.py:14, 34-38
        if (s.m_prop <= 0.0) { tr = 1; id = 1; }
    } else {
        dn = s.y < 1.0;
        if (s.y < -10.0) { tr = 1; id = 1; }
        else if (s.m_prop <= 0.0) { tr = 1; id = 2; }
        else if (s.theta > PD_PI + 2.0 * (PD_PI / 180.0)) { tr = 1; id = 3; }
        else if (q > R(65000)) { tr = 1; id = 4; }
        else if (g1 > R(6.0)) { tr = 1; id = 5; }
        else if (s.vy > 0.0) { tr = 1; id = 6; }
    }
    o.reward = R(0); o.done = dn; o.truncated = tr; o.trunc_id = id;
}

template <typename R, int PHASE, int RTD>
__device__ __forceinline__ void rtd_eval(const State &s, R g1, R u0, Rtd<R> &o) {
    const Scalars<R> &c = SC<R>();
    R vx = (R)s.vx, vy = (R)s.vy;
    R speed = m_sqrt(vx * vx + vy * vy);
    R rho = isa_rho<R>((R)s.y);
    R q = R(0.5) * rho * (speed * speed);
    R a_eff = s.vy < 0.0 ? (R)fabs(s.gamma - s.theta - PD_PI) : (R)fabs(s.theta - s.gamma);
    const double theta_lim = PD_PI + 2.0 * (PD_PI / 180.0);
    int tr = 0, id = 0, dn = 0;
    R reward = R(0);
    if (RTD == 0 && PHASE == 0) {
        if (s.y < 0.0) { tr = 1; id = 1; }
        else if (s.m_prop <= 0.0) { tr = 1; id = 2; }
        else if (s.theta > theta_lim) { tr = 1; id = 3; }
        else if (q > R(65000)) { tr = 1; id = 4; }
        else if (s.vy > 0.0) { tr = 1; id = 6; }
        else if (g1 > R(6.0)) { tr = 1; id = 7; }
        dn = (s.y > 0.0 && s.y < 1.0 && speed < R(5.5)) ? 1 : 0;
        if (tr && s.y > 0.0) reward = -(R)fabs(s.y);
        else if (tr && s.y < 0.0) reward = R(200) - speed;
        else if (dn) reward = (R)s.m_prop;
    } else if (RTD == 0) {
        R over = overshoot<R>(s.x, s.y);
        R dist = m_sqrt((R)s.x * (R)s.x + (R)s.y * (R)s.y);
        if (over > R(0.5)) { tr = 1; id = 1; }
        else if (s.m_prop <= 0.0) { tr = 1; id = 2; }
        else if (a_eff > R(10.0 * (PD_PI / 180.0))) { tr = 1; id = 3; }
        else if (q > R(65000)) { tr = 1; id = 4; }
        else if (s.vy > 0.0) { tr = 1; id = 6; }
        else if (g1 > R(6.0)) { tr = 1; id = 7; }
        else if (s.y > 1000.0 && s.vx > 0.0) { tr = 1; id = 8; }
        dn = (dist > R(0) && dist < R(1) && speed < R(2.5)) ? 1 : 0;
        if (tr && over < R(0.5)) reward = -dist;
        else if (tr) reward = R(200) - speed;
        else if (dn) reward = (R)s.m_prop;
    } else {
        // rl closures are shared by both landing phases (rtd_rl.py:194-240)
        if (s.y < -10.0) { tr = 1; id = 1; }
        else if (s.m_prop <= 0.0) { tr = 1; id = 2; }
        else if (s.theta > theta_lim) { tr = 1; id = 3; }
        else if (q > R(65000)) { tr = 1; id = 4; }
        else if (g1 > R(6.0)) { tr = 1; id = 5; }
        else if (s.vy > 0.0) { tr = 1; id = 6; }
        else if (s.vx > 0.01) { tr = 1; id = 7; }
        dn = (s.y > 0.0 && s.y < 1.0 && speed < R(5.0)) ? 1 : 0;
        if (PHASE == 0) {
            // dense pure-throttle reward, rtd_rl.py:286-336 (speed via hypot upstream)
            R sp = m_hypot(vx, vy);
            R qq = R(0.5) * rho * (sp * sp);
            R r = R(0);
            if (qq > R(60000)) {
                R e = (qq - R(60000)) / R(5000);
                r -= m_min(e * e, R(1));
            }
            if (g1 > R(5.5)) {
                R e = (g1 - R(5.5)) / R(0.5);
                r -= m_min(e * e, R(1));
            }
            R prog = (R)((sd.y0 - s.y) / sd.y0);
            R wp = (qq <= R(60000) && g1 <= R(5.5)) ? R(0.5) : R(0.5 * 0.1);
            r += wp * prog;
            if (s.y < 100.0) r += R(5.5) * (R(1) - m_abs(vy) / R(50));
            if (dn && !tr) r += R(400) * (R)s.m_prop / c.mass0;
            else if (tr && s.y > 0.0) r -= R(50) * ((R)fabs(s.y) / c.y0);
            else if (tr && s.y < 0.0) r -= R(50) * (m_abs(vy) / R(10));
            if (!dn || !(tr && s.y < 0.0)) r = r < R(-10) ? R(-10) : (r > R(10) ? R(10) : r);
            reward = r;
        } else {
            // gimballed landing burn, rtd_rl.py:243-269
            R ae = (R)fabs(s.gamma - s.theta - PD_PI);
            R tau = (u0 + R(1)) / R(2);
            const R lmax = R(0.29941239026616734);    // math.log(1 + math.radians(20))
            R r = (R(1.5) - m_log(R(1) + ae) / lmax - tau * R(0.5)) * (R)(1.0 - s.y / sd.y0) * R(2) / R(3);
            if (s.y < 100.0) r += R(1) - m_tanh((speed - R(15)) / R(15));
            if (tr && s.y < 5.0) r += R(1) - m_tanh((speed - R(5)) / R(5));
            if (dn) r += R(5);
            reward = r * c.rl_reward_scale;
        }
    }
    o.reward = reward; o.done = dn; o.truncated = tr; o.trunc_id = id;
}

// An episode cut off by the caller's step cap.  The reference has no cap (env_wrapped_ea.py:209:
// `while not done_or_truncated`), so such a particle would fly on to a truncation; scoring the
// partial sum (0 for the pso closures, which only pay at a terminal state) would rank a stalling
// policy above every crashed one.  The final state is therefore scored with the closures' own
// *truncated* branch (rtd_pso.py:222-229 / 300-316), and the rollout reports truncation id -1.
template <typename R, int PHASE>
__device__ __forceinline__ R cap_reward(const State &s) {
    R vx = (R)s.vx, vy = (R)s.vy;
    R speed = m_sqrt(vx * vx + vy * vy);
    if (PHASE == 0) {
        if (s.y > 0.0) return -(R)fabs(s.y);
        if (s.y < 0.0) return R(200) - speed;
        return R(0);
    }
    R over = overshoot<R>(s.x, s.y);
    R dist = m_sqrt((R)s.x * (R)s.x + (R)s.y * (R)s.y);
    return over < R(0.5) ? -dist : R(200) - speed;
}

// ------------------------------------------------------------------ observations
// pso: env_wrapped_ea.py:97-123 ; rl: env_wrapped_rl_pytorch.py:42-47,167-202 (state is
// rounded to float32 first)
template <typename R, int PHASE, int RTD>
__device__ __forceinline__ void observe(const State &s, R *o) {
    const Scalars<R> &c = SC<R>();
    if constexpr (PHASE >= 2) {
        // env_wrapped_rl_pytorch.py:169-177, 199-201: the state is rounded to float32; phases 2..4
        // divide a float32 array in place by the float64 normalisation vector (float64 divide,
        // float32 result); P-control forms (1 - y/norm)*2 - 1 in float64
        if constexpr (phase_ascent(PHASE)) {
            const double v[8] = {s.x, s.y, s.vx, s.vy, s.theta, s.theta_dot, s.alpha, s.mass};
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = (R)(float)((double)(float)v[k] / sd.norm8[k]);
        } else if constexpr (PHASE == 4) {
            const double v[4] = {s.theta, s.theta_dot, s.gamma, s.alpha};
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = (R)(float)((double)(float)v[k] / sd.norm8[k]);
        } else if constexpr (PHASE == 6) {
            const double v[2] = {s.theta, s.theta_dot};
#pragma unroll
            for (int k = 0; k < 2; ++k) o[k] = (R)(float)((double)(float)v[k] / sd.norm8[k]);
        } else {
            o[0] = (R)((1.0 - (double)(float)s.y / sd.norm_y) * 2 - 1);
        }
        return;
    }
    if (RTD == 0) {
        if (PHASE == 0) {
            o[0] = (R)s.y / c.norm_y;
            o[1] = (R)s.vy / c.norm_vy;
        } else {
            o[0] = (R)s.x / c.norm_x;
            o[1] = (R)s.y / c.norm_y;
            o[2] = (R)s.vx / c.norm_vx;
            o[3] = (R)s.vy / c.norm_vy;
            o[4] = m_tanh(c.k_theta_pso * (R)(s.theta - PD_PI / 2));
        }
    } else {
        float yf = (float)s.y, vyf = (float)s.vy;
        // np.float32 scalar / np.float64 norm -> float64 arithmetic on fp32-rounded inputs
        if (PHASE == 0) {
            o[0] = (R)((1.0 - (double)yf / sd.norm_y) * 2 - 1);
            o[1] = (R)((1.0 - (double)vyf / sd.norm_vy) * 2 - 1);
        } else {
            float th = (float)s.theta, thd = (float)s.theta_dot, gm = (float)s.gamma;
            o[0] = (R)((double)yf / sd.norm_y);
            o[1] = (R)((double)vyf / sd.norm_vy);
            // np.float32 scalar (-, *) Python float stays float32 (NEP 50); math.tanh promotes
            o[2] = (R)tanh((double)__fmul_rn(sf.k_theta_rl, __fsub_rn(th, (float)(PD_PI / 2))));
            o[3] = (R)tanh((double)__fmul_rn(sf.k_thetadot_rl, thd));
            o[4] = (R)tanh((double)__fmul_rn(sf.k_theta_rl, __fsub_rn(gm, (float)(1.5 * PD_PI))));
        }
    }
}

// ------------------------------------------------------------------ one env.step()
// 4 sub-steps with the same action (and, for G, the same actuator memory), g-load window,
// truncation -> done -> reward on the new state.  base_environment.py:99-154.
template <typename R, typename RT, int PHASE, int RTD, bool WIND, int COOP = 1, bool FULL = false>
__device__ __forceinline__ void env_step(State &s, const Action<phase_adim(PHASE)> &act,
                                         ActPrev &prev, WindState &w, const WindCtx &wc,
                                         unsigned int env_id, GWindow<R> &gw, Info<R, FULL> &info,
                                         Rtd<R> &out, R &g1_out, const SharedTables *sh) {
    R vxp = (R)s.vx, vyp = (R)s.vy;
    R v_p = m_sqrt(vxp * vxp + vyp * vyp);
    Control<R> ctl;
    if constexpr (WIND) gust_ahead<phase_nsub(PHASE), COOP>(w, wc, env_id);
#pragma unroll 1
    for (int k = 0; k < phase_nsub(PHASE); ++k)
        substep<R, RT, PHASE, WIND, COOP, FULL>(s, act, prev, w, wc, env_id, info, ctl, sh);
    if (phase_has_actuator_memory(PHASE)) {
        prev.gimbal_deg = ctl.gimbal_deg;
        prev.dl = ctl.dl_cmd;
        prev.dr = ctl.dr_cmd;
    }
    R vx = (R)s.vx, vy = (R)s.vy;
    R v = m_sqrt(vx * vx + vy * vy);
    R g_load = m_abs(v - v_p) / R(0.1) * R(1) / R(9.81);
    R g1 = gwindow_push<R>(gw, g_load);
    g1_out = g1;
    if constexpr (phase_ascent(PHASE)) {
        rtd_ascent<R>(s, out);
    } else if constexpr (PHASE == 4) {
        rtd_ballistic<R>(s, out);
    } else if constexpr (PHASE == 5) {
        rtd_pcontrol<R>(s, g1, act.u[0], act.f32, out);
    } else if constexpr (PHASE == 6) {
        rtd_supervisory<R, 6>(s, g1, out);     // the one closure set that runs this phase upstream
    } else {
        R u0 = (R)act.u[0];
        if (sizeof(R) == 8 && act.f32) u0 = (R)(float)act.u[0];
        rtd_eval<R, PHASE, RTD>(s, g1, u0, out);
    }
}

__device__ __forceinline__ void state_reset(State &s) {
    const double *i = tb.init;
    s.x = i[0]; s.y = i[1]; s.vx = i[2]; s.vy = i[3]; s.theta = i[4]; s.theta_dot = i[5];
    s.gamma = i[6]; s.alpha = i[7]; s.mass = i[8]; s.m_prop = i[9]; s.time = i[10];
}

};  // struct Dev

}  // namespace pd
