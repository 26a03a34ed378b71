// pd_kernels.cuh - kernels of the powered-descent hot path, templated on the compute type.
// Included by pd_fp64.cu (R = double) and pd_fp32.cu (R = float); each TU owns its copy of
// the __constant__ parameter blocks and exports an Impl table to pd_api.cu.
#pragma once
#include <mutex>
#include <set>
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "pd_device.cuh"
#include "pd_impl.h"

namespace pd {

// ------------------------------------------------------------------ SoA views
__device__ __forceinline__ void load_state(const EnvSoA &e, int i, State &s) {
    const size_t B = e.n;
    const double *p = e.st + i;
    s.x = p[0]; s.y = p[B]; s.vx = p[2 * B]; s.vy = p[3 * B]; s.theta = p[4 * B];
    s.theta_dot = p[5 * B]; s.gamma = p[6 * B]; s.alpha = p[7 * B]; s.mass = p[8 * B];
    s.m_prop = p[9 * B]; s.time = p[10 * B];
}
__device__ __forceinline__ void store_state(const EnvSoA &e, int i, const State &s) {
    const size_t B = e.n;
    double *p = e.st + i;
    p[0] = s.x; p[B] = s.y; p[2 * B] = s.vx; p[3 * B] = s.vy; p[4 * B] = s.theta;
    p[5 * B] = s.theta_dot; p[6 * B] = s.gamma; p[7 * B] = s.alpha; p[8 * B] = s.mass;
    p[9 * B] = s.m_prop; p[10 * B] = s.time;
}
template <typename R>
__device__ __forceinline__ void load_window(const EnvSoA &e, int i, GWindow<R> &g) {
    g.n = e.gwin_n[i];
#pragma unroll
    for (int k = 0; k < 10; ++k) g.w[k] = (R)e.gwin[(size_t)k * e.n + i];
}
template <typename R>
__device__ __forceinline__ void store_window(const EnvSoA &e, int i, const GWindow<R> &g) {
    e.gwin_n[i] = g.n;
#pragma unroll
    for (int k = 0; k < 10; ++k) e.gwin[(size_t)k * e.n + i] = (double)g.w[k];
}
__device__ __forceinline__ void load_wind(const EnvSoA &e, int i, WindState &w) {
    const size_t B = e.n;
    w.xu0 = e.wst[i]; w.xu1 = e.wst[B + i]; w.xv0 = e.wst[2 * B + i]; w.xv1 = e.wst[3 * B + i];
    w.sigma_u = e.wst[4 * B + i]; w.sigma_v = e.wst[5 * B + i];
    w.ctr = e.wctr[i];
    w.episode = e.episode[i];
}
__device__ __forceinline__ void store_wind(const EnvSoA &e, int i, const WindState &w) {
    const size_t B = e.n;
    e.wst[i] = w.xu0; e.wst[B + i] = w.xu1; e.wst[2 * B + i] = w.xv0; e.wst[3 * B + i] = w.xv1;
    e.wst[4 * B + i] = w.sigma_u; e.wst[5 * B + i] = w.sigma_v;
    e.wctr[i] = w.ctr;
}
// rl_wrapped_env_pytorch.augment_action for landing_burn (env_wrapped_rl_pytorch.py:144-157)
// `1 + c*abs(u)` is evaluated in float32 when the action is a float32 ndarray (int * np.float32
// stays float32 under NEP 50); math.log then works on the promoted double.
__device__ __forceinline__ double log_compress(double u, double cfac, bool f32) {
    double arg = f32 ? (double)__fadd_rn(1.0f, __fmul_rn((float)cfac, fabsf((float)u)))
                     : 1.0 + cfac * fabs(u);
    return copysign(log(arg) / log(1.0 + cfac), u);
}

template <int A>
__device__ __forceinline__ void read_action(const void *actions, int dtype, size_t idx, Action<A> &a) {
    a.f32 = (dtype == 1);
#pragma unroll
    for (int k = 0; k < A; ++k)
        a.u[k] = dtype == 1 ? (double)((const float *)actions)[idx * A + k]
                            : ((const double *)actions)[idx * A + k];
}

template <int PHASE, int RTD>
__device__ __forceinline__ void shape_action(const Dev &D, Action<phase_adim(PHASE)> &a) {
    if constexpr (PHASE == 5 && RTD == 1) {
        // rl_wrapped_env_pytorch.augment_action (env_wrapped_rl_pytorch.py:158-164):
        // v_ref = (u0 + 1)/2 * speed0, float32 arithmetic for a float32 action
        if (a.f32) a.u[0] = (double)__fmul_rn(__fdiv_rn(__fadd_rn((float)a.u[0], 1.0f), 2.0f), D.sf.speed0);
        else a.u[0] = (a.u[0] + 1.0) / 2.0 * D.sd.speed0;
    }
    if constexpr (PHASE == 1 && RTD == 1) {
        // np.array([...python floats...]) -> float64 action
        a.u[0] = log_compress(a.u[0], 10.0, a.f32);
        a.u[2] = log_compress(a.u[2], 5.0, a.f32);
        a.u[3] = log_compress(a.u[3], 5.0, a.f32);
        a.f32 = false;
    }
}

// small hot tables -> shared memory.  stage_tables: every thread of the block calls it (one
// __syncthreads); thread 0 then starts one TMA bulk copy of the host-replicated image.
// wait_tables: every thread that goes on to read the tables calls it once, as late as possible
// (state / window / action loads are issued in between).
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ SharedTables *aligned_tables(unsigned char *raw) {
    const unsigned a = smem_addr(raw);
    return reinterpret_cast<SharedTables *>(raw + (((a + PD_SH_ALIGN - 1) & ~(PD_SH_ALIGN - 1)) - a));
}
__device__ __forceinline__ void stage_tables(SharedTables *sh, const void *image) {
    const unsigned bar = smem_addr(&sh->bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(PD_SH_IMAGE_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_addr(sh)), "l"(image), "r"(PD_SH_IMAGE_BYTES), "r"(bar) : "memory");
    }
}
__device__ __forceinline__ void wait_tables(SharedTables *sh) {
    const unsigned bar = smem_addr(&sh->bar);
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    }
}

// One block per SM, at most PD_MAX_BLOCK = 448 threads = the 443 lanes per SM that BASELINE config 2
// (65 536 envs over 148 SMs) provides.  The register file is allocated for warp counts rounded up to
// a multiple of 4, so these blocks are capped at 65 536 / 512 = 128 registers per thread (ptxas
// reports 128; a 144-register build is refused at launch).
#ifndef PD_MAX_BLOCK
#define PD_MAX_BLOCK 448
#endif
// Block size so that the batch fills the SMs in whole waves of one block per SM:
// 65 536 envs / 148 SMs -> 448 threads x 147 blocks (14 warps on every SM).
static inline void big_block_config(long long lanes, int n_sm, int &threads, int &blocks) {
    long long waves = (lanes + (long long)n_sm * PD_MAX_BLOCK - 1) / ((long long)n_sm * PD_MAX_BLOCK);
    long long per = (lanes + n_sm * waves - 1) / (n_sm * waves);
    threads = (int)((per + 31) / 32 * 32);
    if (threads < 64) threads = 64;
    if (threads > PD_MAX_BLOCK) threads = PD_MAX_BLOCK;
    blocks = (int)((lanes + threads - 1) / threads);
}

// ------------------------------------------------------------------ reset kernel
template <typename R>
__global__ void reset_kernel(const __grid_constant__ KParams kp, EnvSoA e, const uint8_t *mask, WindCtx wc,
                             const double *sigma_uv) {
    Dev D(kp);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n) return;
    if (mask && !mask[i]) return;
    State s;
    D.state_reset(s);
    store_state(e, i, s);
    e.gwin_n[i] = 0;
#pragma unroll
    for (int k = 0; k < 10; ++k) e.gwin[(size_t)k * e.n + i] = 0.0;
    e.aprev[i] = 0.0; e.aprev[(size_t)e.n + i] = 0.0; e.aprev[2 * (size_t)e.n + i] = 0.0;
    unsigned int ep = e.episode[i] + 1;
    e.episode[i] = ep;
    WindState w;
    D.wind_reset(w, wc, (unsigned)i, ep, sigma_uv);
    store_wind(e, i, w);
    e.trunc_id[i] = 0;
    e.ep_steps[i] = 0;
}

// observation of the current state (first step of a collection run)
template <typename R, int PHASE, int RTD>
__global__ void observe_kernel(const __grid_constant__ KParams kp, EnvSoA e, R *obs) {
    constexpr int O = phase_odim(PHASE);
    Dev D(kp);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n) return;
    State s;
    load_state(e, i, s);
    R o[O];
    D.observe<R, PHASE, RTD>(s, o);
#pragma unroll
    for (int k = 0; k < O; ++k) obs[(size_t)i * O + k] = o[k];
}

// ------------------------------------------------------------------ fused step kernel
// physics (4 sub-steps) + g-window + truncation/done/reward + observation + auto-reset.
#ifndef PD_STEP_MIN_BLOCKS
#define PD_STEP_MIN_BLOCKS 8
#endif
// (448 threads: the register file is allocated for warp counts rounded up to a multiple of 4, so the
// cap is 65 536 / 512 = 128 registers per thread - a launch with __maxnreg__(144) is refused with
// "too many resources requested")
template <typename R, typename RT, int PHASE, int RTD, bool WIND, bool FULL = false>
__global__ void __launch_bounds__(PD_MAX_BLOCK, 1)
step_kernel(const __grid_constant__ KParams kp, EnvSoA e, StepIO io, WindCtx wc, const double *sigma_uv,
            int auto_reset) {
    constexpr int A = phase_adim(PHASE);
    constexpr int O = phase_odim(PHASE);
    Dev D(kp);
    extern __shared__ __align__(16) unsigned char pd_smem[];
    SharedTables &sh = *aligned_tables(pd_smem);
    stage_tables(&sh, kp.tb.sh_image);
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n) return;
    State s;
    load_state(e, i, s);
    GWindow<R> gw;
    load_window<R>(e, i, gw);
    ActPrev prev = {0.0, 0.0, 0.0};
    if (phase_has_actuator_memory(PHASE)) {
        prev.gimbal_deg = e.aprev[i]; prev.dl = e.aprev[(size_t)e.n + i];
        prev.dr = e.aprev[2 * (size_t)e.n + i];
    }
    WindState w = {};
    if (WIND) load_wind(e, i, w);
    Action<A> act;
    read_action<A>(io.actions, io.action_dtype, (size_t)i, act);
    if (!io.raw_actions) shape_action<PHASE, RTD>(D, act);
    Info<R, FULL> info;
    info.rbf_status = 0;
    Rtd<R> out;
    R g1;
    const int ep_steps_in = e.ep_steps[i];      // issued with the other state loads, consumed at the end
    wait_tables(&sh);
    D.env_step<R, RT, PHASE, RTD, WIND, 1, FULL>(s, act, prev, w, wc, (unsigned)i, gw, info, out, g1, &sh);
    if constexpr (RTD == 1) {
        if (io.supervisory) D.rtd_supervisory<R, PHASE>(s, g1, out);
    }
    if (info.rbf_status) atomicOr(e.status, info.rbf_status);
    R obs[O];
    D.observe<R, PHASE, RTD>(s, obs);
    if (io.obs) {
#pragma unroll
        for (int k = 0; k < O; ++k) ((R *)io.obs)[(size_t)i * O + k] = obs[k];
    }
    if (io.reward) ((R *)io.reward)[i] = out.reward;
    if (io.done) io.done[i] = (uint8_t)out.done;
    if (io.truncated) io.truncated[i] = (uint8_t)out.truncated;
    if (io.trunc_id) io.trunc_id[i] = out.trunc_id;
    if (io.dbg) {
        double *d = io.dbg + (size_t)i * (FULL ? 16 + PD_INFO_X : 16);
        if constexpr (FULL) {
#pragma unroll
            for (int k = 0; k < PD_INFO_X; ++k) d[16 + k] = (double)info.x[k];
        }
        d[0] = info.mach; d[1] = info.q; d[2] = info.CL; d[3] = info.CD; d[4] = info.rho;
        d[5] = info.p_atm; d[6] = info.a; d[7] = info.x_cog; d[8] = info.inertia;
        d[9] = info.mass_flow; d[10] = info.throttle; d[11] = info.alpha_eff; d[12] = g1;
        d[13] = info.ug; d[14] = info.vg; d[15] = (double)info.rbf_status;
    }
    e.trunc_id[i] = out.trunc_id;
    int ep_steps = ep_steps_in + 1;
    if (auto_reset && (out.done || out.truncated)) {
        D.state_reset(s);
        gw.n = 0;
#pragma unroll
        for (int k = 0; k < 10; ++k) gw.w[k] = R(0);
        prev.gimbal_deg = prev.dl = prev.dr = 0.0;
        unsigned int ep = e.episode[i] + 1;
        e.episode[i] = ep;
        if (WIND) D.wind_reset(w, wc, (unsigned)i, ep, sigma_uv);
        ep_steps = 0;
        D.observe<R, PHASE, RTD>(s, obs);
    }
    if (io.next_obs) {
#pragma unroll
        for (int k = 0; k < O; ++k) ((R *)io.next_obs)[(size_t)i * O + k] = obs[k];
    }
    e.ep_steps[i] = ep_steps;
    store_state(e, i, s);
    store_window<R>(e, i, gw);
    if (phase_has_actuator_memory(PHASE)) {
        e.aprev[i] = prev.gimbal_deg; e.aprev[(size_t)e.n + i] = prev.dl;
        e.aprev[2 * (size_t)e.n + i] = prev.dr;
    }
    if (WIND) store_wind(e, i, w);
}

// ------------------------------------------------------------------ per-particle actor MLP
// simple_actor (env_wrapped_ea.py:18-44): Linear(in,8)+ReLU, NH x [Linear(8,8)+ReLU],
// Linear(8,out)+Tanh in float32.  Parameters in named_parameters() order, transposed on the
// device to [param][particle] so that a warp's 32 lanes read 32 consecutive floats: a
// warp-level batch of 32 independent GEMVs, one particle per lane.
template <int IN, int OUT, int NH>
__device__ __forceinline__ void actor_mlp(const float *__restrict__ wT, size_t stride, size_t col,
                                          const float *obs, float *act) {
    constexpr int H = 8;
    float h[H], g[H];
    const float *p = wT + col;
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < IN; ++k) acc = fmaf(__ldg(p + (size_t)(j * IN + k) * stride), obs[k], acc);
        h[j] = acc;
    }
    p += (size_t)(H * IN) * stride;
#pragma unroll
    for (int j = 0; j < H; ++j) { h[j] = fmaxf(h[j] + __ldg(p + (size_t)j * stride), 0.f); }
    p += (size_t)H * stride;
#pragma unroll 1
    for (int l = 0; l < NH; ++l) {
#pragma unroll
        for (int j = 0; j < H; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < H; ++k) acc = fmaf(__ldg(p + (size_t)(j * H + k) * stride), h[k], acc);
            g[j] = acc;
        }
        p += (size_t)(H * H) * stride;
#pragma unroll
        for (int j = 0; j < H; ++j) h[j] = fmaxf(g[j] + __ldg(p + (size_t)j * stride), 0.f);
        p += (size_t)H * stride;
    }
#pragma unroll
    for (int j = 0; j < OUT; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < H; ++k) acc = fmaf(__ldg(p + (size_t)(j * H + k) * stride), h[k], acc);
        act[j] = acc;
    }
    p += (size_t)(OUT * H) * stride;
#pragma unroll
    for (int j = 0; j < OUT; ++j) act[j] = tanhf(act[j] + __ldg(p + (size_t)j * stride));
}

static __global__ void init_queue_kernel(int *q, int v) { *q = v; }

// weights [n_particles][n_params] -> wT [n_params][n_particles]
static __global__ void transpose_weights_kernel(const float *__restrict__ w, float *__restrict__ wT,
                                         int n_particles, int n_params) {
    __shared__ float tile[32][33];
    int p0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int p = p0 + r, k = k0 + threadIdx.x;
        tile[r][threadIdx.x] = (p < n_particles && k < n_params) ? w[(size_t)p * n_params + k] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = k0 + r, p = p0 + threadIdx.x;
        if (k < n_params && p < n_particles) wT[(size_t)k * n_particles + p] = tile[threadIdx.x][r];
    }
}

// ------------------------------------------------------------------ cooperative per-particle MLP
// The lanes of a cooperative group (COOP = 8 or 32) all hold the same episode.  The hidden width of
// simple_actor is 8, so lane j (mod 8) owns hidden unit j of every layer: its weight rows - IN + 1,
// NH x 9 and 9 floats - are fetched ONCE per episode into a thread-private column of shared memory,
// a layer is 8 FMAs + 8 shuffles, and the per-step reload of all 249 / 372 weights through L2 by
// every lane (37 % of a lone episode's step latency, profiles/r2_rollout_mode2_before.txt) is gone.
// The FMA order per unit is that of actor_mlp, so both paths return bit-identical actions.
template <int IN, int NH>
__host__ __device__ constexpr int mlp_words() { return (IN + 1) + NH * 9 + 9; }

template <int IN, int OUT, int NH>
__device__ __forceinline__ void coop_mlp_load(const float *__restrict__ wT, size_t stride, size_t col, float *wcol) {
    constexpr int H = 8;
    const int j = threadIdx.x & 7;
    const float *p = wT + col;
    int n = 0;
    const int nt = blockDim.x;
#pragma unroll
    for (int k = 0; k < IN; ++k) wcol[(n++) * nt] = __ldg(p + (size_t)(j * IN + k) * stride);
    wcol[(n++) * nt] = __ldg(p + (size_t)(H * IN + j) * stride);
    size_t off = (size_t)(H * IN + H);
#pragma unroll
    for (int l = 0; l < NH; ++l) {
#pragma unroll
        for (int k = 0; k < H; ++k) wcol[(n++) * nt] = __ldg(p + (off + j * H + k) * stride);
        wcol[(n++) * nt] = __ldg(p + (off + H * H + j) * stride);
        off += H * H + H;
    }
    const int o = j < OUT ? j : 0;            // output layer: unit j < OUT (the others repeat unit 0)
#pragma unroll
    for (int k = 0; k < H; ++k) wcol[(n++) * nt] = __ldg(p + (off + o * H + k) * stride);
    wcol[(n++) * nt] = __ldg(p + (off + OUT * H + o) * stride);
}

template <int IN, int OUT, int NH, int COOP>
__device__ __forceinline__ void coop_mlp_forward(const float *wcol, const float *obs, float *act) {
    constexpr int H = 8;
    const int nt = blockDim.x;
    const int lane = threadIdx.x & 31;
    const int base = lane & ~7;                // the 8 lanes that hold units 0..7 for this lane
    const unsigned gmask = (COOP == 32 ? 0xffffffffu : ((1u << COOP) - 1u)) << (lane & ~(COOP - 1));
    float h[H];
    int n = 0;
    {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < IN; ++k) acc = fmaf(wcol[(n++) * nt], obs[k], acc);
        const float u = fmaxf(acc + wcol[(n++) * nt], 0.f);
#pragma unroll
        for (int k = 0; k < H; ++k) h[k] = __shfl_sync(gmask, u, base + k);
    }
#pragma unroll
    for (int l = 0; l < NH; ++l) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < H; ++k) acc = fmaf(wcol[(n++) * nt], h[k], acc);
        const float u = fmaxf(acc + wcol[(n++) * nt], 0.f);
#pragma unroll
        for (int k = 0; k < H; ++k) h[k] = __shfl_sync(gmask, u, base + k);
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) acc = fmaf(wcol[(n++) * nt], h[k], acc);
    const float u = tanhf(acc + wcol[(n++) * nt]);
#pragma unroll
    for (int o = 0; o < OUT; ++o) act[o] = __shfl_sync(gmask, u, base + o);
}

// ------------------------------------------------------------------ persistent rollout kernel
// One episode per thread (or per cooperative group of COOP lanes) from reset to done/truncated:
// objective_function (env_wrapped_ea.py:200-222) for POLICY_MLP, an env.step loop over a tape for
// POLICY_TAPE, LandingBurn.run_closed_loop for POLICY_CLASSICAL.
// MODE bit 0: hand-off - an episode still running after io.handoff_steps is written to the output
//             continuation records instead of being finished.
// MODE bit 1: the episodes are the input continuation records of an earlier stage; the launch is
//             a no-op unless their number is in (io.run_if_gt, io.run_if_le].
// Ragged episode lengths (P: 101 .. 4000+ steps) are served in stages of growing cooperation:
// one lane per episode while there are more episodes than lanes, 8 lanes, then 32 lanes for the
// handful of long ones whose sequential latency bounds the generation.
// MAXB: block-size bound.  The last few stragglers run one warp per episode and at most MAXB / 32
// of them per SM, so the (32, *, 128) instantiations are compiled for 4 warps per SM and may use
// 255 registers instead of the 128 the 448-thread blocks are held to.
template <typename R, typename RT, int PHASE, int RTD, bool WIND, int POLICY, int COOP, int MODE = 0,
          int MAXB = PD_MAX_BLOCK>
__global__ void __launch_bounds__(MAXB, 1)
rollout_kernel(const __grid_constant__ KParams kp, RolloutIO io, WindCtx wc, const double *sigma_uv, int *status) {
    constexpr int A = phase_adim(PHASE);
    constexpr int O = phase_odim(PHASE);
    constexpr bool FROM_RECORDS = (MODE & 2) != 0, HANDOFF = (MODE & 1) != 0;
    constexpr bool COOP_MLP = POLICY == 0 && COOP >= 8;
    constexpr int NH = PHASE == 0 ? 3 : 4;
    int n_work = io.n_episodes;
    if constexpr (FROM_RECORDS) {
        const int cnt = min(*io.cont_count, io.cont_cap);
        if (cnt <= io.run_if_gt || cnt > io.run_if_le) return;       // another stage variant's job
        n_work = cnt;
    }
    Dev D(kp);
    extern __shared__ __align__(16) unsigned char pd_smem[];
    SharedTables &sh = *aligned_tables(pd_smem);
    float *wcol = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(&sh) + sizeof(SharedTables)) + threadIdx.x;
    stage_tables(&sh, kp.tb.sh_image);
    wait_tables(&sh);
    // Persistent lanes with a work queue: a lane (group) that finishes pulls the next episode
    // index instead of idling until the slowest lane of its warp is done.
    // First assignment strided over the blocks warp by warp: c < lanes / COOP pieces of work occupy
    // c * COOP / 32 warps spread over all SMs instead of filling the first few blocks to 14 warps.
    int e = (((int)(threadIdx.x >> 5) * (int)gridDim.x + (int)blockIdx.x) * 32 + (int)(threadIdx.x & 31)) / COOP;   // FROM_RECORDS: record index
    const bool writer = (threadIdx.x & (COOP - 1)) == 0;
    bool active = e < n_work;
    int eid = e;                                                  // episode the lane works on
    State s;
    GWindow<R> gw;
    ActPrev prev;
    WindState w = {};
    Info<R> info;
    info.rbf_status = 0;
    double total = 0.0;
    int t = 0;
    size_t col = 0;
    auto begin_episode = [&]() {
        if constexpr (FROM_RECORDS) {
            const size_t cap = (size_t)io.cont_cap;
            const double *d = io.cont_d + e;
            const int *ci = io.cont_i + e;
            eid = ci[0]; t = ci[cap]; gw.n = ci[2 * cap]; w.ctr = (unsigned)ci[3 * cap];
            s.x = d[0]; s.y = d[cap]; s.vx = d[2 * cap]; s.vy = d[3 * cap]; s.theta = d[4 * cap];
            s.theta_dot = d[5 * cap]; s.gamma = d[6 * cap]; s.alpha = d[7 * cap]; s.mass = d[8 * cap];
            s.m_prop = d[9 * cap]; s.time = d[10 * cap];
#pragma unroll
            for (int k = 0; k < 10; ++k) gw.w[k] = (R)d[(11 + k) * cap];
            prev.gimbal_deg = d[21 * cap]; prev.dl = d[22 * cap]; prev.dr = d[23 * cap];
            w.xu0 = d[24 * cap]; w.xu1 = d[25 * cap]; w.xv0 = d[26 * cap]; w.xv1 = d[27 * cap];
            w.sigma_u = d[28 * cap]; w.sigma_v = d[29 * cap];
            w.episode = io.generation + 1u;
            total = d[30 * cap];
            info.q = R(0);
        } else {
            eid = e;
            D.state_reset(s);
            gw.n = 0;
#pragma unroll
            for (int k = 0; k < 10; ++k) gw.w[k] = R(0);
            prev.gimbal_deg = prev.dl = prev.dr = 0.0;
            if (WIND) D.wind_reset(w, wc, (unsigned)eid, io.generation + 1u, sigma_uv);
            info.q = R(0);
            total = 0.0;
            t = 0;
        }
        col = (size_t)(eid / io.n_seeds);
        if constexpr (COOP_MLP) coop_mlp_load<O, A, NH>(io.wT, io.w_stride, col, wcol);
    };
    if (active) begin_episode();
    while (__any_sync(0xffffffffu, active)) {
        if (!active) continue;
        Action<A> act;
        bool stop = false;
        if constexpr (POLICY == 0) {
            R obs[O];
            D.observe<R, PHASE, 0>(s, obs);
            float of[O], af[A];
#pragma unroll
            for (int k = 0; k < O; ++k) of[k] = (float)obs[k];
            if constexpr (COOP_MLP) coop_mlp_forward<O, A, NH, COOP>(wcol, of, af);
            else actor_mlp<O, A, NH>(io.wT, io.w_stride, col, of, af);
            act.f32 = true;
#pragma unroll
            for (int k = 0; k < A; ++k) act.u[k] = (double)af[k];
            if (io.act_out && writer) {
#pragma unroll
                for (int k = 0; k < A; ++k) io.act_out[((size_t)t * io.n_episodes + eid) * A + k] = af[k];
            }
        } else if constexpr (POLICY == 1) {
            read_action<A>(io.actions, io.action_dtype, (size_t)t * io.n_episodes + eid, act);
            shape_action<PHASE, RTD>(D, act);
        } else {
            // classical P controller on v_ref(y) (landing_burn_pure_throttle.py:261-339)
            double alpha_eff = s.gamma - s.theta - PD_PI;
            stop = !(s.m_prop > 0.0 && s.y > 1.0 && (double)info.q < 65e3 && s.vy < 0.0 &&
                     alpha_eff < 5.0 * (180.0 / PD_PI));
            double speed = sqrt(s.vx * s.vx + s.vy * s.vy);
            double v_ref = D.sd.v_opt_a * (s.y * s.y) + D.sd.v_opt_b * s.y;
            double nn = -0.10 * (v_ref - speed) + 0.0;
            nn = nn < 0.0 ? 0.0 : (nn > 1.0 ? 1.0 : nn);
            act.u[0] = 2.0 * (nn - 0.5);
            act.f32 = false;
        }
        Rtd<R> out;
        out.reward = R(0); out.done = 0; out.truncated = 0; out.trunc_id = 0;
        int tid = -1;
        if (!stop) {
            if constexpr (POLICY == 2) {
                Control<R> ctl;
                if constexpr (WIND) D.gust_ahead<4, COOP>(w, wc, (unsigned)eid);
#pragma unroll 1
                for (int k = 0; k < 4; ++k)
                    D.substep<R, RT, PHASE, WIND, COOP>(s, act, prev, w, wc, (unsigned)eid, info, ctl, &sh);
            } else {
                R g1;
                D.env_step<R, RT, PHASE, RTD, WIND, COOP>(s, act, prev, w, wc, (unsigned)eid, gw, info, out, g1, &sh);
            }
            total -= (double)out.reward;
            if (io.traj && writer) {
                double *p = io.traj + ((size_t)t * io.n_episodes + eid) * 11;
                p[0] = s.x; p[1] = s.y; p[2] = s.vx; p[3] = s.vy; p[4] = s.theta; p[5] = s.theta_dot;
                p[6] = s.gamma; p[7] = s.alpha; p[8] = s.mass; p[9] = s.m_prop; p[10] = s.time;
            }
            if (io.rewards && writer) io.rewards[(size_t)t * io.n_episodes + eid] = (double)out.reward;
            ++t;
            if (out.done || out.truncated) { tid = out.trunc_id; stop = true; }
            else if (t >= io.max_steps) {
                stop = true;         // step cap: truncation id stays -1
                if constexpr (POLICY == 0) {
                    // per-particle MLP = PSO fitness: score the cut-off episode as a truncation
                    const R rc = D.cap_reward<R, PHASE>(s);
                    total -= (double)rc;
                    if (io.rewards && writer) io.rewards[(size_t)(t - 1) * io.n_episodes + eid] += (double)rc;
                }
            }
        }
        bool handoff = false;
        if constexpr (HANDOFF) {
            if (!stop && t >= io.handoff_steps) {
                // still running: append the complete episode state to the output continuation records
                int r = 0;
                if (writer) r = atomicAdd(io.out_count, 1);
                if (COOP > 1) {
                    const int leader = (threadIdx.x & 31) & ~(COOP - 1);
                    const unsigned gmask = (COOP == 32 ? 0xffffffffu : ((1u << COOP) - 1u)) << leader;
                    r = __shfl_sync(gmask, r, leader);
                }
                if (r < io.out_cap) {
                    if (writer) {
                        const size_t cap = (size_t)io.out_cap;
                        double *d = io.out_d + r;
                        int *ci = io.out_i + r;
                        ci[0] = eid; ci[cap] = t; ci[2 * cap] = gw.n; ci[3 * cap] = (int)w.ctr;
                        d[0] = s.x; d[cap] = s.y; d[2 * cap] = s.vx; d[3 * cap] = s.vy; d[4 * cap] = s.theta;
                        d[5 * cap] = s.theta_dot; d[6 * cap] = s.gamma; d[7 * cap] = s.alpha; d[8 * cap] = s.mass;
                        d[9 * cap] = s.m_prop; d[10 * cap] = s.time;
#pragma unroll
                        for (int k = 0; k < 10; ++k) d[(11 + k) * cap] = (double)gw.w[k];
                        d[21 * cap] = prev.gimbal_deg; d[22 * cap] = prev.dl; d[23 * cap] = prev.dr;
                        d[24 * cap] = w.xu0; d[25 * cap] = w.xu1; d[26 * cap] = w.xv0; d[27 * cap] = w.xv1;
                        d[28 * cap] = w.sigma_u; d[29 * cap] = w.sigma_v;
                        d[30 * cap] = total;
                    }
                    handoff = true;
                    stop = true;
                }
            }
        }
        if (stop) {
            if (!handoff && writer) {
                if (io.ret) io.ret[eid] = total;
                if (io.steps) io.steps[eid] = t;
                if (io.trunc_id) io.trunc_id[eid] = tid;
                if (io.terminal) {
                    double *p = io.terminal + (size_t)eid * 11;
                    p[0] = s.x; p[1] = s.y; p[2] = s.vx; p[3] = s.vy; p[4] = s.theta; p[5] = s.theta_dot;
                    p[6] = s.gamma; p[7] = s.alpha; p[8] = s.mass; p[9] = s.m_prop; p[10] = s.time;
                }
            }
            if (writer) e = atomicAdd(io.queue, 1);
            if (COOP > 1) {
                const int leader = (threadIdx.x & 31) & ~(COOP - 1);
                const unsigned gmask = (COOP == 32 ? 0xffffffffu : ((1u << COOP) - 1u)) << leader;
                e = __shfl_sync(gmask, e, leader);
            }
            active = e < n_work;
            if (active) begin_episode();
        }
    }
    if (info.rbf_status) atomicOr(status, info.rbf_status);
}

// ------------------------------------------------------------------ AoS <-> SoA
static __global__ void get_state_kernel(EnvSoA e, double *state, double *gwin, int *nwin, double *aprev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n) return;
    if (state)
        for (int k = 0; k < 11; ++k) state[(size_t)i * 11 + k] = e.st[(size_t)k * e.n + i];
    if (gwin)
        for (int k = 0; k < 10; ++k) gwin[(size_t)i * 10 + k] = e.gwin[(size_t)k * e.n + i];
    if (nwin) nwin[i] = e.gwin_n[i];
    if (aprev)
        for (int k = 0; k < 3; ++k) aprev[(size_t)i * 3 + k] = e.aprev[(size_t)k * e.n + i];
}
static __global__ void set_state_kernel(EnvSoA e, const double *state, const double *gwin, const int *nwin,
                                 const double *aprev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n) return;
    if (state)
        for (int k = 0; k < 11; ++k) e.st[(size_t)k * e.n + i] = state[(size_t)i * 11 + k];
    if (gwin)
        for (int k = 0; k < 10; ++k) e.gwin[(size_t)k * e.n + i] = gwin[(size_t)i * 10 + k];
    if (nwin) e.gwin_n[i] = nwin[i];
    if (aprev)
        for (int k = 0; k < 3; ++k) e.aprev[(size_t)k * e.n + i] = aprev[(size_t)i * 3 + k];
}

// ------------------------------------------------------------------ launch tables
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of every kernel
// instantiation: remember (kernel, device) pairs instead of one flag per process.
static void ensure_smem(const void *kernel, int device) {
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> done;
    std::lock_guard<std::mutex> lock(mu);
    if (done.insert({kernel, device}).second)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PD_SH_BYTES);
}
#define PD_SMEM_OPT_IN(kernel, lc) ensure_smem(reinterpret_cast<const void *>(&kernel), (lc).device)

template <typename R, typename RT>
struct Launch {
    template <int PHASE, int RTD, bool WIND>
    static void step_t(const LaunchCtx &lc, const EnvSoA &e, const StepIO &io, const WindCtx &wc, const double *sig,
                       int auto_reset, cudaStream_t st) {
        int threads, blocks;
        big_block_config(e.n, lc.n_sm, threads, blocks);
        if constexpr (sizeof(R) == 8 && !WIND) {
            // full-info diagnostic variant (fp64, no wind): the scalar drop-in env and the
            // trajectory export use it; never on the throughput path
            if (io.dbg_full) {
                PD_SMEM_OPT_IN((step_kernel<R, RT, PHASE, RTD, WIND, true>), lc);
                step_kernel<R, RT, PHASE, RTD, WIND, true><<<blocks, threads, PD_SH_BYTES, st>>>(*lc.kp, e, io, wc, sig, auto_reset);
                return;
            }
        }
        PD_SMEM_OPT_IN((step_kernel<R, RT, PHASE, RTD, WIND>), lc);
        step_kernel<R, RT, PHASE, RTD, WIND><<<blocks, threads, PD_SH_BYTES, st>>>(*lc.kp, e, io, wc, sig, auto_reset);
    }
    static void step(const LaunchCtx &lc, int phase, int rtd, int wind, const EnvSoA &e, const StepIO &io,
                     const WindCtx &wc, const double *sig, int auto_reset, cudaStream_t st) {
        int key = phase * 4 + rtd * 2 + (wind ? 1 : 0);
        switch (key) {
            case 0: step_t<0, 0, false>(lc, e, io, wc, sig, auto_reset, st); break;
            case 1: step_t<0, 0, true>(lc, e, io, wc, sig, auto_reset, st); break;
            case 2: step_t<0, 1, false>(lc, e, io, wc, sig, auto_reset, st); break;
            case 3: step_t<0, 1, true>(lc, e, io, wc, sig, auto_reset, st); break;
            case 4: step_t<1, 0, false>(lc, e, io, wc, sig, auto_reset, st); break;
            case 5: step_t<1, 0, true>(lc, e, io, wc, sig, auto_reset, st); break;
            case 6: step_t<1, 1, false>(lc, e, io, wc, sig, auto_reset, st); break;
            case 7: step_t<1, 1, true>(lc, e, io, wc, sig, auto_reset, st); break;
            // phases 2..6 exist with the rl closures only (pd_create rejects type 'pso')
            case 10: step_t<2, 1, false>(lc, e, io, wc, sig, auto_reset, st); break;
            case 11: step_t<2, 1, true>(lc, e, io, wc, sig, auto_reset, st); break;
            case 14: step_t<3, 1, false>(lc, e, io, wc, sig, auto_reset, st); break;
            case 15: step_t<3, 1, true>(lc, e, io, wc, sig, auto_reset, st); break;
            case 18: step_t<4, 1, false>(lc, e, io, wc, sig, auto_reset, st); break;
            case 19: step_t<4, 1, true>(lc, e, io, wc, sig, auto_reset, st); break;
            case 22: step_t<5, 1, false>(lc, e, io, wc, sig, auto_reset, st); break;
            case 23: step_t<5, 1, true>(lc, e, io, wc, sig, auto_reset, st); break;
            case 26: step_t<6, 1, false>(lc, e, io, wc, sig, auto_reset, st); break;
            default: step_t<6, 1, true>(lc, e, io, wc, sig, auto_reset, st); break;
        }
    }
    template <int PHASE, int RTD, bool WIND, int POLICY, int COOP, int MODE, int MAXB = PD_MAX_BLOCK>
    static void roll_launch(const LaunchCtx &lc, int blocks, int threads, const RolloutIO &io, const WindCtx &wc,
                            const double *sig, int *status, cudaStream_t st) {
        constexpr int O = phase_odim(PHASE);
        constexpr int NH = PHASE == 0 ? 3 : 4;
        // cooperative per-particle MLP: a thread-private column of weight rows behind the tables
        const unsigned smem = PD_SH_BYTES + ((POLICY == 0 && COOP >= 8) ? mlp_words<O, NH>() * threads * 4 + 16 : 0);
        static std::mutex mu;
        static std::set<int> done;
        {
            std::lock_guard<std::mutex> lock(mu);
            if (done.insert(lc.device).second)
                cudaFuncSetAttribute(rollout_kernel<R, RT, PHASE, RTD, WIND, POLICY, COOP, MODE, MAXB>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(PD_SH_BYTES + 52 * PD_MAX_BLOCK * 4 + 16));
        }
        init_queue_kernel<<<1, 1, 0, st>>>(io.queue, blocks * threads / COOP);
        rollout_kernel<R, RT, PHASE, RTD, WIND, POLICY, COOP, MODE, MAXB><<<blocks, threads, smem, st>>>(*lc.kp, io, wc, sig, status);
    }
    template <int PHASE, int RTD, bool WIND, int POLICY>
    static void roll_t(const LaunchCtx &lc, const RolloutIO &io_in, const WindCtx &wc, const double *sig, int *status,
                       cudaStream_t st) {
        // Persistent grids (one block per SM) fed by a work queue.  Fewer episodes than ~3/4 of the
        // GPU's lanes / 8: 8 lanes co-operate on each episode from the start (they split the 100 RBF
        // terms per sub-step and the 8 hidden units of the actor), which both fills the SMs and cuts
        // the per-step latency that bounds a generation by its longest episode.
        RolloutIO io = io_in;
        const int n_sm = lc.n_sm;
        const long long L = (long long)n_sm * PD_MAX_BLOCK;
        const bool coop = (long long)io.n_episodes * 8 <= L * 3 / 4;
        const int lanes_per = coop ? 8 : 1;
        int threads, blocks;
        big_block_config((long long)io.n_episodes * lanes_per, n_sm, threads, blocks);
        if (blocks > n_sm) blocks = n_sm;              // persistent: the queue feeds the rest
        if constexpr (POLICY == 0) {
            const bool staged = io.handoff_steps > 0 && io.cont_d && io.out_d;
            const int h1 = io.handoff_steps;
            if (staged && h1 < io.max_steps) {
                // Stage chain: reset -> h1 -> h2 -> 2 h2 -> 4 h2 ... -> end.  The survivors of a stage are
                // appended to continuation records (two buffers, the stages ping-pong) and the next
                // stage picks its cooperation from their number c, which only exists on the device: it
                // is launched in its 1-, 8- and 32-lane variants and the two whose window (run_if_gt,
                // run_if_le] does not hold c exit at once.  Per env step a stage costs about
                //   1 lane : 31 us (c = 2) ... 34 us (c = 8 288), 60 us with every lane busy
                //   8 lanes: 10 us (c <= 2 072), 14.4 us at c = 8 288 = one wave of 8-lane groups
                //   32 lanes: 7.4-8.0 us (c <= 592), 10.8 us at c = 2 072 = one wave of warps
                // (fp32 build, no wind, tools/lone_episode_latency.py), which cross at c = 2 waves of
                // 8-lane groups and at c = 1 wave of warps.
                const int T8 = io.lanes8_below > 0 ? io.lanes8_below : (int)(L / 4);
                const int T32 = io.lanes32_below > 0 ? io.lanes32_below : (int)(L / 32);
                const int T32s = T32 < 4 * n_sm ? T32 : 4 * n_sm;   // one wave of 4 warps per SM
                struct Buf { double *d; int *i; int *count; int cap; };
                const Buf buf[2] = {{io.cont_d, io.cont_i, io.cont_count, io.cont_cap},
                                    {io.out_d, io.out_i, io.out_count, io.out_cap}};
                auto reads = [&](RolloutIO &r, int k) {
                    r.cont_d = buf[k].d; r.cont_i = buf[k].i; r.cont_count = buf[k].count; r.cont_cap = buf[k].cap;
                };
                auto writes = [&](RolloutIO &r, int k) {
                    r.out_d = buf[k].d; r.out_i = buf[k].i; r.out_count = buf[k].count; r.out_cap = buf[k].cap;
                    cudaMemsetAsync(buf[k].count, 0, sizeof(int), st);
                };
                // PD_ROLLOUT_TRACE=1 (diagnostic): time every stage with events and print the survivor
                // counts to stderr; synchronises the stream, never set in production
                const bool trace = getenv("PD_ROLLOUT_TRACE") != nullptr;
                std::vector<cudaEvent_t> ev;
                std::vector<int> bounds, srcs;
                auto mark = [&](int boundary, int k) {
                    if (!trace) return;
                    cudaEvent_t x;
                    cudaEventCreate(&x);
                    cudaEventRecord(x, st);
                    ev.push_back(x); bounds.push_back(boundary); srcs.push_back(k);
                };
                mark(0, -1);
                RolloutIO first = io;
                writes(first, 0);
                if (coop) roll_launch<PHASE, RTD, WIND, POLICY, 8, 1>(lc, blocks, threads, first, wc, sig, status, st);
                else roll_launch<PHASE, RTD, WIND, POLICY, 1, 1>(lc, blocks, threads, first, wc, sig, status, st);
                int src = 0;
                long long h = io.handoff2_steps > h1 ? io.handoff2_steps : 2LL * h1;
                std::vector<int> counts;
                auto snapshot = [&](int k) {          // survivor count of the stage just enqueued
                    if (!trace) return;
                    int c = 0;
                    cudaStreamSynchronize(st);
                    cudaMemcpy(&c, buf[k].count, sizeof(int), cudaMemcpyDeviceToHost);
                    counts.push_back(c);
                };
                mark(h1, 0);
                snapshot(0);
                for (; h < io.max_steps; h *= 2) {
                    RolloutIO r = io;
                    reads(r, src);
                    writes(r, src ^ 1);
                    r.handoff_steps = (int)h;
                    r.run_if_gt = T8; r.run_if_le = 0x7fffffff;
                    roll_launch<PHASE, RTD, WIND, POLICY, 1, 3>(lc, n_sm, PD_MAX_BLOCK, r, wc, sig, status, st);
                    r.run_if_gt = T32; r.run_if_le = T8;
                    roll_launch<PHASE, RTD, WIND, POLICY, 8, 3>(lc, n_sm, PD_MAX_BLOCK, r, wc, sig, status, st);
                    r.run_if_gt = T32s; r.run_if_le = T32;
                    roll_launch<PHASE, RTD, WIND, POLICY, 32, 3>(lc, n_sm, PD_MAX_BLOCK, r, wc, sig, status, st);
                    r.run_if_gt = -1; r.run_if_le = T32s;
                    roll_launch<PHASE, RTD, WIND, POLICY, 32, 3, 128>(lc, n_sm, 128, r, wc, sig, status, st);
                    src ^= 1;
                    mark((int)h, src);
                    snapshot(src);
                }
                RolloutIO last = io;
                reads(last, src);
                last.run_if_gt = T8; last.run_if_le = 0x7fffffff;
                roll_launch<PHASE, RTD, WIND, POLICY, 1, 2>(lc, n_sm, PD_MAX_BLOCK, last, wc, sig, status, st);
                last.run_if_gt = T32; last.run_if_le = T8;
                roll_launch<PHASE, RTD, WIND, POLICY, 8, 2>(lc, n_sm, PD_MAX_BLOCK, last, wc, sig, status, st);
                last.run_if_gt = T32s; last.run_if_le = T32;
                roll_launch<PHASE, RTD, WIND, POLICY, 32, 2>(lc, n_sm, PD_MAX_BLOCK, last, wc, sig, status, st);
                last.run_if_gt = -1; last.run_if_le = T32s;
                roll_launch<PHASE, RTD, WIND, POLICY, 32, 2, 128>(lc, n_sm, 128, last, wc, sig, status, st);
                if (trace) {
                    mark(io.max_steps, -1);
                    cudaStreamSynchronize(st);
                    fprintf(stderr, "[pd rollout trace] %d episodes (T8 %d, T32 %d):", io.n_episodes, T8, T32);
                    for (size_t k = 1; k < ev.size(); ++k) {
                        float ms = 0.f;
                        cudaEventElapsedTime(&ms, ev[k - 1], ev[k]);
                        if (k - 1 < counts.size())
                            fprintf(stderr, "  ->%d %.2f ms, %d left;", bounds[k], ms, counts[k - 1]);
                        else
                            fprintf(stderr, "  ->end %.2f ms", ms);
                    }
                    fprintf(stderr, "\n");
                    for (cudaEvent_t x : ev) cudaEventDestroy(x);
                }
                return;
            }
        }
        if (coop) roll_launch<PHASE, RTD, WIND, POLICY, 8, 0>(lc, blocks, threads, io, wc, sig, status, st);
        else roll_launch<PHASE, RTD, WIND, POLICY, 1, 0>(lc, blocks, threads, io, wc, sig, status, st);
    }
    static int rollout(const LaunchCtx &lc, int policy, int phase, int rtd, int wind, const RolloutIO &io,
                       const WindCtx &wc, const double *sig, int *status, cudaStream_t st) {
        if (phase > 1) return 1;    // whole-episode rollouts: the two landing phases
        if (policy == 0) {          // per-particle MLP: pso rtd only
            int key = phase * 2 + (wind ? 1 : 0);
            switch (key) {
                case 0: roll_t<0, 0, false, 0>(lc, io, wc, sig, status, st); break;
                case 1: roll_t<0, 0, true, 0>(lc, io, wc, sig, status, st); break;
                case 2: roll_t<1, 0, false, 0>(lc, io, wc, sig, status, st); break;
                default: roll_t<1, 0, true, 0>(lc, io, wc, sig, status, st); break;
            }
            return 0;
        }
        if (policy == 1) {
            int key = phase * 4 + rtd * 2 + (wind ? 1 : 0);
            switch (key) {
                case 0: roll_t<0, 0, false, 1>(lc, io, wc, sig, status, st); break;
                case 1: roll_t<0, 0, true, 1>(lc, io, wc, sig, status, st); break;
                case 2: roll_t<0, 1, false, 1>(lc, io, wc, sig, status, st); break;
                case 3: roll_t<0, 1, true, 1>(lc, io, wc, sig, status, st); break;
                case 4: roll_t<1, 0, false, 1>(lc, io, wc, sig, status, st); break;
                case 5: roll_t<1, 0, true, 1>(lc, io, wc, sig, status, st); break;
                case 6: roll_t<1, 1, false, 1>(lc, io, wc, sig, status, st); break;
                default: roll_t<1, 1, true, 1>(lc, io, wc, sig, status, st); break;
            }
            return 0;
        }
        if (policy == 2) {
            if (phase != 0) return 1;
            if (wind) roll_t<0, 0, true, 2>(lc, io, wc, sig, status, st);
            else roll_t<0, 0, false, 2>(lc, io, wc, sig, status, st);
            return 0;
        }
        return 1;
    }
};

template <typename R>
static void impl_observe(const LaunchCtx &lc, int phase, int rtd, const EnvSoA &e, void *obs, cudaStream_t st) {
    int threads = 128, blocks = (e.n + threads - 1) / threads;
    const KParams &kp = *lc.kp;
    int key = phase * 2 + rtd;
    switch (key) {
        case 0: observe_kernel<R, 0, 0><<<blocks, threads, 0, st>>>(kp, e, (R *)obs); break;
        case 1: observe_kernel<R, 0, 1><<<blocks, threads, 0, st>>>(kp, e, (R *)obs); break;
        case 2: observe_kernel<R, 1, 0><<<blocks, threads, 0, st>>>(kp, e, (R *)obs); break;
        case 3: observe_kernel<R, 1, 1><<<blocks, threads, 0, st>>>(kp, e, (R *)obs); break;
        case 5: observe_kernel<R, 2, 1><<<blocks, threads, 0, st>>>(kp, e, (R *)obs); break;
        case 7: observe_kernel<R, 3, 1><<<blocks, threads, 0, st>>>(kp, e, (R *)obs); break;
        case 9: observe_kernel<R, 4, 1><<<blocks, threads, 0, st>>>(kp, e, (R *)obs); break;
        case 11: observe_kernel<R, 5, 1><<<blocks, threads, 0, st>>>(kp, e, (R *)obs); break;
        default: observe_kernel<R, 6, 1><<<blocks, threads, 0, st>>>(kp, e, (R *)obs); break;
    }
}

template <typename R, typename RT>
static void impl_reset(const LaunchCtx &lc, const EnvSoA &e, const uint8_t *mask, const WindCtx &wc, const double *sig,
                       cudaStream_t st) {
    int threads = 128, blocks = (e.n + threads - 1) / threads;
    reset_kernel<R><<<blocks, threads, 0, st>>>(*lc.kp, e, mask, wc, sig);
}

static void impl_get_state(const EnvSoA &e, double *state, double *gwin, int *nwin, double *aprev,
                           cudaStream_t st) {
    int threads = 128, blocks = (e.n + threads - 1) / threads;
    get_state_kernel<<<blocks, threads, 0, st>>>(e, state, gwin, nwin, aprev);
}
static void impl_set_state(const EnvSoA &e, const double *state, const double *gwin, const int *nwin,
                           const double *aprev, cudaStream_t st) {
    int threads = 128, blocks = (e.n + threads - 1) / threads;
    set_state_kernel<<<blocks, threads, 0, st>>>(e, state, gwin, nwin, aprev);
}
static void impl_transpose(const float *w, float *wT, int n_particles, int n_params, cudaStream_t st) {
    dim3 grid((n_particles + 31) / 32, (n_params + 31) / 32), block(32, 8);
    transpose_weights_kernel<<<grid, block, 0, st>>>(w, wT, n_particles, n_params);
}
}  // namespace pd
