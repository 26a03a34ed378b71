"""GPU parity tests: the CUDA path (through the C ABI, via the ctypes host layer) against
(a) fixtures produced by the unmodified reference (tests/golden) and (b) the CPU oracle
run on the same seeded inputs.

Tolerances (north_star): single-step next state and reward within 1e-12 relative in the
fp64 build (absolute floor per component for values that pass through zero) and 1e-5
relative in the fp32 build; done / truncated / truncation id exact.
"""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

P, G = "landing_burn_pure_throttle", "landing_burn"

# Absolute floors per state component [x, y, vx, vy, theta, theta_dot, gamma, alpha, m, mp, t]:
# "1e-12 relative" is measured against max(|reference|, floor).  The floors are the magnitudes
# the landing burn lives at, except for the pitch channel: theta_dot integrates the aerodynamic
# moment, whose C_L comes from a thin-plate-spline sum with a cancellation factor
# kappa = sum|c_i phi_i| / |f| of up to ~6e4 - one ulp of log() or a different summation order
# moves C_L by ~5e-12 relative in *any* implementation, including between two LAPACK/BLAS builds
# running the reference itself.  That bounds theta_dot to ~3e-13 rad/s per 0.025 s
# sub-step (P) and ~1.2e-12 per 0.1 s sub-step (G); theta/alpha inherit dt * that.
FLOOR = {
    "landing_burn_pure_throttle": np.array([1e3, 1e3, 1e2, 1e2, 1.0, 2.5, 1.0, 1.0, 1e5, 1e5, 1e2]),
    "landing_burn": np.array([1e3, 1e3, 1e2, 1e2, 2.0, 10.0, 1.0, 2.0, 1e5, 1e5, 1e2]),
}


def state_err(a, b, phase="landing_burn_pure_throttle"):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), FLOOR[phase]), axis=-1)


@pytest.fixture(scope="module")
def envs_mod():
    from psso_sac_for_powered_descent_b200 import envs
    return envs


def _load_fixture_batch(envs_mod, g, phase, precision, rtd="pso"):
    n = len(g["state"])
    env = envs_mod.BatchedRocketEnv(n, rtd, phase, precision=precision)
    env.set_state(g["state"], g["win"], g["nwin"].astype(np.int32), g["aprev"])
    return env


@pytest.mark.parametrize("tag,phase", [("P", P), ("G", G)])
@pytest.mark.parametrize("key,akey", [("o64", "act64"), ("o32", "act32")])
def test_single_step_vs_reference_fp64(envs_mod, golden, tag, phase, key, akey):
    g = golden(f"single_step_{tag}.npz")
    env = _load_fixture_batch(envs_mod, g, phase, "fp64")
    act = torch.as_tensor(g[akey]).cuda()
    dbg = torch.zeros(env.n_envs, 16, dtype=torch.float64, device="cuda")
    obs, rew, done, trunc, tid = env.step(act, dbg=dbg)
    env.check_status()
    st, gw, nw, ap = env.get_state(full=True)
    ref = g[key]
    err = state_err(st.cpu().numpy(), ref[:, :11], phase)
    assert err.max() < 1e-12, (err.argmax(), err.max())
    r = rew.cpu().numpy()
    assert np.max(np.abs(r - ref[:, 11]) / np.maximum(np.abs(ref[:, 11]), 1.0)) < 1e-12
    assert np.array_equal(done.cpu().numpy().astype(float), ref[:, 12])
    assert np.array_equal(trunc.cpu().numpy().astype(float), ref[:, 13])
    assert np.array_equal(tid.cpu().numpy().astype(float), ref[:, 14])
    if phase == G:
        assert np.max(np.abs(ap.cpu().numpy() - ref[:, 15:18])) < 1e-13
    # intermediates of the last sub-step (looser: C_D / C_L carry the 5e4 cancellation factor
    # of the thin-plate-spline sum)
    d = dbg.cpu().numpy()
    cols = list(g["out_cols"][18:])
    for name, tol in (("mach", 1e-12), ("q", 1e-12), ("rho", 1e-12), ("p_atm", 1e-12), ("a", 1e-13),
                      ("x_cog", 1e-13), ("inertia", 1e-13), ("mass_flow", 1e-12),
                      ("throttle", 1e-12), ("CL", 2e-9), ("CD", 2e-9), ("g1", 1e-10)):
        j = cols.index(name)
        refv = ref[:, 18 + j]
        e = np.max(np.abs(d[:, j if name != "g1" else 12] - refv) / np.maximum(np.abs(refv), 1e-3))
        assert e < tol, (name, e)


@pytest.mark.parametrize("tag,phase", [("P", P), ("G", G)])
def test_single_step_vs_reference_fp32(envs_mod, golden, tag, phase):
    g = golden(f"single_step_{tag}.npz")
    env = _load_fixture_batch(envs_mod, g, phase, "fp32")
    obs, rew, done, trunc, tid = env.step(torch.as_tensor(g["act32"]).cuda())
    env.check_status()
    st = env.get_state().cpu().numpy()
    ref = g["o32"]
    err = state_err(st, ref[:, :11], phase)
    assert err.max() < 1e-5, (err.argmax(), err.max())
    # flags: exact.  A row may only differ if one of the closures' thresholded quantities of the
    # REFERENCE's next state (y, q - 65 kPa, speed - 5.5, g - 6, vy, theta - theta_lim, m_prop ...)
    # lies within the fp32 build's own 1e-5 state tolerance of its threshold; any other mismatch fails.
    d, t, i = done.cpu().numpy(), trunc.cpu().numpy(), tid.cpu().numpy()
    same = (d == ref[:, 12]) & (t == ref[:, 13]) & (i == ref[:, 14])
    from oracle import pd_oracle as O
    from psso_sac_for_powered_descent_b200.parity import threshold_margins
    jg = list(g["out_cols"]).index("g1")
    for k in np.nonzero(~same)[0]:
        m = threshold_margins(phase, "pso", ref[k, :11], ref[k, jg], O.isa(ref[k, 1])[0])
        key = min(m, key=m.get)
        assert m[key] < 1e-5, (f"row {k}: flags {(d[k], t[k], i[k])} vs reference {tuple(ref[k, 12:15])} with no "
                               f"threshold nearby (nearest: {key}, margin {m[key]:.2e})")
    print(f"fp32 single step {tag}: {int((~same).sum())} of {len(same)} fixture rows differ in a flag")
    ok = same
    r = rew.cpu().numpy().astype(float)
    assert np.max(np.abs(r - ref[:, 11])[ok] / np.maximum(np.abs(ref[:, 11][ok]), 1.0)) < 1e-5


def test_tape_replay_golden_trajectory(envs_mod, golden):
    """The reference's committed 1281-step controller trajectory, replayed in one launch."""
    g = golden("p_tape_replay.npz")
    env = envs_mod.BatchedRocketEnv(1, "pso", P, precision="fp64")
    act = torch.as_tensor(g["u0"]).reshape(-1, 1, 1).cuda()
    out = env.rollout_tape(act, record=True)
    env.check_status()
    assert int(out["steps"][0]) == 1281 and int(out["trunc_id"][0]) == 0
    traj = out["traj"][:, 0].cpu().numpy()
    err = state_err(traj, g["states"])
    assert err[:4].max() < 1e-12
    # the pitch channel is unstable: a 1e-15 difference grows ~10x every 10 steps at first
    assert err[:50].max() < 1e-7
    # 5124 sub-steps: terminal metrics tolerance 1e-6
    assert err.max() < 1e-6, err.max()
    rew = out["rewards"][:, 0].cpu().numpy()
    assert abs(rew[-1] - 474318.95042647305) < 1e-3      # done reward = propellant left [kg]
    assert abs(-float(out["ret"][0]) - 474318.95042647305) < 1e-3


@pytest.mark.parametrize("tag,phase", [("P", P), ("G", G)])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_pso_fitness_vs_reference(envs_mod, golden, tag, phase, precision):
    g = golden(f"pso_fitness_{tag}.npz")
    model = envs_mod.pso_wrapped_env(flight_phase=phase, precision=precision)
    fit, steps, tid, term = model.evaluate(g["positions"], terminal=True)
    fit, steps, tid = fit.cpu().numpy(), steps.cpu().numpy(), tid.cpu().numpy()
    # Episodes whose length changes when every action is nudged by one float32 ulp *in the
    # reference itself* (tools/make_golden.py) cannot be pinned by any implementation whose
    # fp32 MLP rounds differently from torch-CPU; they are sanity-checked only.
    wc = g["well_conditioned"].astype(bool)
    assert wc.sum() >= 10
    assert np.isfinite(fit).all() and (tid >= 0).all()
    if precision == "fp64":
        assert np.array_equal(steps[wc], g["steps"][wc])
        assert np.array_equal(tid[wc], g["trunc_id"][wc])
        # per-particle fp32 MLP: hidden sums cancel, so torch-CPU's summation order shows up
        # at ~1e-5 relative in single actions; terminal metrics tolerance 1e-4
        assert np.max(np.abs(fit - g["fitness"])[wc] / np.abs(g["fitness"][wc])) < 1e-4
        err = state_err(term.cpu().numpy()[wc], g["terminal_state"][wc], phase)
        assert err.max() < 1e-3
    else:
        assert np.mean(steps[wc] == g["steps"][wc]) >= 0.8
        ok = wc & (steps == g["steps"])
        assert np.max(np.abs(fit - g["fitness"])[ok] / np.abs(g["fitness"][ok])) < 1e-3


@pytest.mark.parametrize("tag,phase", [("P", P), ("G", G)])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_pso_many_particles_vs_reference(envs_mod, golden, tag, phase, precision):
    """256 random particles per phase through the unmodified reference's objective_function
    (tools/make_golden.py pso_fitness_many): episode length, truncation id and fitness.
    Particles whose episode changes length when every action is nudged by one float32 ulp IN THE
    REFERENCE are ill-conditioned for any implementation whose fp32 MLP rounds differently from
    torch-CPU; the stated agreement numbers are asserted on the well-conditioned ones and printed
    for all."""
    g = golden(f"pso_many_{tag}.npz")
    model = envs_mod.pso_wrapped_env(flight_phase=phase, precision=precision)
    fit, steps, tid, term = model.evaluate(g["positions"], terminal=True)
    fit, steps, tid = fit.cpu().numpy(), steps.cpu().numpy(), tid.cpu().numpy()
    wc = g["well_conditioned"].astype(bool)
    assert wc.sum() >= 200
    same_len = steps == g["steps"]
    same_id = tid == g["trunc_id"]
    rel = np.abs(fit - g["fitness"]) / np.maximum(np.abs(g["fitness"]), 1.0)
    print(f"pso_many {tag} {precision}: same length {same_len.mean():.4f} (well-conditioned "
          f"{same_len[wc].mean():.4f}), same id {same_id[wc].mean():.4f}, max fitness rel err on "
          f"same-length well-conditioned {rel[wc & same_len].max():.2e}")
    assert np.isfinite(fit).all() and (tid >= 0).all()
    if precision == "fp64":
        # the fp64 build differs from the reference only by the fp32 MLP summation order
        assert same_len[wc].mean() >= 0.99 and same_id[wc].mean() >= 0.99
        assert rel[wc & same_len].max() < 1e-4
    else:
        # production build: stated agreement on >= 200 well-conditioned particles per phase
        # (landing_burn fitness evaluations run the fp64 instantiation in every handle: a phase
        # whose pitch channel amplifies ~50x per step turns fp32 rounding into episode lengths)
        assert same_len[wc].mean() >= 0.99, same_len[wc].mean()
        assert same_id[wc & same_len].all()
        assert rel[wc & same_len].max() < 1e-3


@pytest.mark.parametrize("tag,phase", [("P", P), ("G", G)])
@pytest.mark.parametrize("key,dtype", [("f64", torch.float64), ("f32", torch.float32)])
def test_batch_tape_vs_reference(envs_mod, golden, tag, phase, key, dtype):
    """SURVEY 8(d) config-2 parity subset: 1 024 envs x 64 steps of random actions from reset through
    the UNMODIFIED reference env (committed fixture, float64 tape and the same values as float32),
    against the fp64 build: every step's flags and reward, states at steps 1, 2, 4, ..., 64."""
    g = golden(f"batch_tape_{tag}.npz")
    acts = torch.as_tensor(g["actions"]).to(dtype).cuda()          # [E, T, A]
    E, T, A = acts.shape
    env = envs_mod.BatchedRocketEnv(E, "pso", phase, precision="fp64")
    out = env.rollout_tape(acts.permute(1, 0, 2).contiguous(), record=True)
    env.check_status()
    steps_ref, flags_ref = g[f"steps_{key}"], g[f"flags_{key}"]
    steps = out["steps"].cpu().numpy()
    tid = out["trunc_id"].cpu().numpy()
    traj = out["traj"].cpu().numpy()                                # [T, E, 11]
    rew = out["rewards"].cpu().numpy()                              # [T, E]
    ended_ref = steps_ref < T
    same_len = steps == steps_ref
    # G under random actions is chaotic: its pitch channel amplifies a perturbation ~50x per 0.4 s
    # env step, so a 1e-13 rounding difference (another libm) decides some episode lengths after
    # 8+ steps - the reference's own float32 and float64 tapes end 21 % of these episodes at
    # different steps.  The bulk must agree; flags are compared wherever the final state still does.
    need = 0.999 if phase == P else 0.90
    assert same_len.mean() >= need, same_len.mean()
    last_id = flags_ref[np.arange(E), steps_ref - 1, 2]
    term_ok = same_len & (state_err(out["terminal"].cpu().numpy(), g[f"last_{key}"], phase) <
                          (1e-6 if phase == P else 1e-3))
    assert term_ok.mean() >= (0.999 if phase == P else 0.60), term_ok.mean()
    assert np.array_equal(np.where(ended_ref, last_id, -1)[term_ok], tid[term_ok])
    # the first 4 steps are pinned for EVERY env, before the chaos has had time to act (G grows a
    # 1e-12 single-step difference ~20x per step: measured 3e-8 after 4 steps)
    for j, k in enumerate(g["keep_steps"][:3]):
        live = steps_ref > k
        e4 = state_err(traj[k][live], g[f"states_{key}"][:, j][live], phase)
        assert e4.max() < (1e-10 if phase == P else (1e-11, 1e-9, 1e-6)[j]), (k, e4.max())
        fl = flags_ref[:, k]
        assert np.array_equal(fl[live & (steps > k), 0] + fl[live & (steps > k), 1] > 0,
                              (steps == k + 1)[live & (steps > k)])
    worst = np.zeros(E)
    for j, k in enumerate(g["keep_steps"]):
        ref = g[f"states_{key}"][:, j]
        live = same_len & (steps_ref > k)
        if live.any():
            worst[live] = np.maximum(worst[live], state_err(traj[k][live], ref[live], phase))
    live = same_len
    ridx = np.minimum(steps_ref, T) - 1
    rr = g[f"rewards_{key}"][np.arange(E), ridx]
    rerr = np.abs(rew[ridx, np.arange(E)] - rr) / np.maximum(np.abs(rr), 1.0)
    w = np.sort(worst[same_len])
    print(f"batch tape {tag} {key}: same length {same_len.mean():.4f}; state err median {w[len(w) // 2]:.2e} "
          f"p99 {w[int(0.99 * len(w))]:.2e} max {w[-1]:.2e}; terminal reward err max {rerr[live].max():.2e}")
    # 64 steps = 256 sub-steps of growth on the 1e-12 single-step bar; the aero interpolant is
    # discontinuous (a neighbour-set boundary crossed one sub-step earlier separates two
    # trajectories by ~1e-5) and the pitch channel is unstable: the bulk stays tight, none blows up
    assert w[len(w) // 2] < (1e-9 if phase == P else 1e-4), w[len(w) // 2]
    assert w[int(0.99 * len(w))] < (1e-5 if phase == P else 1.0)


@pytest.mark.parametrize("phase,n_steps", [(P, 1000), (G, 200)])
def test_fp32_vs_fp64_config2_tape(phase, n_steps):
    """BASELINE config 2 at full size (65 536 envs, random float32 actions, auto-reset) through the
    fp32 production build and the oracle-pinned fp64 build in lock step: termination-flag and
    episode-length agreement, with the conditioning baseline (fp64 vs fp64 under a 1-ulp action
    nudge) measured beside it.  Every recorded P disagreement must sit on a threshold."""
    from psso_sac_for_powered_descent_b200 import parity
    r = parity.fp32_vs_fp64_tape(65536, n_steps, phase=phase)
    b = parity.fp32_vs_fp64_tape(65536, n_steps, phase=phase, test="ulp")
    print({k: r[k] for k in ("flag_match_frac", "episode_same_length_frac", "episodes", "flag_mismatches",
                             "max_translational_err")},
          "baseline", {k: b[k] for k in ("flag_match_frac", "episode_same_length_frac", "flag_mismatches")})
    if phase == P:
        assert r["flag_match_frac"] >= 0.99999 and r["episode_same_length_frac"] >= 0.9995
        for rec in r["first_mismatches"]:
            if "margin" in rec:
                assert rec["margin"] < 1e-5, rec
    else:
        # chaotic phase: a 1-ulp action change alone alters > 10 % of the episode lengths; the fp32
        # build must stay within a factor 3 of that intrinsic disagreement
        assert (1 - r["episode_same_length_frac"]) <= 3.0 * (1 - b["episode_same_length_frac"]) + 0.01
        assert r["flag_match_frac"] >= 0.95


def test_pso_best_actor_known_answer(envs_mod, golden):
    """The reference's own saved best actor (data/pso_saves/.../PSO_different_starting_point):
    823 steps, fitness -590 404.49 (= propellant left at touchdown)."""
    g = golden("pso_best_actor_P.npz")
    model = envs_mod.pso_wrapped_env(flight_phase=P, precision="fp64")
    f = model.objective_function(g["weights"])
    assert model.last_steps == 823 and model.truncation_id() == 0
    assert abs(f - float(g["stored_fitness"])) / abs(float(g["stored_fitness"])) < 1e-6
    assert abs(f - float(g["fitness"])) / abs(float(g["fitness"])) < 1e-6
    assert len(model.bounds) == 249 and model.bounds[0] == (-1.5, 1.5)


def test_classical_controller_config1(envs_mod, golden):
    g = golden("classical_rollout_P.npz")
    env = envs_mod.BatchedRocketEnv(1, "pso", P, precision="fp64")
    out = env.rollout_classical(1, max_steps=4000, record=True)
    env.check_status()
    n = int(out["steps"][0])
    assert n == int(g["steps"]) == 1281
    last = out["terminal"][0].cpu().numpy()
    assert state_err(last, g["states"][-1]) < 1e-6
    # landing metrics: touchdown vy, altitude, horizontal speed, fuel left
    assert abs(last[3] - (-1.3753416481551386)) < 1e-5
    assert abs(last[1] - 0.8717687151879727) < 1e-5
    assert abs(last[9] - 474318.95042647305) < 1e-2
    traj = out["traj"][:n, 0].cpu().numpy()
    assert state_err(traj[:4], g["states"][:4]).max() < 1e-12
    assert state_err(traj[:20], g["states"][:20]).max() < 1e-9


def test_wind_noise_tape(envs_mod, golden):
    g = golden("wind_sequence_P.npz")
    env = envs_mod.BatchedRocketEnv(1, "pso", P, enable_wind=True, stochastic_wind=True,
                                    horiontal_wind_percentile=int(g["percentile"]), precision="fp64")
    env.set_wind_tape(g["tape"][None, :], [[float(g["sigma_u"]), float(g["sigma_v"])]])
    env.reset()
    dbg = torch.zeros(1, 16, dtype=torch.float64, device="cuda")
    worst = 0.0
    for k, a in enumerate(g["actions"]):
        env.step(torch.tensor([[a]], dtype=torch.float32, device="cuda"), dbg=dbg)
        if k % 10 == 0 or k == len(g["actions"]) - 1:
            st = env.get_state()[0].cpu().numpy()
            worst = max(worst, float(state_err(st, g["states"][k])))
            ugvg = dbg[0, 13:15].cpu().numpy()
            assert np.max(np.abs(ugvg - g["ug_vg"][k])) < 1e-9 * max(1.0, abs(g["ug_vg"][k][0]))
            if k <= 20:
                assert state_err(st, g["states"][k]) < 1e-11
    env.check_status()
    # 1040 sub-steps of error growth through the unstable pitch channel
    assert worst < 1e-7, worst


@pytest.mark.parametrize("tag,phase", [("P", P), ("G", G)])
def test_rl_wrapper_sequence(envs_mod, golden, tag, phase):
    g = golden(f"rl_sequence_{tag}.npz")
    env = envs_mod.rl_wrapped_env_pytorch(flight_phase=phase, enable_wind=False, trajectory_length=1,
                                          discount_factor=0.99, precision="fp64")
    o = env.reset()
    assert env.state_dim == g["obs"].shape[1] and env.action_dim == g["actions"].shape[1]
    assert np.max(np.abs(o - g["obs"][0])) < 1e-12
    for k, a in enumerate(g["actions"]):
        o, r, d, t, info = env.step(a)
        assert np.max(np.abs(o - g["obs"][k + 1])) < 2e-6, k     # fp32-rounded observation
        assert abs(r - g["rewards"][k]) < 1e-6 * max(1.0, abs(g["rewards"][k])), k
        assert d == bool(g["done"][k]) and t == bool(g["truncated"][k]), k
    assert env.truncation_id() == int(g["trunc_id"])


@pytest.mark.parametrize("phase,adim", [(P, 1), (G, 4)])
def test_batch_vs_oracle_random(envs_mod, oracle_tables, phase, adim):
    """Seeded random actions, 48 envs x 12 steps (P) / 5 steps (G, 0.4 s each) from reset:
    CUDA fp64 vs the CPU oracle."""
    from oracle import pd_oracle as O
    rng = np.random.default_rng(11)
    B, T = 48, (12 if phase == P else 5)
    acts = rng.uniform(-1, 1, (T, B, adim))
    env = envs_mod.BatchedRocketEnv(B, "pso", phase, precision="fp64")
    env.reset()
    cu_states, cu_flags = [], []
    for t in range(T):
        obs, rew, done, trunc, tid = env.step(torch.as_tensor(acts[t]).cuda())
        cu_states.append(env.get_state().cpu().numpy())
        cu_flags.append((done.cpu().numpy().copy(), trunc.cpu().numpy().copy(), tid.cpu().numpy().copy(),
                         rew.cpu().numpy().copy(), obs.cpu().numpy().copy()))
    env.check_status()
    worst_per_env = []
    for b in range(0, B, 3):
        worst = 0.0
        o = O.OracleEnv(phase, "pso", tables=oracle_tables)
        o.reset()
        m = O.PsoModel.__new__(O.PsoModel)
        m.env, m.flight_phase = o, phase
        for t in range(T):
            s, r, d, tr, info = o.step(acts[t, b])
            worst = max(worst, float(state_err(cu_states[t][b], np.array(s, float), phase)))
            assert bool(cu_flags[t][0][b]) == d and bool(cu_flags[t][1][b]) == tr
            assert int(cu_flags[t][2][b]) == o.truncation_id
            assert abs(cu_flags[t][3][b] - r) <= 1e-9 * max(1.0, abs(r))
            # observation function in isolation: oracle obs of the CUDA state
            assert np.max(np.abs(cu_flags[t][4][b] - m.obs(cu_states[t][b]))) < 1e-12
            if d or tr:
                break
        worst_per_env.append(worst)
    # 12 steps = 48 sub-steps of error growth on top of the 1e-12 single-step bar.  The
    # reference's aero interpolant is discontinuous (it jumps by up to 1.3e-2 where the 50-NN
    # set changes), so a trajectory that crosses such a boundary one sub-step earlier or later
    # than the oracle separates by ~1e-5; the bulk of the envs must stay tight, none may blow up.
    w = np.sort(np.array(worst_per_env))
    assert w[int(0.75 * len(w))] < (5e-11 if phase == P else 1e-8), w
    assert w[-1] < 1e-3, w


def test_auto_reset_and_reset_mask(envs_mod):
    B = 256
    env = envs_mod.BatchedRocketEnv(B, "pso", P, precision="fp32", auto_reset=True)
    init = np.array(env.params.initial_state)
    a = torch.full((B, 1), -1.0, dtype=torch.float32, device="cuda")   # u0=-1: q>65 kPa at step 101
    n_done = 0
    for t in range(110):
        obs, rew, done, trunc, tid = env.step(a)
        if trunc.any():
            n_done += int(trunc.sum())
            st = env.get_state().cpu().numpy()
            idx = trunc.cpu().numpy().astype(bool)
            assert np.allclose(st[idx], init, rtol=0, atol=0)           # reset inside the step
            assert (tid[trunc.bool()] == 4).all()
            assert t == 100
    assert n_done == B
    env.reset()
    env.step(a)
    mask = torch.zeros(B, dtype=torch.uint8, device="cuda")
    mask[::2] = 1
    st = env.reset(mask).cpu().numpy()
    assert np.array_equal(st[::2], np.tile(init, (B // 2, 1)))
    assert not np.array_equal(st[1::2], np.tile(init, (B // 2, 1)))


def test_full_size_properties(envs_mod):
    """BASELINE config 2 size (65 536 envs): determinism, finite states, constant-action
    episode lengths known from the reference (u0 = -1/0/+1 -> 101/131/397 steps, ids 4/4/6)."""
    B = 65536
    env = envs_mod.BatchedRocketEnv(B, "pso", P, precision="fp32")
    u = torch.zeros(B, 1, dtype=torch.float32, device="cuda")
    u[0::3] = -1.0
    u[2::3] = 1.0
    acts = u.unsqueeze(0).expand(420, B, 1).contiguous()
    out = env.rollout_tape(acts)
    out2 = env.rollout_tape(acts)
    env.check_status()
    steps, tid = out["steps"].cpu().numpy(), out["trunc_id"].cpu().numpy()
    assert np.array_equal(steps, out2["steps"].cpu().numpy())
    assert torch.equal(out["terminal"], out2["terminal"])
    assert np.isfinite(out["terminal"].cpu().numpy()).all()
    assert (steps[0::3] == 101).all() and (tid[0::3] == 4).all()
    assert (steps[1::3] == 131).all() and (tid[1::3] == 4).all()
    assert (steps[2::3] == 397).all() and (tid[2::3] == 6).all()


# --------------------------------------------------------------------------- shared SAC actor
def _make_actor(state_dim, action_dim, hidden=256, seed=0):
    """Same module structure and default init as src/agents/sac_pytorch.Actor (:129-160)."""
    import torch.nn as nn
    torch.manual_seed(seed)

    class Actor(nn.Module):
        def __init__(self):
            super().__init__()
            layers, d = [], state_dim
            for _ in range(2):
                layers += [nn.Linear(d, hidden), nn.ReLU()]
                d = hidden
            self.layers = nn.Sequential(*layers)
            self.mean = nn.Linear(hidden, action_dim)
            self.log_std = nn.Linear(hidden, action_dim)
            self.max_action = 1.0

        def forward(self, s):
            x = self.layers(s)
            return self.mean(x), torch.clamp(self.log_std(x), -20, 2)
    return Actor()


@pytest.mark.parametrize("phase,O,A", [(P, 2, 1), (G, 5, 4)])
def test_shared_actor_tensor_core_vs_torch_fp32(envs_mod, phase, O, A):
    env = envs_mod.BatchedRocketEnv(256, "rl", phase, precision="fp32", auto_reset=True)
    actor = _make_actor(O, A)
    # make the GEMM matter: scale the hidden layer so activations are O(1)
    with torch.no_grad():
        actor.layers[2].weight.mul_(4.0)
    n = 1000                                       # ragged last tile (1000 = 7*128 + 104)
    g = torch.Generator().manual_seed(1)
    obs = (torch.rand(n, O, generator=g) * 2 - 1)
    with torch.no_grad():
        mean_ref, _ = actor(obs)
        act_ref = torch.tanh(mean_ref)
    act32, mean32 = env.actor_forward(actor, obs.cuda(), deterministic=True, fp32_path=True, want_mean=True)
    assert torch.allclose(mean32.cpu(), mean_ref, rtol=1e-5, atol=2e-6)
    assert torch.allclose(act32.cpu(), act_ref, rtol=1e-5, atol=2e-6)
    act_tc, mean_tc = env.actor_forward(actor, obs.cuda(), deterministic=True, want_mean=True)
    # IEEE-half operands (11-bit significand), fp32 accumulation over K = 256: ~3e-4 of the activation
    # scale (bfloat16 operands gave 2e-3 and a 1e-2 bound here)
    scale = float(mean_ref.abs().max())
    err_m, err_a = float((mean_tc.cpu() - mean_ref).abs().max()), float((act_tc.cpu() - act_ref).abs().max())
    print(f"tensor-core actor {phase}: max |mean - torch fp32| = {err_m:.2e} (scale {scale:.2f}), action {err_a:.2e}")
    assert err_m < 1.5e-3 * max(scale, 1.0)
    assert err_a < 1.5e-3
    # and it must not be a trivially-zero output
    assert float(mean_tc.abs().max()) > 0.05


def test_collect_shared_actor_matches_stepwise(envs_mod):
    """collect() == manual loop of actor_forward + step on a twin env (deterministic actor)."""
    B, T = 512, 6
    actor = _make_actor(2, 1)
    e1 = envs_mod.BatchedRocketEnv(B, "rl", P, precision="fp32", auto_reset=True)
    e2 = envs_mod.BatchedRocketEnv(B, "rl", P, precision="fp32", auto_reset=True)
    out = e1.collect(actor, T, deterministic=True)
    nv = np.array(e2.params.norm_vals)
    st = e2.get_state().cpu().numpy().astype(np.float32)
    obs = torch.as_tensor(np.stack([(1 - st[:, 1] / nv[0]) * 2 - 1, (1 - st[:, 3] / nv[1]) * 2 - 1], 1),
                          dtype=torch.float32).cuda()
    for t in range(T):
        assert torch.allclose(out["obs"][t], obs, atol=1e-6)
        a = e2.actor_forward(actor, obs, deterministic=True)
        assert torch.equal(out["actions"][t], a)
        o, r, d, tr, tid = e2.step(a)
        assert torch.equal(out["rewards"][t], r) and torch.equal(out["done"][t], d)
        assert torch.equal(out["next_obs"][t], o)
        obs = e2.next_obs.clone()
    e1.check_status()
    # stochastic mode: actions differ between envs and stay in (-1, 1)
    out2 = e1.collect(actor, 2, deterministic=False, seed=3)
    a = out2["actions"]
    assert float(a.abs().max()) <= 1.0 and float(a.std()) > 1e-3


@pytest.mark.parametrize("mode", ["zc_all", "zc_out", "copy"])
def test_step_host_matches_step(envs_mod, mode, monkeypatch):
    """Host-facing step (numpy in / numpy out; mapped pinned memory or explicit copies + CUDA
    graph) == device step."""
    monkeypatch.setenv("PD_HOST_STEP", mode)
    B = 1024
    e1 = envs_mod.BatchedRocketEnv(B, "pso", P, precision="fp32", auto_reset=True)
    e2 = envs_mod.BatchedRocketEnv(B, "pso", P, precision="fp32", auto_reset=True)
    rng = np.random.default_rng(2)
    for t in range(5):
        a = rng.uniform(-1, 1, (B, 1)).astype(np.float32)
        obs, rew, done, trunc, tid = e1.step_host(torch.as_tensor(a).pin_memory() if t % 2 else a)
        o2, r2, d2, t2, i2 = e2.step(torch.as_tensor(a).cuda())     # interleaved handles: re-activation
        assert np.array_equal(obs, o2.cpu().numpy()) and np.array_equal(rew, r2.cpu().numpy())
        assert np.array_equal(done, d2.cpu().numpy()) and np.array_equal(trunc, t2.cpu().numpy())
        assert torch.equal(e1.trunc_id, i2)
    assert torch.equal(e1.get_state(), e2.get_state())


# --------------------------------------------------------------------------- RL rtd, wind, API
@pytest.mark.parametrize("tag,phase,adim", [("P", P, 1), ("G", G, 4)])
def test_rl_single_step_vs_oracle(envs_mod, golden, oracle_tables, tag, phase, adim):
    """type='rl' closures (rtd_rl.py:194-336) on the fixture states: reward, flags, ids, and the
    fp32-rounded observation of rl_wrapped_env_pytorch; G actions go through augment_action."""
    from oracle import pd_oracle as O
    g = golden(f"single_step_{tag}.npz")
    n = 96
    env = envs_mod.BatchedRocketEnv(n, "rl", phase, precision="fp64", trajectory_length=1, discount_factor=0.99)
    env.set_state(g["state"][:n], g["win"][:n], g["nwin"][:n].astype(np.int32), g["aprev"][:n])
    act = g["act32"][:n]
    obs, rew, done, trunc, tid = env.step(torch.as_tensor(act).cuda())
    st = env.get_state().cpu().numpy()
    rl = O.RlEnv(phase, tables=oracle_tables)
    for i in range(n):
        rl.env.reset()
        rl.env.set_state(g["state"][i], g["win"][i][:g["nwin"][i]], *g["aprev"][i])
        o, r, d, t, info = rl.step(act[i])
        assert state_err(st[i], np.array(rl.env.state, float), phase) < 1e-12, i
        assert (bool(done[i]), bool(trunc[i]), int(tid[i])) == (d, t, rl.env.truncation_id), i
        assert abs(float(rew[i]) - r) <= 1e-9 * max(1.0, abs(r)), (i, float(rew[i]), r)
        assert np.max(np.abs(obs[i].cpu().numpy() - o)) < 1e-12, i


def test_pso_rollout_with_wind_tape_vs_oracle(envs_mod, oracle_tables):
    """Persistent rollout kernel with the gust filter on: same noise tape and sigmas as the
    oracle, per-particle MLP in the loop (fp64 build)."""
    from oracle import pd_oracle as O
    rng = np.random.default_rng(21)
    n = 6
    pos = rng.uniform(-1.5, 1.5, (n, 249))
    tape = rng.standard_normal((n, 8 * 700))
    sig = np.stack([rng.uniform(0.5, 2.25, n), rng.uniform(1.25, 2.0, n)], 1)
    model = envs_mod.pso_wrapped_env(flight_phase=P, enable_wind=True, stochastic_wind=True,
                                     horiontal_wind_percentile=50, precision="fp64", max_steps=700)
    model._b.n_envs = n            # the tape is indexed by episode
    model._b.set_wind_tape(tape, sig)
    fit, steps, tid = model.evaluate(pos)
    model._b.check_status()
    for i in range(n):
        ref = O.PsoModel(P, enable_wind=True, stochastic_wind=True, horiontal_wind_percentile=50,
                         tables=oracle_tables, wind_noise=dict(sigma_u=sig[i, 0], sigma_v=sig[i, 1], tape=tape[i]),
                         max_steps=700)
        f = ref.objective_function(pos[i])
        assert int(steps[i]) == ref.steps, (i, int(steps[i]), ref.steps)
        assert abs(float(fit[i]) - f) <= 1e-4 * abs(f), (i, float(fit[i]), f)


def test_wind_seeds_and_determinism(envs_mod):
    """Philox gusts: seeds differ, a given (seed, particle, wind-seed) is reproducible, sigma in the
    reference's ranges (vonkarman.py:62-63)."""
    rng = np.random.default_rng(4)
    pos = torch.as_tensor(rng.uniform(-1.5, 1.5, (64, 249)).astype(np.float32)).cuda()
    m1 = envs_mod.pso_wrapped_env(flight_phase=P, enable_wind=True, stochastic_wind=True, precision="fp32", seed=11)
    f1, s1, t1 = m1._b.rollout_pso(pos, n_seeds=8)
    f1b, _, _ = m1._b.rollout_pso(pos, n_seeds=8)
    assert torch.equal(f1, f1b)
    m2 = envs_mod.pso_wrapped_env(flight_phase=P, enable_wind=True, stochastic_wind=True, precision="fp32", seed=12)
    f2, _, _ = m2._b.rollout_pso(pos, n_seeds=8)
    assert not torch.equal(f1, f2)
    per_seed = f1.reshape(64, 8)
    assert float((per_seed.max(1).values - per_seed.min(1).values).max()) > 0.0     # gusts matter


def test_scalar_drop_in_api(envs_mod):
    """rocket_environment_pre_wrap mirror: same call shapes as the reference."""
    env = envs_mod.rocket_environment_pre_wrap(type="pso", flight_phase=P, enable_wind=False)
    s0 = env.reset()
    assert len(s0) == 11 and s0[1] == 30028.385497767023
    outs = []
    for a in ((0.25,), [0.25], np.array([0.25]), np.array([[0.25]])):
        env.reset()
        s, r, d, t, info = env.step(a)
        outs.append(s)
        assert isinstance(r, float) and isinstance(d, bool) and isinstance(t, bool)
        assert {"mach_number", "dynamic_pressure", "CL", "CD", "g_load_1_sec_window", "action_info"} <= set(info)
    assert outs[0] == outs[1] == outs[2] == outs[3]
    env.reset()
    s32, *_ = env.step(np.array([0.25], dtype=np.float32))      # float32 action: NEP-50 path
    assert s32 != outs[0] and abs(s32[8] - outs[0][8]) < 1.0
    with pytest.raises(TypeError):          # upstream: pso closures of this phase have the wrong arity
        envs_mod.rocket_environment_pre_wrap(type="pso", flight_phase="subsonic", enable_wind=False)
    with pytest.raises(NotImplementedError):    # does not run upstream either
        envs_mod.rocket_environment_pre_wrap(type="rl", flight_phase="landing_burn_ACS", enable_wind=False)
    with pytest.raises(TypeError):              # upstream: rl closures of the flip-over take one argument
        envs_mod.rocket_environment_pre_wrap(type="rl", flight_phase="flip_over_boostbackburn", enable_wind=False)
    with pytest.raises(AssertionError):
        envs_mod.rocket_environment_pre_wrap(type="pso", flight_phase="nonsense", enable_wind=False)
    m = envs_mod.pso_wrapped_env(flight_phase=G)
    assert len(m.bounds) == 372 and "0_weight_0" in m.mock_dictionary_of_opt_params


def test_full_size_G_and_cooperative_consistency(envs_mod):
    """The 8-lane cooperative rollout (small swarms) and the 1-lane work-queue rollout (large
    swarms) must agree on the same particles (fp64, well within chaotic growth for short episodes)."""
    rng = np.random.default_rng(8)
    pos = rng.uniform(-1.5, 1.5, (20000, 372)).astype(np.float32)
    m = envs_mod.pso_wrapped_env(flight_phase=G, precision="fp64")
    big = m._b.rollout_pso(torch.as_tensor(pos).cuda())               # 20000 episodes: 1 lane each
    small = m._b.rollout_pso(torch.as_tensor(pos[:512]).cuda())       # 512 episodes: 8 lanes each
    m._b.check_status()
    same = (big[1][:512] == small[1]).float().mean()
    assert float(same) > 0.9                                          # ill-conditioned G episodes aside
    ok = (big[1][:512] == small[1])
    rel = ((big[0][:512] - small[0]).abs() / small[0].abs())[ok]
    assert float(rel.median()) < 1e-9 and float(rel.max()) < 1e-2
    assert np.isfinite(big[0].cpu().numpy()).all() and int((big[2] < 0).sum()) == 0


def test_device_swarm_update_matches_numpy(envs_mod):
    """pd_pso_update against a numpy restatement of the reference update rule
    (particle_swarm_optimisation.py:431-436, 517-521, 112-118), and a short optimisation run."""
    from psso_sac_for_powered_descent_b200 import pso
    model = envs_mod.pso_wrapped_env(flight_phase=G, precision="fp32", max_steps=256)
    params = dict(pso.landing_burn_pso_params, pop_size=512, generations=20)
    sw = pso.DeviceSwarm(model, 512, params, seed=3, max_steps=256)
    x0, v0 = sw.x.clone(), sw.v.clone()
    fit = sw.step()
    assert fit.shape == (512,) and torch.isfinite(fit).all()
    # personal bests after generation 0 = the evaluated positions
    assert torch.equal(sw.best, x0) and torch.equal(sw.best_fit, fit)
    for k in range(2):
        idx = torch.nonzero(sw.swarm_of_all == k).flatten()
        j = idx[torch.argmin(fit[idx])]
        assert torch.equal(sw.swarm_best[k], x0[j]) and float(sw.swarm_best_fit[k]) == float(fit[j])
    # v1 = w v0 + c1 r1 (pb - x) + c2 r2 (lb - x) with pb == x  ->  every row of v1 is a scalar
    # multiple of (lb - x0); x1 = clip(x0 + v1)
    lb = sw.swarm_best[sw.swarm_of_all.long()]
    d = lb - x0
    ratio = sw.v / torch.where(d.abs() > 1e-9, d, torch.ones_like(d))
    big = d.abs() > 1e-3
    # (the sub-swarm best itself has lb - x0 == 0: its velocity must stay zero)
    r2 = torch.stack([ratio[i][big[i]].median() if bool(big[i].any()) else ratio.new_zeros(())
                      for i in range(512)])
    assert int((~big.any(dim=1)).sum()) == 2
    assert float(r2.min()) >= 0.0 and float(r2.max()) <= 1.0 and float(r2.std()) > 0.1
    assert torch.allclose(sw.v, r2[:, None] * d, atol=1e-9)
    assert torch.allclose(sw.x, torch.clamp(x0 + sw.v, -1.5, 1.5), atol=1e-12)
    assert torch.equal(sw.weights, sw.x.to(torch.float32))
    best0 = sw.global_best_fitness
    for _ in range(6):
        sw.step()
    assert sw.global_best_fitness <= best0
    assert (sw.best_fit <= fit + 1e-12).all()


def test_device_swarm_follows_host_dropin_exactly(envs_mod, tmp_path):
    """A seeded run of the device-resident swarm and of the host drop-in (rng='philox') started from
    the same positions: identical best-fitness history, generation by generation, through sharing
    (every 4), migration (every 3) and the re-initialisation (generation 10) - and the files
    the reference's loaders read."""
    import csv
    import pickle
    from psso_sac_for_powered_descent_b200 import pso
    params = dict(pso.landing_burn_pso_params, pop_size=96, generations=30, communication_freq=4,
                  migration_freq=3, re_initialise_generation=10, re_initialise_number_of_particles=60)
    model = envs_mod.pso_wrapped_env(flight_phase=G, precision="fp32", max_steps=256)
    host = pso.ParticleSubswarmOptimisation(G, save_interval=0, model=model, pso_params=params, seed=9,
                                            rng="philox", base_save_dir=str(tmp_path / "host"))
    dev = pso.DeviceSwarm(model, 96, params, seed=9, max_steps=256, positions=host.position.copy(),
                          base_save_dir=str(tmp_path / "dev"))
    for g in range(24):
        host.step_generation(g)
        dev.step()
    assert dev.global_best_fitness_array == host.global_best_fitness_array
    assert dev.N_total == len(host.position) == 60 and dev.members == host.members
    assert np.array_equal(dev.x.cpu().numpy(), host.position)
    assert np.array_equal(dev.v.cpu().numpy(), host.velocity)
    assert np.array_equal(dev.best_fit.cpu().numpy(), host.best_fitness)
    assert np.array_equal(dev.swarm_best_fit.cpu().numpy(), np.array(host.subswarm_best_fitnesses))
    assert np.array_equal(dev.gbest_pos.cpu().numpy(), host.global_best_position)
    d = dev.save_metrics()
    dev.save_results(); dev.save()
    h = list(csv.reader(open(f"{d}/fitness_history.csv")))
    assert h[0][:3] == ["Generation", "Global_Best_Fitness", "Average_Fitness"] and len(h) == 25
    assert [float(r[1]) for r in h[1:]] == host.global_best_fitness_array
    host_avg = host.average_particle_fitness_array
    assert np.allclose([float(r[2]) for r in h[1:]], host_avg, rtol=1e-12)
    sub = list(csv.reader(open(f"{d}/subswarm_0_metrics.csv")))
    assert len(sub) == 25 and sub[0][1] == "best_fitness"
    with open(tmp_path / "dev" / "saves" / "swarm.pkl", "rb") as f:
        sw = pickle.load(f)
    assert [len(x) for x in sw] == [len(m) for m in host.members]
    assert np.array_equal(sw[1][0]["position"], host.swarms[1][0]["position"])
    # resume on the device from that file (load_swarms renumbers the particles in list order, as a reload
    # of the reference's lists does): same particles, same bests, same fitness of the next generation
    dev2 = pso.DeviceSwarm(model, 96, params, seed=9, max_steps=256)
    dev2.load_swarms(str(tmp_path / "dev" / "saves" / "swarm.pkl"))
    dev2.generation = dev.generation
    order = torch.as_tensor([i for m in dev.members for i in m], device="cuda")
    assert [len(m) for m in dev2.members] == [len(m) for m in dev.members]
    assert torch.equal(dev2.x, dev.x[order]) and torch.equal(dev2.best_fit, dev.best_fit[order])
    assert torch.equal(dev2.swarm_best_fit, dev.swarm_best_fit) and dev2.global_best_fitness == dev.global_best_fitness
    fa, fb = dev.step().clone(), dev2.step().clone()
    assert torch.equal(fa[:60][order], fb[:60])
    host2 = pso.ParticleSubswarmOptimisation(G, save_interval=0, model=model, pso_params=params, seed=1,
                                             base_save_dir=str(tmp_path / "h2"), write_metrics=False)
    host2.load_swarms(str(tmp_path / "dev" / "saves" / "swarm.pkl"))       # the reference's resume path
    assert host2.global_best_fitness == min(p["best_fitness"] for x in sw for p in x)
    hdr = open(tmp_path / "dev" / "particle_subswarm_optimisation_results.csv").readline().split(",")
    assert hdr[0] == "Algorithm" and hdr[1] == "0_weight_0" and len(hdr) == 374


def test_optimisers_follow_the_reference_optimiser(envs_mod, golden, tmp_path):
    """The recorded run of the UNMODIFIED reference optimiser (tests/golden/pso_run_reference.npz: 7
    generations with sharing, migration, re-initialisation) followed (1) by the host drop-in with the
    CUDA fitness evaluation - same positions in every generation, bit for bit - and (2) by checking
    that the device-resident swarm started from the same positions evaluates the same first
    generation."""
    import json
    import sys
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).parent))
    from test_host_cpu import _follow_reference_run
    from psso_sac_for_powered_descent_b200 import pso
    g = golden("pso_run_reference.npz")
    phase = str(g["phase"])
    params = dict(pso.PSO_PARAMS[phase], **json.loads(str(g["knobs"])))
    model = envs_mod.pso_wrapped_env(flight_phase=phase, precision="fp64", max_steps=8192)
    opt = pso.ParticleSubswarmOptimisation(phase, save_interval=0, model=model, pso_params=params,
                                           seed=int(g["seed"]), rng="reference", base_save_dir=str(tmp_path))
    x0 = opt.position.copy()
    _follow_reference_run(g, opt, fit_tol=1e-4)
    dev = pso.DeviceSwarm(model, len(x0), params, seed=int(g["seed"]), max_steps=8192, positions=x0)
    fit = dev.step()[:len(x0)].cpu().numpy()
    ref_min = [g["g0_metrics"][k][2] for k in range(2)]
    assert np.allclose([fit[:4].min(), fit[4:].min()], ref_min, rtol=1e-4)
    assert abs(dev.global_best_fitness - g["g0_global"][0]) <= 1e-4 * abs(g["g0_global"][0])


@pytest.mark.parametrize("n", [1, 33, 449, 1000])
def test_ragged_batch_sizes_bitwise(envs_mod, golden, n):
    """Batch sizes that do not fill a warp / a block / a wave: every env's result is bit-identical
    to the same env stepped inside the 192-env fixture batch (no cross-lane dependence)."""
    g = golden("single_step_P.npz")
    ref_env = _load_fixture_batch(envs_mod, g, P, "fp64")
    act = torch.as_tensor(g["act32"]).cuda()
    ref_env.step(act)
    ref_state = ref_env.get_state().cpu().numpy()
    ref_rew = ref_env.reward.cpu().numpy().copy()
    idx = np.arange(n) % len(g["state"])
    env = envs_mod.BatchedRocketEnv(n, "pso", P, precision="fp64")
    env.set_state(g["state"][idx], g["win"][idx], g["nwin"][idx].astype(np.int32), g["aprev"][idx])
    obs, rew, done, trunc, tid = env.step(act[torch.as_tensor(idx).cuda()])
    env.check_status()
    assert np.array_equal(env.get_state().cpu().numpy(), ref_state[idx])
    assert np.array_equal(rew.cpu().numpy(), ref_rew[idx])


def test_config4_batch_on_one_gpu(envs_mod):
    """BASELINE config 4's full batch (1 048 576 envs, RL closures, stochastic wind) fits and steps
    on a single B200; auto-reset keeps every lane inside the aero tables."""
    n = 1 << 20
    env = envs_mod.BatchedRocketEnv(n, "rl", P, enable_wind=True, stochastic_wind=True,
                                    horiontal_wind_percentile=50, precision="fp32", auto_reset=True, seed=1)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(0)
    for k in range(12):
        obs, rew, done, trunc, tid = env.step(torch.rand(n, 1, device="cuda", generator=gen) * 2 - 1)
    env.check_status()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    st = env.get_state()
    assert float(st[:, 1].max()) <= 30028.385497767023 + 1e-6 and float(st[:, 9].min()) > 0


def test_rollout_straggler_handoff(envs_mod):
    """pd_set_rollout_handoff: episodes still running after N steps are finished by a second,
    8-lane cooperative pass.  Episodes that end before the hand-off are bit-identical to the
    one-pass rollout; handed-off ones resume from their exact state (same step count and fitness
    unless the different summation order of the cooperative RBF sums flips a chaotic episode)."""
    from psso_sac_for_powered_descent_b200 import _native as N
    n = 16384
    rng = np.random.default_rng(7)
    pos = torch.as_tensor(rng.uniform(-1.5, 1.5, (n, 249)).astype(np.float32)).cuda()
    res = {}
    for steps in (0, 300):
        env = envs_mod.BatchedRocketEnv(1, "pso", P, precision="fp64")
        N.check(env.lib.pd_set_rollout_handoff(env._h, steps))
        fit, st, tid, term = env.rollout_pso(pos, max_steps=1500, terminal=True)
        env.check_status()
        res[steps] = (fit.cpu().numpy(), st.cpu().numpy(), tid.cpu().numpy(), term.cpu().numpy())
    f0, s0, t0, x0 = res[0]
    f1, s1, t1, x1 = res[300]
    early = s0 < 300
    assert early.sum() > 0.5 * n and (~early).sum() > 100
    assert np.array_equal(f0[early], f1[early]) and np.array_equal(s0[early], s1[early])
    assert np.array_equal(t0[early], t1[early]) and np.array_equal(x0[early], x1[early])
    late = ~early
    assert np.isfinite(f1[late]).all() and (s1[late] >= 300).all()
    same = s0[late] == s1[late]
    assert same.mean() > 0.9
    assert np.max(np.abs(f0[late][same] - f1[late][same]) / np.maximum(np.abs(f0[late][same]), 1.0)) < 1e-6


@pytest.mark.parametrize("precision,wind", [("fp64", True), ("fp32", True), ("fp32", False)])
def test_rollout_stage_chain_every_lane_choice(envs_mod, precision, wind):
    """Doubling stage chain (pd_set_rollout_stages) with the lane choice of the record-fed stages
    forced to 1, 8 and 32 lanes per episode (pd_set_rollout_lanes), windy and not: every variant
    resumes every episode from its exact state.  Same Philox counters whichever lane draws the gust
    noise, so with wind the episodes are the same episodes; what differs is the summation order of
    the cooperative RBF sums, the FMA contraction of the other kernel instantiation and, in the
    fp32 build, atan2 against its incremental form, which a few chaotic episodes turn into a
    different length."""
    from psso_sac_for_powered_descent_b200 import _native as N
    n = 8192                          # more than 3/4 of the lanes / 8: the first stage is one lane per episode
    rng = np.random.default_rng(11)
    pos = torch.as_tensor(rng.uniform(-1.5, 1.5, (n, 249)).astype(np.float32)).cuda()
    env = envs_mod.BatchedRocketEnv(1, "pso", P, precision=precision, enable_wind=wind, stochastic_wind=wind, seed=5)
    out = {}
    for name, stages, lanes in (("one pass", (0, 0), (0, 0)), ("default", (128, 256), (0, 0)),
                                ("1 lane", (64, 128), (1, 1)), ("8 lanes", (64, 128), (1 << 20, 1)),
                                ("32 lanes", (64, 128), (1 << 20, 1 << 20))):
        N.check(env.lib.pd_set_rollout_stages(env._h, *stages))
        N.check(env.lib.pd_set_rollout_lanes(env._h, *lanes))
        fit, st, tid = env.rollout_pso(pos, max_steps=1200)
        env.check_status()
        out[name] = (fit.cpu().numpy(), st.cpu().numpy(), tid.cpu().numpy())
    f0, s0, t0 = out["one pass"]
    assert (s0 > 256).sum() > 40 and np.isfinite(f0).all()
    for name in ("default", "1 lane", "8 lanes", "32 lanes"):
        f, s, t = out[name]
        first = 128 if name == "default" else 64
        early = s0 < first            # finished inside the first stage: the same instructions
        assert np.array_equal(f[early], f0[early]) and np.array_equal(s[early], s0[early])
        same = s == s0
        assert same.mean() > 0.97, (name, same.mean())
        assert np.array_equal(t[same], t0[same])
        rel = np.abs(f[same] - f0[same]) / np.maximum(np.abs(f0[same]), 1.0)
        # rounding-level differences grow along an episode (parity.py conditioning baseline): bound the
        # bulk tightly and the worst episode loosely
        assert np.quantile(rel, 0.99) < (1e-9 if precision == "fp64" else 1e-4), (name, np.quantile(rel, 0.99))
        assert rel.max() < (1e-5 if precision == "fp64" else 5e-2), (name, rel.max())
    # record-fed 1-lane stages: the same arithmetic in another instantiation of the kernel, whose
    # FMA contraction differs in places (first difference 4e-16 in theta_dot, one step after the
    # resume) - rounding-level, then amplified like any perturbation
    f, s, t = out["1 lane"]
    assert (s == s0).mean() > 0.999
    with pytest.raises(RuntimeError):
        N.check(env.lib.pd_set_rollout_lanes(env._h, 8, 16))


@pytest.mark.parametrize("phase,rtd", [(P, "pso"), (G, "rl")])
def test_aero_patches_against_exact_sums(envs_mod, phase, rtd):
    """fp32 build: C_L / C_D come from bicubic patches of the 50-term thin-plate sums
    (csrc/pd_patch.h) unless exact_aero is set.  The builder keeps a patch only if it reproduces the
    exact sum to 1e-8 at 25 check points; here 65 536 spread-out states take one step through both
    variants: the new states agree far below the fp32 build's own rounding and every flag is equal
    except where a thresholded quantity sits within fp32 resolution of its threshold."""
    B = 65536
    adim = 1 if phase == P else 4
    patched = envs_mod.BatchedRocketEnv(B, rtd, phase, precision="fp32", auto_reset=True)
    exact = envs_mod.BatchedRocketEnv(B, rtd, phase, precision="fp32", auto_reset=True, exact_aero=True)
    st = patched.aero_patch_stats()
    assert st["cd_patches"] > 100000 and st["cl_patches"] > 1000000
    assert st["cd_rejected"] < 0.005 * st["cd_patches"] and st["cl_rejected"] < 0.005 * st["cl_patches"]
    assert 0.0 < st["max_abs_error_in_use"] <= 1e-8
    assert exact.aero_patch_stats()["cl_patches"] == 0
    gen = torch.Generator(device="cuda").manual_seed(4)
    for _ in range(24 if phase == G else 150):          # spread the batch over the flight envelope
        exact.step(torch.rand(B, adim, device="cuda", generator=gen) * 2 - 1)
    worst, bulk, flag_diff = 0.0, 0.0, 0
    for _ in range(8):
        state = exact.get_state(full=True)
        patched.set_state(*state)
        act = torch.rand(B, adim, device="cuda", generator=gen) * 2 - 1
        oe = exact.step(act)
        op = patched.step(act)
        same = (oe[2] == op[2]) & (oe[3] == op[3])
        flag_diff += int((~same).sum())
        live = same & ~(oe[2].bool() | oe[3].bool())     # ended episodes were reset in place
        a, b = patched.get_state()[live], exact.get_state()[live]
        # theta_dot is the integral of a small difference of large moments: absolute scale 1 rad/s
        scale = b.abs().clamp_min(torch.tensor([1e3, 1e3, 10., 10., 1., 1., 1., 1., 1e3, 1e3, 1.], device="cuda",
                                               dtype=torch.float64))
        per_env = ((a - b).abs() / scale).max(dim=1).values
        worst = max(worst, float(per_env.max()))
        bulk = max(bulk, float(torch.quantile(per_env, 0.999)))
        assert float(torch.quantile((op[1][live].double() - oe[1][live].double()).abs(), 0.999)) < 1e-6
    patched.check_status(); exact.check_status()
    # landing_burn's pitch channel amplifies a one-ulp change of a float force by orders of magnitude
    # inside one 0.4 s step on the few envs that tumble (theta_dot 0.35 -> -1.1 rad/s in the worst one)
    assert bulk < (1e-7 if phase == P else 1e-6), bulk
    assert worst < (1e-7 if phase == P else 1e-3), worst
    assert flag_diff <= 4, flag_diff


def test_aero_patch_cache_is_shared_and_releasable(envs_mod):
    """One patch set per GPU and table serves every fp32 handle of the process; it stays cached after the
    last handle is gone (the next pd_create does not rebuild it) until pd_release_aero_patches."""
    import gc
    from psso_sac_for_powered_descent_b200 import _native as N
    lib = N.load_library()
    a = envs_mod.BatchedRocketEnv(64, "pso", P, precision="fp32")
    b = envs_mod.BatchedRocketEnv(64, "rl", G, precision="fp32")
    st = a.aero_patch_stats()
    assert st == b.aero_patch_stats()
    assert lib.pd_release_aero_patches() == 0          # sets in use are never freed
    a.step(torch.zeros(64, 1, device="cuda")); a.check_status()
    a.close(); b.close()
    gc.collect()
    freed = lib.pd_release_aero_patches()
    assert freed in (0, 64 * (st["cd_patches"] + st["cl_patches"]))     # 0: another live handle still holds them
    c = envs_mod.BatchedRocketEnv(64, "pso", P, precision="fp32")       # rebuilt (or still cached)
    assert c.aero_patch_stats() == st
    c.step(torch.zeros(64, 1, device="cuda")); c.check_status()


# --------------------------------------------------------------------------- round-2 robustness
def test_captured_graph_survives_other_handles(envs_mod):
    """A handle's constants travel with every launch (a __grid_constant__ kernel parameter): a CUDA
    graph captured on one handle replays correctly after other handles - another phase, another
    precision - have been created and stepped in between (round 1 needed pd_activate for that)."""
    B = 2048
    rng = np.random.default_rng(3)
    acts = torch.as_tensor(rng.uniform(-1, 1, (6, B, 1)).astype(np.float32)).cuda()
    ref = envs_mod.BatchedRocketEnv(B, "pso", P, precision="fp32", auto_reset=True)
    for k in range(6):
        ref.step(acts[k])
    want = ref.get_state().clone()
    env = envs_mod.BatchedRocketEnv(B, "pso", P, precision="fp32", auto_reset=True)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        env.step(acts[0])                    # warm-up launch (module load, shared-memory opt-in)
        env.reset()
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for k in range(6):
                env.step(acts[k])
    other = envs_mod.BatchedRocketEnv(512, "rl", G, precision="fp32", auto_reset=True)
    other64 = envs_mod.BatchedRocketEnv(512, "pso", P, precision="fp64", enable_wind=True, stochastic_wind=True)
    other.step(torch.zeros(512, 4, device="cuda"))
    other64.step(torch.zeros(512, 1, dtype=torch.float64, device="cuda"))
    with torch.cuda.stream(stream):
        graph.replay()
        other.step(torch.zeros(512, 4, device="cuda"))       # interleaved with the replay's stream
    torch.cuda.synchronize()
    assert torch.equal(env.get_state(), want)
    env.check_status(); other.check_status(); other64.check_status()


def test_gust_noise_is_fresh_every_episode_and_generation(envs_mod):
    """Upstream's reset() re-seeds and rebuilds the gust filters (vonkarman.py:86-96): consecutive
    episodes of one env, and consecutive PSO generations, must see different noise; the stream of
    a particle must not depend on how the swarm is sharded (index0)."""
    B = 64
    env = envs_mod.BatchedRocketEnv(B, "pso", P, enable_wind=True, stochastic_wind=True, precision="fp64", seed=3)
    dbg = torch.zeros(B, 16, dtype=torch.float64, device="cuda")
    a = torch.full((B, 1), 0.8, dtype=torch.float32, device="cuda")    # brakes: below 15 km after ~17 s, where gusts act
    runs = []
    for episode in range(2):
        env.reset()
        vg = []
        for t in range(300):
            obs, rew, done, trunc, tid = env.step(a, dbg=dbg)
            assert not bool((done | trunc).any())
            vg.append(dbg[:, 14].clone())                # v-gust: pure filter output, no altitude profile
        runs.append(torch.stack(vg))
    first, second = runs[0][:, 0], runs[1][:, 0]
    active = (first != 0) & (second != 0)
    assert int(active.sum()) > 50
    assert not torch.equal(first, second)                               # not a replay of episode 1 ...
    ratio = second[active] / first[active]
    assert float(ratio.std()) > 1e-2 * float(ratio.abs().mean())        # ... and not a rescaled replay either
    assert not torch.equal(runs[0][:, 0], runs[0][:, 1])                # envs differ among themselves
    # rollouts: generation-dependent noise, sharding-independent streams
    rng = np.random.default_rng(4)
    pos = torch.as_tensor(rng.uniform(-1.5, 1.5, (16, 249)).astype(np.float32)).cuda()
    m = envs_mod.pso_wrapped_env(flight_phase=P, enable_wind=True, stochastic_wind=True, precision="fp32", seed=11)
    f0, _, _ = m._b.rollout_pso(pos, n_seeds=4, generation=0)
    f0b, _, _ = m._b.rollout_pso(pos, n_seeds=4, generation=0)
    f1, _, _ = m._b.rollout_pso(pos, n_seeds=4, generation=1)
    assert torch.equal(f0, f0b) and not torch.equal(f0, f1)
    fs, _, _ = m._b.rollout_pso(pos[8:], n_seeds=4, generation=0, index0=8)
    assert torch.equal(fs, f0[8 * 4:])


def test_step_cap_is_scored_as_truncation(envs_mod):
    """The reference's episode loop has no step cap; an episode cut by max_steps keeps truncation id -1
    and is scored with the closures' truncated branch at its final state (P: +|y| while airborne), so a
    stalling policy cannot outrank a crash."""
    rng = np.random.default_rng(5)
    pos = rng.uniform(-1.5, 1.5, (256, 249))
    m = envs_mod.pso_wrapped_env(flight_phase=P, precision="fp64", max_steps=40)
    with pytest.warns(RuntimeWarning, match="max_steps"):
        fit, steps, tid, term = m.evaluate(pos, terminal=True)
    assert m.capped == 256 and (tid == -1).all() and (steps == 40).all()
    y = term[:, 1]
    assert (y > 0).all() and torch.allclose(fit, y.abs(), rtol=1e-12)
    m2 = envs_mod.pso_wrapped_env(flight_phase=P, precision="fp64", max_steps=4096)
    fit2, steps2, tid2 = m2.evaluate(pos[:32])
    assert (tid2 >= 0).all() and m2.capped == 0
