"""Generate tests/golden/*.npz by running the UNMODIFIED reference (container-only).

Every array written here is an output of the reference's own code
(rocket_environment_pre_wrap, pso_wrapped_env, rl_wrapped_env_pytorch, LandingBurn
loop) imported through tools/ref_harness.py; inputs (states, actions, weights,
noise tapes) are stored alongside so the oracle and the CUDA path can be fed the
identical data on a box where /root/reference does not exist.

    python tools/make_golden.py            # rewrites tests/golden/
"""
import io
import contextlib
import json
import math
import os
import random
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tools.ref_harness import load_reference, NoiseTapeWind  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
P, G = "landing_burn_pure_throttle", "landing_burn"


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def info_row(info):
    ai = info["action_info"]
    return [info["mach_number"], info["dynamic_pressure"], info["CL"], info["CD"],
            info["air_density"], info["atmospheric_pressure"], info["speed_of_sound"],
            info["x_cog"], info["inertia"], float(info["mass_flow"]), float(ai["throttle"]),
            info["alpha_effective"], info["g_load_1_sec_window"]]


INFO_COLS = ["mach", "q", "CL", "CD", "rho", "p_atm", "a", "x_cog", "inertia", "mass_flow",
             "throttle", "alpha_eff", "g1"]


def tape_replay():
    import pandas as pd
    from src.envs.base_environment import rocket_environment_pre_wrap
    g = pd.read_csv("data/reference_trajectory/landing_burn_controls_pure_throttle/"
                    "state_action_landing_burn_pure_throttle_control.csv")
    env = rocket_environment_pre_wrap(type="pso", flight_phase=P, enable_wind=False)
    env.reset()
    u = g["u0"].values
    S, R, D, T, TID, INFO = [], [], [], [], [], []
    for k in range(len(u)):
        s, r, d, t, info = quiet(env.step, np.array([[u[k]]]))
        S.append([float(v) for v in s]); R.append(float(r)); D.append(d); T.append(t)
        TID.append(env.truncation_id); INFO.append(info_row(info))
        if d or t:
            break
    csv_cols = ["x[m]", "y[m]", "vx[m/s]", "vy[m/s]", "theta[rad]", "theta_dot[rad/s]",
                "gamma[rad]", "alpha[rad]", "mass[kg]", "masspropellant[kg]", "time[s]"]
    np.savez_compressed(os.path.join(OUT, "p_tape_replay.npz"), u0=u, states=np.array(S),
                        rewards=np.array(R), done=np.array(D), truncated=np.array(T),
                        trunc_id=np.array(TID), info=np.array(INFO), info_cols=INFO_COLS,
                        csv_states=g[csv_cols].values)
    print("tape_replay", len(S), "steps; final reward", R[-1], "done", D[-1])


def _collect_states(phase, rng, n_roll, act_dim):
    """States visited by short random-policy rollouts (reference env, pso rtd)."""
    from src.envs.base_environment import rocket_environment_pre_wrap
    env = rocket_environment_pre_wrap(type="pso", flight_phase=phase, enable_wind=False)
    out = []
    for r in range(n_roll):
        env.reset()
        bias = rng.uniform(-1, 1, act_dim)
        prev = env.state
        win = []
        for k in range(3000):
            a = np.clip(bias + 0.5 * rng.uniform(-1, 1, act_dim), -1, 1)
            s, rew, d, t, info = quiet(env.step, a.astype(np.float64))
            prevs = (0.0, 0.0, 0.0)
            if phase == G:
                prevs = (env.gimbal_angle_deg_prev, env.delta_command_left_rad_prev,
                         env.delta_command_right_rad_prev)
            out.append(([float(v) for v in s], [float(v) for v in prev], list(env.g_loads_window),
                        [float(v) for v in prevs]))
            prev = s
            if d or t:
                break
    return out


def single_step(phase, tag, n=192, seed=0):
    """(state, prev_state, g-window, actuator prevs, action) -> one reference step."""
    from src.envs.base_environment import rocket_environment_pre_wrap
    rng = np.random.default_rng(seed)
    act_dim = 1 if phase == P else 4
    pool = _collect_states(phase, rng, 6 if phase == P else 24, act_dim)
    env = rocket_environment_pre_wrap(type="pso", flight_phase=phase, enable_wind=False)
    rows = dict(state=[], prev=[], win=[], nwin=[], aprev=[], act32=[], act64=[],
                o64=[], o32=[])
    for i in range(n):
        s, sp, win, aprev = pool[rng.integers(len(pool))]
        s = np.array(s); sp = np.array(sp)
        if i % 6 == 1:      # touchdown window: exercises done (0<y<1, slow) and y<0
            s = s.copy()
            s[1] = rng.uniform(0.05, 2.5)
            s[3] = -rng.uniform(0.5, 7.0)
            s[2] = rng.uniform(-1.0, 1.0)
            s[0] = rng.uniform(-1.0, 1.0) if phase == G else s[0]
            s[6] = math.atan2(s[3], s[2]) % (2 * math.pi)
            s[7] = s[4] - s[6]
        elif i % 3 == 1:    # perturb: low-altitude / slow cases to exercise done / y<0 branches
            s = s.copy()
            s[1] = rng.uniform(-2.0, 40.0)
            s[3] = -rng.uniform(0.5, 30.0)
            s[2] = rng.uniform(-3.0, 3.0)
            s[6] = math.atan2(s[3], s[2]) % (2 * math.pi)
            s[7] = s[4] - s[6]
        elif i % 3 == 2:    # generic perturbation
            s = s * (1 + 0.02 * rng.standard_normal(11))
            s[6] = math.atan2(s[3], s[2]) % (2 * math.pi)
            s[7] = s[4] - s[6]
        a64 = rng.uniform(-1, 1, act_dim)
        a32 = a64.astype(np.float32)
        win = list(win)
        outs = {}
        for key, a in (("o64", a32.astype(np.float64)), ("o32", a32)):
            env.reset()
            env.state = [np.float64(v) for v in s]
            # invariant of the reference: previous_state is the state the step starts from
            env.previous_state = env.state
            env.g_loads_window = list(win)
            if phase == G:
                env.gimbal_angle_deg_prev = aprev[0]
                env.delta_command_left_rad_prev = aprev[1]
                env.delta_command_right_rad_prev = aprev[2]
            ns, r, d, t, info = quiet(env.step, a)
            ap = [0.0, 0.0, 0.0]
            if phase == G:
                ap = [float(env.gimbal_angle_deg_prev), float(env.delta_command_left_rad_prev),
                      float(env.delta_command_right_rad_prev)]
            outs[key] = [float(v) for v in ns] + [float(r), float(d), float(t),
                                                  float(env.truncation_id)] + ap + info_row(info)
        rows["state"].append(s); rows["prev"].append(sp)
        w = np.zeros(10); w[:len(win)] = win
        rows["win"].append(w); rows["nwin"].append(len(win)); rows["aprev"].append(aprev)
        rows["act32"].append(a32); rows["act64"].append(a32.astype(np.float64))
        rows["o64"].append(outs["o64"]); rows["o32"].append(outs["o32"])
    cols = ["x", "y", "vx", "vy", "theta", "theta_dot", "gamma", "alpha", "mass", "m_prop", "time",
            "reward", "done", "truncated", "trunc_id", "gimbal_prev", "dl_prev", "dr_prev"] + INFO_COLS
    np.savez_compressed(os.path.join(OUT, f"single_step_{tag}.npz"),
                        **{k: np.array(v) for k, v in rows.items()}, out_cols=cols)
    o = np.array(rows["o64"])
    print("single_step", tag, n, "done", int(o[:, 12].sum()), "trunc ids",
          np.unique(o[:, 14], return_counts=True))


def pso_fitness(phase, tag, n=12, seed=0):
    """objective_function of randomly initialised particles + a conditioning probe: the same
    episode replayed through the reference env with every action nudged by one float32 ulp.
    Episodes whose length changes under that nudge are ill-conditioned in the reference itself
    (phase G tumbles within ~20 steps) and are only sanity-checked by the parity tests."""
    from src.envs.pso.env_wrapped_ea import pso_wrapped_env
    from src.envs.base_environment import rocket_environment_pre_wrap
    model = pso_wrapped_env(flight_phase=phase, enable_wind=False)
    random.seed(seed)
    # exactly ParticleSubswarmOptimisation.initialize_swarms' draw order
    pos = [np.array([random.uniform(b[0], b[1]) for b in model.bounds]) for _ in range(n)]
    env = rocket_environment_pre_wrap(type="pso", flight_phase=phase, enable_wind=False)
    adim = 1 if phase == P else 4
    fit, steps, tid, term, cond, acts_all = [], [], [], [], [], []
    for p_ in pos:
        f = quiet(model.objective_function, p_)
        fit.append(float(f)); tid.append(model.env.truncation_id())
        term.append([float(v) for v in model.env.env.state])
        model.reset()
        # re-drive the same episode to record the float32 actions
        model.individual_update_model(p_)
        obs = model.env.reset()
        acts = []
        while True:
            a = model.actor.forward(obs)
            acts.append(a.detach().numpy().copy())
            obs, r, dn, tr, info = quiet(model.env.step, a)
            if dn or tr:
                break
        steps.append(len(acts))
        ok = True
        for direction in (2.0, -2.0):
            env.reset()
            tot, k = 0.0, 0
            for a in acts:
                a2 = np.nextafter(a.astype(np.float32), np.float32(direction)).astype(np.float32)
                s_, r, dn, tr, info = quiet(env.step, a2)
                tot -= r; k += 1
                if dn or tr:
                    break
            if k != len(acts) or not (dn or tr) or abs(tot - f) > 1e-6 * abs(f):
                ok = False
        cond.append(ok)
        pad = np.zeros((512, adim), np.float32)
        pad[:min(len(acts), 512)] = np.array(acts)[:512]
        acts_all.append(pad)
    np.savez_compressed(os.path.join(OUT, f"pso_fitness_{tag}.npz"), positions=np.array(pos),
                        fitness=np.array(fit), steps=np.array(steps), trunc_id=np.array(tid),
                        terminal_state=np.array(term), seed=seed, well_conditioned=np.array(cond),
                        actions=np.array(acts_all))
    print("pso_fitness", tag, "fitness", np.round(fit, 3), "steps", steps, "tid", tid, "well-conditioned", cond)


def _pso_many_worker(args):
    """One particle through the reference's own objective_function (fresh model state per call),
    plus the +-1-ulp action conditioning probe of pso_fitness()."""
    phase, p_, probe = args
    from src.envs.pso.env_wrapped_ea import pso_wrapped_env
    from src.envs.base_environment import rocket_environment_pre_wrap
    g = globals()
    key = ("_many_model", phase)
    if key not in g:
        g[key] = (quiet(pso_wrapped_env, flight_phase=phase, enable_wind=False),
                  quiet(rocket_environment_pre_wrap, type="pso", flight_phase=phase, enable_wind=False))
    model, env = g[key]
    model.reset()
    model.individual_update_model(p_)
    obs = model.env.reset()
    acts, tot = [], 0.0
    while True:
        a = model.actor.forward(obs)
        acts.append(a.detach().numpy().copy())
        obs, r, dn, tr, info = quiet(model.env.step, a)
        tot -= r
        if dn or tr:
            break
    f, n = float(tot), len(acts)
    tid = model.env.truncation_id()
    term = [float(v) for v in model.env.env.state]
    ok = True
    if probe:
        for direction in (2.0, -2.0):
            env.reset()
            t2, k = 0.0, 0
            for a in acts:
                a2 = np.nextafter(a.astype(np.float32), np.float32(direction)).astype(np.float32)
                s_, r, dn, tr, info = quiet(env.step, a2)
                t2 -= r; k += 1
                if dn or tr:
                    break
            if k != n or not (dn or tr) or abs(t2 - f) > 1e-6 * abs(f):
                ok = False
    return f, n, tid, term, ok


def pso_fitness_many(phase, tag, n=256, seed=100, procs=None):
    """>= 256 randomly initialised particles per phase through the unmodified reference
    (objective_function semantics, Pool over particles as parallel_evaluate does): fitness, episode
    length, truncation id, terminal state and the +-1-ulp conditioning flag.  Pins the step-count
    agreement of the fp32 production build with a stated number."""
    import multiprocessing as mp
    import torch
    torch.set_num_threads(1)        # forked workers: no OpenMP pool to inherit
    from src.envs.pso.env_wrapped_ea import pso_wrapped_env
    model = quiet(pso_wrapped_env, flight_phase=phase, enable_wind=False)
    random.seed(seed)
    pos = [np.array([random.uniform(b[0], b[1]) for b in model.bounds]) for _ in range(n)]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs or os.cpu_count()) as pool:
        res = pool.map(_pso_many_worker, [(phase, p_, True) for p_ in pos], chunksize=4)
    np.savez_compressed(os.path.join(OUT, f"pso_many_{tag}.npz"), positions=np.array(pos),
                        fitness=np.array([r[0] for r in res]), steps=np.array([r[1] for r in res]),
                        trunc_id=np.array([r[2] for r in res]), terminal_state=np.array([r[3] for r in res]),
                        well_conditioned=np.array([r[4] for r in res]), seed=seed)
    st = np.array([r[1] for r in res])
    print("pso_many", tag, n, "particles; steps min/median/max", st.min(), int(np.median(st)), st.max(),
          "well-conditioned", int(sum(r[4] for r in res)), "ids",
          np.unique([r[2] for r in res], return_counts=True))


BATCH_KEEP = (0, 1, 3, 7, 15, 31, 63)


def _batch_tape_worker(args):
    phase, acts32, use32 = args
    from src.envs.base_environment import rocket_environment_pre_wrap
    g = globals()
    key = ("_batch_env", phase)
    if key not in g:
        g[key] = quiet(rocket_environment_pre_wrap, type="pso", flight_phase=phase, enable_wind=False)
    env = g[key]
    env.reset()
    T = len(acts32)
    flags = np.zeros((T, 3), np.int8)
    rew = np.zeros(T)
    kept = np.full((len(BATCH_KEEP), 11), np.nan)
    n = 0
    last = None
    for k in range(T):
        a = acts32[k] if use32 else acts32[k].astype(np.float64)
        s, r, d, t, info = quiet(env.step, a)
        flags[k] = (int(d), int(t), int(env.truncation_id))
        rew[k] = float(r)
        last = [float(v) for v in s]
        if k in BATCH_KEEP:
            kept[BATCH_KEEP.index(k)] = last
        n = k + 1
        if d or t:
            break
    return flags, rew, kept, n, last


def batch_tape(phase, tag, n_envs=1024, n_steps=64, seed=200, procs=None):
    """SURVEY 8(d) config-2 parity subset: n_envs x n_steps random U(-1,1) actions from reset through
    the unmodified reference env (pso closures), once with the float64 tape and once with the same
    values as float32 (NEP-50 path).  Episodes stop at done / truncated.  Kept: every step's flags
    and reward, the state at steps BATCH_KEEP and the last state."""
    import multiprocessing as mp
    rng = np.random.default_rng(seed)
    adim = 1 if phase == P else 4
    acts = rng.uniform(-1, 1, (n_envs, n_steps, adim)).astype(np.float32)
    # a quarter of the envs get a biased policy so that episodes live longer / end differently
    bias = rng.uniform(-0.8, 0.8, (n_envs, 1, adim)).astype(np.float32)
    acts[::4] = np.clip(bias[::4] + 0.3 * acts[::4], -1, 1)
    ctx = mp.get_context("fork")
    out = {"actions": acts, "keep_steps": np.array(BATCH_KEEP)}
    with ctx.Pool(procs or os.cpu_count()) as pool:
        for key, use32 in (("f64", False), ("f32", True)):
            res = pool.map(_batch_tape_worker, [(phase, acts[i], use32) for i in range(n_envs)], chunksize=8)
            out[f"flags_{key}"] = np.array([r[0] for r in res])
            out[f"rewards_{key}"] = np.array([r[1] for r in res])
            out[f"states_{key}"] = np.array([r[2] for r in res])
            out[f"steps_{key}"] = np.array([r[3] for r in res])
            out[f"last_{key}"] = np.array([r[4] for r in res])
    np.savez_compressed(os.path.join(OUT, f"batch_tape_{tag}.npz"), **out)
    print("batch_tape", tag, n_envs, "envs; steps f64 min/median/max", out["steps_f64"].min(),
          int(np.median(out["steps_f64"])), out["steps_f64"].max(), "ended",
          int((out["steps_f64"] < n_steps).sum()), "f32 length differs in",
          int((out["steps_f64"] != out["steps_f32"]).sum()))


def pso_best_actor():
    """The reference's own saved best P actor: weights -> fitness / trajectory."""
    import pandas as pd
    from src.envs.pso.env_wrapped_ea import pso_wrapped_env
    d = "data/pso_saves/landing_burn_pure_throttle/PSO_different_starting_point/"
    res = pd.read_csv(d + "particle_subswarm_optimisation_results.csv")
    w = res.iloc[0].values[1:-1].astype(float)      # col 0 = 'Algorithm', last = 'Best Fitness'
    stored = float(res.iloc[0].values[-1])
    model = pso_wrapped_env(flight_phase=P, enable_wind=False)
    # record the per-step trajectory by driving the loop ourselves with the reference objects
    model.individual_update_model(w)
    obs = model.env.reset()
    S, A, R = [], [], []
    tot = 0
    while True:
        a = model.actor.forward(obs)
        obs, r, dn, tr, info = quiet(model.env.step, a)
        S.append([float(v) for v in model.env.env.state]); A.append(float(a.detach().numpy()[0]))
        R.append(float(r)); tot -= r
        if dn or tr:
            break
    ref_states = pd.read_csv(d + "trajectory_data/states.csv").values
    ref_actions = pd.read_csv(d + "trajectory_data/actions.csv").values
    np.savez_compressed(os.path.join(OUT, "pso_best_actor_P.npz"), weights=w, stored_fitness=stored,
                        fitness=float(tot), steps=len(S), states=np.array(S), actions=np.array(A),
                        rewards=np.array(R), trunc_id=model.env.truncation_id(),
                        stored_states=ref_states, stored_actions=ref_actions)
    print("best actor: stored", stored, "re-run", float(tot), "steps", len(S))


def rl_sequence(phase, tag, n_steps, seed=1):
    from src.envs.rl.env_wrapped_rl_pytorch import rl_wrapped_env_pytorch
    rng = np.random.default_rng(seed)
    env = rl_wrapped_env_pytorch(flight_phase=phase, enable_wind=False, trajectory_length=1,
                                 discount_factor=0.99)
    act_dim = env.action_dim
    obs0 = env.reset()
    O, R, D, T, A, S = [np.array(obs0, float)], [], [], [], [], []
    bias = 0.6 if phase == P else 0.0
    for k in range(n_steps):
        a = np.clip(bias + 0.4 * rng.uniform(-1, 1, act_dim), -1, 1).astype(np.float32)
        o, r, d, t, info = quiet(env.step, a)
        O.append(np.array(o, float)); R.append(r); D.append(d); T.append(t); A.append(a)
        S.append([float(v) for v in env.env.state])
        if d or t:
            break
    np.savez_compressed(os.path.join(OUT, f"rl_sequence_{tag}.npz"), actions=np.array(A),
                        obs=np.array(O), rewards=np.array(R), done=np.array(D),
                        truncated=np.array(T), states=np.array(S),
                        trunc_id=env.truncation_id())
    print("rl_sequence", tag, len(R), "steps, last reward", R[-1], "trunc", T[-1], env.truncation_id())


S_, U_, B_, C_ = "subsonic", "supersonic", "ballistic_arc_descent", "landing_burn_pure_throttle_Pcontrol"


def rl_sequence_other(phase, tag, n_steps, seed=11, mode="random"):
    """RL-wrapper sequences of the phases outside the landing burns.
    mode 'csv': replay the reference's own recorded controller actions (float32, as a torch
    policy would hand them over); 'random'; 'track' (phase C: follow the v_opt(y) profile)."""
    import pandas as pd
    from src.envs.rl.env_wrapped_rl_pytorch import rl_wrapped_env_pytorch
    rng = np.random.default_rng(seed)
    env = quiet(rl_wrapped_env_pytorch, flight_phase=phase, enable_wind=False, trajectory_length=1000,
                discount_factor=0.99)
    act_dim = env.action_dim
    csv_states = np.zeros((0, 11))
    if mode == "csv":
        f = {S_: "subsonic_state_action_ascent_control.csv",
             U_: "supersonic_state_action_ascent_control.csv"}[phase]
        g = pd.read_csv("data/reference_trajectory/ascent_controls/" + f)
        tape = g[["u0", "u1"]].values.astype(np.float32)
        cols = ["x[m]", "y[m]", "vx[m/s]", "vy[m/s]", "theta[rad]", "theta_dot[rad/s]", "gamma[rad]",
                "alpha[rad]", "mass[kg]", "mass_propellant[kg]", "time[s]"]
        csv_states = g[cols].values
        n_steps = min(n_steps, len(tape))
    obs0 = env.reset()
    O, R, D, T, A, S = [np.array(obs0, float)], [], [], [], [], []
    a_opt, b_opt = -7.445479767703873e-07, 0.056435801747276214
    for k in range(n_steps):
        if mode == "csv":
            a = tape[k]
        elif mode == "track":
            y = float(env.env.state[1])
            a = np.array([np.clip(2 * (a_opt * y * y + b_opt * y) / env.speed0 - 1, -1, 1)], np.float32)
        else:
            a = rng.uniform(-1, 1, act_dim).astype(np.float32)
        o, r, d, t, info = quiet(env.step, a)
        O.append(np.array(o, float)); R.append(r); D.append(d); T.append(t); A.append(a)
        S.append([float(v) for v in env.env.state])
        if d or t:
            break
    np.savez_compressed(os.path.join(OUT, f"rl_sequence_{tag}.npz"), actions=np.array(A),
                        obs=np.array(O), rewards=np.array(R), done=np.array(D),
                        truncated=np.array(T), states=np.array(S), trunc_id=env.truncation_id(),
                        csv_states=csv_states, obs_dtype=str(np.asarray(o).dtype))
    print("rl_sequence", tag, len(R), "steps, last reward", R[-1], "done", D[-1], "trunc", T[-1],
          env.truncation_id())
    return np.array(S)


def single_step_other(phase, tag, pool, n=96, seed=21):
    """(state, g-window, action) -> one step of the reference base env (type='rl'), float64 and
    float32 action arrays (NEP-50 paths of the ascent / RCS / P-control decomposers)."""
    from src.envs.base_environment import rocket_environment_pre_wrap
    rng = np.random.default_rng(seed)
    env = quiet(rocket_environment_pre_wrap, type="rl", flight_phase=phase, enable_wind=False,
                trajectory_length=1000, discount_factor=0.99)
    act_dim = {S_: 2, U_: 2, B_: 1, C_: 1}[phase]
    rows = dict(state=[], win=[], nwin=[], act32=[], act64=[], o64=[], o32=[])
    for i in range(n):
        s = np.array(pool[rng.integers(len(pool))], float)
        if i % 3 == 1:
            s = s * (1 + 0.01 * rng.standard_normal(11))
            s[6] = math.atan2(s[3], s[2]) % (2 * math.pi)
            s[7] = s[4] - s[6]
        elif i % 3 == 2 and phase == C_:
            s[1] = rng.uniform(-12.0, 150.0); s[3] = -rng.uniform(0.1, 60.0); s[2] = rng.uniform(-2, 2)
            s[6] = math.atan2(s[3], s[2]) % (2 * math.pi)
            s[7] = s[4] - s[6]
        nwin = int(rng.integers(0, 11))
        win = list(rng.uniform(0.0, 0.7 if i % 5 else 7.0, nwin))
        if phase == C_:
            a32 = rng.uniform(0.0, 1100.0, act_dim).astype(np.float32)     # v_ref [m/s]
        else:
            a32 = rng.uniform(-1, 1, act_dim).astype(np.float32)
        outs = {}
        for key, a in (("o64", a32.astype(np.float64)), ("o32", a32)):
            quiet(env.reset)
            env.state = [np.float64(v) for v in s]
            env.previous_state = env.state
            env.g_loads_window = list(win)
            ns, r, d, t, info = quiet(env.step, a)
            outs[key] = [float(v) for v in ns] + [float(r), float(d), float(t), float(env.truncation_id),
                                                  info["mach_number"], info["dynamic_pressure"],
                                                  info["CL"], info["CD"], info["x_cog"], info["inertia"],
                                                  float(info["mass_flow"]), info["g_load_1_sec_window"]]
        w = np.zeros(10); w[:nwin] = win
        rows["state"].append(s); rows["win"].append(w); rows["nwin"].append(nwin)
        rows["act32"].append(a32); rows["act64"].append(a32.astype(np.float64))
        rows["o64"].append(outs["o64"]); rows["o32"].append(outs["o32"])
    cols = ["x", "y", "vx", "vy", "theta", "theta_dot", "gamma", "alpha", "mass", "m_prop", "time",
            "reward", "done", "truncated", "trunc_id", "mach", "q", "CL", "CD", "x_cog", "inertia",
            "mass_flow", "g1"]
    np.savez_compressed(os.path.join(OUT, f"single_step_{tag}.npz"),
                        **{k: np.array(v) for k, v in rows.items()}, out_cols=cols)
    o = np.array(rows["o64"])
    print("single_step", tag, n, "done", int(o[:, 12].sum()), "trunc ids",
          np.unique(o[:, 14], return_counts=True))


INFO_FULL_KEYS = ["mach_number", "mach_number_max", "CL", "CD", "drag", "lift", "d_cp_cg", "d_thrust_cg",
                  "x_cog", "inertia", "dynamic_pressure", "mass_flow", "fuel_percentage_consumed",
                  "control_force_parallel", "control_force_perpendicular", "control_force_x",
                  "control_force_y", "aero_force_x", "aero_force_y", "gravity_force_y",
                  "atmospheric_pressure", "air_density", "speed_of_sound", "ug", "vg", "alpha_effective",
                  "g_load_1_sec_window"]
ACC_KEYS = ["acceleration_x_component_control", "acceleration_y_component_control",
            "acceleration_x_component_drag", "acceleration_y_component_drag",
            "acceleration_x_component_lift", "acceleration_y_component_lift",
            "acceleration_x_component_gravity", "acceleration_y_component_gravity",
            "acceleration_x_component", "acceleration_y_component", "acceleration_x_component_wind",
            "acceleration_y_component_wind"]
MOM_KEYS = ["control_moment_z", "aero_moment_z", "moments_z", "theta_dot_dot", "M_wind_z"]


def info_full(n_steps=25, seed=31):
    """The complete per-step `info` dict of the reference (rockets_physics.py:649-702) along short
    episodes of four phases: pins the trajectory / info export."""
    from src.envs.base_environment import rocket_environment_pre_wrap
    rng = np.random.default_rng(seed)
    out = {}
    for tag, phase, typ, adim in (("P", P, "pso", 1), ("G", G, "pso", 4), ("U", U_, "rl", 2), ("B", B_, "rl", 1)):
        env = quiet(rocket_environment_pre_wrap, type=typ, flight_phase=phase, enable_wind=False,
                    trajectory_length=1000, discount_factor=0.99)
        quiet(env.reset)
        A, V, ACT = [], [], []
        for k in range(n_steps):
            a = (0.3 * rng.uniform(-1, 1, adim)).astype(np.float64)
            s, r, d, t, info = quiet(env.step, a)
            row = [float(info[k_]) for k_ in INFO_FULL_KEYS]
            row += [float(info["acceleration_dict"][k_]) for k_ in ACC_KEYS]
            row += [float(info["moment_dict"][k_]) for k_ in MOM_KEYS]
            ai = info["action_info"]
            ACT.append([float(ai.get("throttle", 0.0) or 0.0), float(ai.get("gimbal_angle_deg", 0.0)),
                        float(ai.get("delta_command_left_rad", 0.0)), float(ai.get("delta_command_right_rad", 0.0))])
            A.append(a); V.append(row)
            if d or t:
                break
        out[f"actions_{tag}"] = np.array(A); out[f"values_{tag}"] = np.array(V)
        out[f"action_info_{tag}"] = np.array(ACT)
    np.savez_compressed(os.path.join(OUT, "info_full.npz"), keys=INFO_FULL_KEYS, acc_keys=ACC_KEYS,
                        mom_keys=MOM_KEYS, **out)
    print("info_full", {k: v.shape for k, v in out.items() if k.startswith("values")})


def stored_info_csv(n_rows=120):
    """The reference's own committed per-step `info_data.csv` of its saved best P actor (written on
    the author's machine by save_trajectory_data, particle_swarm_optimisation.py:787-809): numeric
    columns of the first rows, with the stored actions that produced them."""
    import pandas as pd
    d = "data/pso_saves/landing_burn_pure_throttle/PSO_different_starting_point/trajectory_data/"
    info = pd.read_csv(d + "info_data.csv")
    acts = pd.read_csv(d + "actions.csv").values[:n_rows]
    cols = [c for c in info.columns if info[c].dtype.kind == "f" or info[c].dtype.kind == "i"]
    np.savez_compressed(os.path.join(OUT, "stored_info_P.npz"), columns=np.array(cols),
                        values=info[cols].values[:n_rows].astype(float), actions=acts.astype(float),
                        all_columns=np.array(list(info.columns)))
    print("stored_info", len(cols), "numeric columns of", len(info.columns), "rows", n_rows)


def supervisory_steps(n=64, seed=41):
    """type='supervisory' closures (rtd_supervisory_mock.py): one reference step from states of the
    existing single-step fixtures, all six working phases."""
    from src.envs.base_environment import rocket_environment_pre_wrap
    rng = np.random.default_rng(seed)
    out = {}
    for tag, phase in (("P", P), ("G", G), ("S", S_), ("U", U_), ("B", B_), ("C", C_)):
        g = np.load(os.path.join(OUT, f"single_step_{tag}.npz"), allow_pickle=True)
        env = quiet(rocket_environment_pre_wrap, type="supervisory", flight_phase=phase, enable_wind=False)
        idx = rng.choice(len(g["state"]), size=min(n, len(g["state"])), replace=False)
        rows = []
        for i in idx:
            quiet(env.reset)
            env.state = [np.float64(v) for v in g["state"][i]]
            env.previous_state = env.state
            env.g_loads_window = [float(v) for v in g["win"][i][:g["nwin"][i]]]
            if phase == G:
                env.gimbal_angle_deg_prev, env.delta_command_left_rad_prev, env.delta_command_right_rad_prev = \
                    [float(v) for v in g["aprev"][i]]
            try:
                ns, r, d, t, info = quiet(env.step, g["act64"][i])
                rows.append([float(v) for v in ns] + [float(r), float(d), float(t), float(env.truncation_id)])
            except NameError:      # upstream bug: the g-load branch prints an undefined name
                rows.append([np.nan] * 11 + [0.0, np.nan, 1.0, 5.0])
        out[f"idx_{tag}"] = idx; out[f"out_{tag}"] = np.array(rows)
    np.savez_compressed(os.path.join(OUT, "supervisory_step.npz"), **out)
    print("supervisory", {k: v.shape for k, v in out.items() if k.startswith("out")},
          {k: np.nansum(v[:, 13]) for k, v in out.items() if k.startswith("out")})


def supervisory_wrapper_sequences(n_steps=8, seed=43):
    """supervisory_wrapper (env_wrapped_supervisory.py): reset observation + a few steps per phase with
    the reference's own normalisation values (an 8-vector of ones-free made-up scales for landing_burn,
    whose stock 7-vector does not broadcast upstream)."""
    from src.envs.supervisory.env_wrapped_supervisory import supervisory_wrapper
    from src.envs.utils.input_normalisation import find_input_normalisation_vals
    rng = np.random.default_rng(seed)
    out = {}
    for tag, phase, adim in (("P", P, 1), ("G", G, 4), ("S", S_, 2), ("U", U_, 2), ("B", B_, 1), ("C", C_, 1)):
        nv = quiet(find_input_normalisation_vals, phase)
        if phase == G:
            nv = np.array([6e3, 4e4, 300.0, 1200.0, 2.0, 0.1, 3.5, 2e6])
        env = quiet(supervisory_wrapper, nv, flight_phase=phase)
        obs = [quiet(env.reset)]
        acts = rng.uniform(-1, 1, size=(n_steps, adim))
        if phase in (S_, U_):
            acts[:, 1] = rng.uniform(0.2, 1.0, size=n_steps)
        flags = []
        for a in acts:
            o, r, d, t, _ = quiet(env.step, a if adim > 1 else a)
            obs.append(np.asarray(o, dtype=np.float64).reshape(-1))
            flags.append([float(r), float(d), float(t), float(env.truncation_id())])
            if d or t:
                break
        out[f"nv_{tag}"] = np.asarray(nv, dtype=np.float64)
        out[f"act_{tag}"] = acts[:len(flags)]
        out[f"obs_{tag}"] = np.array([np.asarray(o, dtype=np.float64).reshape(-1) for o in obs])
        out[f"flags_{tag}"] = np.array(flags)
    np.savez_compressed(os.path.join(OUT, "supervisory_wrapper.npz"), **out)
    print("supervisory_wrapper", {k: v.shape for k, v in out.items() if k.startswith("obs")})


F_ = "flip_over_boostbackburn"


def flip_over(n=96, seed=51):
    """flip_over_boostbackburn (rockets_physics.py:63-92, 542-560, 755-780) under type='supervisory'
    - the one closure set that runs this phase upstream (rtd_supervisory_mock.py:34-38, 57-61):
    (1) the reference's committed controller recording (u0 + states per 0.1 s step) and its replay
    through env.step from reset; (2) single steps from perturbed states of that recording with
    float64 and float32 actions and a non-zero gimbal memory; (3) the supervisory_wrapper
    observation / step sequence."""
    import pandas as pd
    from src.envs.base_environment import rocket_environment_pre_wrap
    from src.envs.supervisory.env_wrapped_supervisory import supervisory_wrapper
    from src.envs.utils.input_normalisation import find_input_normalisation_vals
    rng = np.random.default_rng(seed)
    g = pd.read_csv("data/reference_trajectory/flip_over_and_boostbackburn_controls/"
                    "state_action_flip_over_and_boostbackburn_control.csv")
    cols = ["x[m]", "y[m]", "vx[m/s]", "vy[m/s]", "theta[rad]", "theta_dot[rad/s]", "gamma[rad]",
            "alpha[rad]", "mass[kg]", "mass_propellant[kg]", "time[s]"]
    out = {"csv_states": g[cols].values, "csv_u0": g["u0"].values,
           "csv_gimbal_cmd_deg": g["gimbalanglecommanded[deg]"].values}
    env = quiet(rocket_environment_pre_wrap, type="supervisory", flight_phase=F_, enable_wind=False)
    out["initial_state"] = np.array([float(v) for v in quiet(env.reset)])
    S, FL, GD = [], [], []
    for u in g["u0"].values:
        s, r, d, t, info = quiet(env.step, np.array([u]))
        S.append([float(v) for v in s]); FL.append([float(r), float(d), float(t), float(env.truncation_id)])
        GD.append(float(np.asarray(env.gimbal_angle_deg).reshape(-1)[0]))
        if d or t:
            break
    out["replay_states"], out["replay_flags"], out["replay_gimbal_deg"] = np.array(S), np.array(FL), np.array(GD)
    pool = np.array(S)
    rows = dict(state=[], win=[], nwin=[], aprev=[], act32=[], act64=[], o64=[], o32=[])
    for i in range(n):
        k = rng.integers(len(pool))
        s = pool[k].copy()
        if i % 3 == 1:
            s = s * (1 + 0.01 * rng.standard_normal(11))
            s[6] = math.atan2(s[3], s[2]) % (2 * math.pi)
            s[7] = s[4] - s[6]
        elif i % 3 == 2:          # around the done threshold vx < -60 and propellant exhaustion
            s[2] = rng.uniform(-75.0, -45.0)
            s[6] = math.atan2(s[3], s[2]) % (2 * math.pi)
            s[7] = s[4] - s[6]
            if i % 2:
                s[9] = rng.uniform(-500.0, 15000.0)
        gprev = float(np.float32(rng.uniform(-10, 10))) if i % 4 else 0.0      # float32-representable
        nwin = int(rng.integers(0, 11))
        win = list(rng.uniform(0.0, 3.0, nwin))
        a32 = rng.uniform(-1, 1, 1).astype(np.float32)
        outs = {}
        for key, a in (("o64", a32.astype(np.float64)), ("o32", a32)):
            quiet(env.reset)
            env.state = [np.float64(v) for v in s]
            env.previous_state = env.state
            env.g_loads_window = list(win)
            # the memory is what the previous step of the same dtype left behind
            env.gimbal_angle_deg = gprev if key == "o64" else np.array([gprev], dtype=np.float32)
            ns, r, d, t, info = quiet(env.step, a)
            outs[key] = [float(v) for v in ns] + [float(r), float(d), float(t), float(env.truncation_id),
                                                  float(np.asarray(env.gimbal_angle_deg).reshape(-1)[0]),
                                                  info["mach_number"], info["dynamic_pressure"], info["CL"],
                                                  info["CD"], info["x_cog"], info["inertia"],
                                                  float(info["mass_flow"]), info["g_load_1_sec_window"]]
        w = np.zeros(10); w[:nwin] = win
        rows["state"].append(s); rows["win"].append(w); rows["nwin"].append(nwin)
        rows["aprev"].append([float(np.float32(gprev)), 0.0, 0.0])
        rows["act32"].append(a32); rows["act64"].append(a32.astype(np.float64))
        rows["o64"].append(outs["o64"]); rows["o32"].append(outs["o32"])
    out.update({f"ss_{k}": np.array(v) for k, v in rows.items()})
    out["ss_cols"] = ["x", "y", "vx", "vy", "theta", "theta_dot", "gamma", "alpha", "mass", "m_prop", "time",
                      "reward", "done", "truncated", "trunc_id", "gimbal_deg", "mach", "q", "CL", "CD",
                      "x_cog", "inertia", "mass_flow", "g1"]
    nv = quiet(find_input_normalisation_vals, F_)
    wenv = quiet(supervisory_wrapper, nv, flight_phase=F_)
    obs = [np.asarray(quiet(wenv.reset), dtype=np.float64).reshape(-1)]
    acts = rng.uniform(-1, 1, size=(12, 1))
    flags = []
    for a in acts:
        o, r, d, t, _ = quiet(wenv.step, a)
        obs.append(np.asarray(o, dtype=np.float64).reshape(-1))
        flags.append([float(r), float(d), float(t), float(wenv.truncation_id())])
    out["w_nv"], out["w_act"], out["w_obs"], out["w_flags"] = np.asarray(nv, float), acts, np.array(obs), np.array(flags)
    np.savez_compressed(os.path.join(OUT, "flip_over.npz"), **out)
    o = np.array(rows["o64"])
    print("flip_over: replay", len(S), "steps, max |replay - csv|",
          float(np.max(np.abs(np.array(S) - g[cols].values[:len(S)]))), "; single steps", n, "done",
          int(o[:, 12].sum()), "trunc", int(o[:, 13].sum()))


def pso_run(seed=12, phase=P):
    """A short seeded run of the UNMODIFIED reference optimiser (ParticleSubswarmOptimisation.run,
    serial evaluation): 8 particles in 2 sub-swarms, 7 generations with sharing (every 2), migration
    (every 3) and the re-initialisation (generation 4).  Recorded per generation: the positions that
    were evaluated (per sub-swarm, in list order), the per-sub-swarm metrics and the global best;
    at the end the swarm dicts.  The reference writes its metrics under cwd, so it runs in a
    scratch directory whose `data/*`, `src`, `configs` are links into the read-only checkout."""
    import tempfile
    import configs.evolutionary_algorithms_config as cfg
    import src.particle_swarm_optimisation.particle_swarm_optimisation as ref_pso
    params = cfg.landing_burn_pure_throttle_pso_params if phase == P else cfg.landing_burn_pso_params
    saved = dict(params)
    knobs = dict(pop_size=8, generations=7, communication_freq=2, migration_freq=3, re_initialise_generation=4,
                 re_initialise_number_of_particles=6)
    params.update(knobs)
    root = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="pd_pso_run_")
    os.makedirs(os.path.join(tmp, "data"))
    for name in os.listdir(os.path.join(root, "data")):
        if name != "pso_saves":
            os.symlink(os.path.join(root, "data", name), os.path.join(tmp, "data", name))
    os.makedirs(os.path.join(tmp, "data", "pso_saves"))
    for name in ("src", "configs"):
        os.symlink(os.path.join(root, name), os.path.join(tmp, name))
    os.chdir(tmp)
    try:
        random.seed(seed)
        np.random.seed(seed)
        opt = quiet(ref_pso.ParticleSubswarmOptimisation, flight_phase=phase, save_interval=10 ** 6,
                    enable_wind=False, use_multiprocessing=False)
        gens = []
        orig = opt.save_generation_metrics

        def record(metrics, generation):
            gens.append(dict(
                positions=[np.array([p["position"] for p in sw]) for sw in opt.swarms],
                best_fitness=[np.array([p["best_fitness"] for p in sw]) for sw in opt.swarms],
                swarm_metrics=[[m["best_fitness"], m["avg_fitness"], m["min_fitness"], m["max_fitness"],
                                m["std_fitness"], m["num_particles"]] for m in metrics["swarm_metrics"]],
                global_best=metrics["global_best_fitness"], global_avg=metrics["global_avg_fitness"]))
            orig(metrics, generation)
        opt.save_generation_metrics = record
        best_pos, best_fit = quiet(opt.run)
        files = {}
        for f in ("subswarm_0_metrics.csv", "global_metrics.csv"):
            files[f] = open(os.path.join(opt.metrics_dir, f)).read()
    finally:
        os.chdir(root)
        params.clear(); params.update(saved)
    out = dict(seed=seed, phase=phase, knobs=json.dumps(knobs), n_generations=len(gens),
               global_best_fitness_array=np.array(opt.global_best_fitness_array),
               global_best_position=np.array(best_pos), global_best_fitness=best_fit,
               final_positions=np.array([p["position"] for sw in opt.swarms for p in sw]),
               final_velocities=np.array([p["velocity"] for sw in opt.swarms for p in sw]),
               final_sizes=np.array([len(sw) for sw in opt.swarms]),
               metrics_csv_subswarm_0=files["subswarm_0_metrics.csv"], metrics_csv_global=files["global_metrics.csv"])
    for g, rec in enumerate(gens):
        for k in range(len(rec["positions"])):
            out[f"g{g}_pos_{k}"] = rec["positions"][k]
            out[f"g{g}_pbest_{k}"] = rec["best_fitness"][k]
        out[f"g{g}_metrics"] = np.array(rec["swarm_metrics"], dtype=float)
        out[f"g{g}_global"] = np.array([rec["global_best"], rec["global_avg"]])
    np.savez_compressed(os.path.join(OUT, "pso_run_reference.npz"), **out)
    print("pso_run:", len(gens), "generations; global best history", opt.global_best_fitness_array,
          "final sizes", [len(sw) for sw in opt.swarms])


def ascent_csv():
    """The reference's own committed ascent controller recordings (actions + states per 0.1 s
    step) - golden vectors written on the author's machine, copied verbatim."""
    import pandas as pd
    cols = ["x[m]", "y[m]", "vx[m/s]", "vy[m/s]", "theta[rad]", "theta_dot[rad/s]", "gamma[rad]",
            "alpha[rad]", "mass[kg]", "mass_propellant[kg]", "time[s]"]
    out = {}
    for tag, f in (("S", "subsonic"), ("U", "supersonic")):
        g = pd.read_csv(f"data/reference_trajectory/ascent_controls/{f}_state_action_ascent_control.csv")
        out[f"actions_{tag}"] = g[["u0", "u1"]].values
        out[f"states_{tag}"] = g[cols].values
    np.savez_compressed(os.path.join(OUT, "ascent_csv.npz"), **out)
    print("ascent_csv", {k: v.shape for k, v in out.items()})


def wind_sequence(n_steps=260, seed=3):
    from src.envs.base_environment import rocket_environment_pre_wrap
    rng = np.random.default_rng(seed)
    tape = rng.standard_normal(8 * n_steps + 16)
    sig_u, sig_v = 1.7, 1.4
    nt = NoiseTapeWind(tape, sig_u, sig_v)
    nt.install()
    try:
        env = rocket_environment_pre_wrap(type="pso", flight_phase=P, enable_wind=True,
                                          stochastic_wind=True, horiontal_wind_percentile=50)
        nt.pos = 0
        env.reset()
        nt.pos = 0
        S, UG = [], []
        # strong braking so the episode gets below 15 km where the gust filter switches on
        for k in range(n_steps):
            a = np.array([0.9 + 0.1 * math.sin(0.05 * k)], dtype=np.float32)
            s, r, d, t, info = quiet(env.step, a)
            S.append([float(v) for v in s]); UG.append([float(info["ug"]), float(info["vg"])])
            if d or t:
                break
        Adu, Bdu = env.wind_generator.von_karman_generator_class.u_filter.Ad, \
            env.wind_generator.von_karman_generator_class.u_filter.Bd
        Adv, Bdv = env.wind_generator.von_karman_generator_class.v_filter.Ad, \
            env.wind_generator.von_karman_generator_class.v_filter.Bd
    finally:
        nt.uninstall()
    np.savez_compressed(os.path.join(OUT, "wind_sequence_P.npz"), tape=tape, sigma_u=sig_u,
                        sigma_v=sig_v, states=np.array(S), ug_vg=np.array(UG),
                        actions=np.array([0.9 + 0.1 * math.sin(0.05 * k) for k in range(len(S))],
                                         dtype=np.float32),
                        Adu=Adu, Bdu=Bdu, Adv=Adv, Bdv=Bdv, tape_used=nt.pos, percentile=50)
    print("wind_sequence", len(S), "steps; min y", min(s[1] for s in S), "tape used", nt.pos,
          "last ug", UG[-1])


def wind_sequence_ascent(n_steps=80, seed=13):
    """Supersonic ascent with stochastic wind (gust filter active below 15 km) driven by an explicit
    noise tape; recorded controller actions as float32 and as float64 (the float32 force sums of
    the ascent decomposer meet the np.float64 wind force here)."""
    import pandas as pd
    from src.envs.base_environment import rocket_environment_pre_wrap
    rng = np.random.default_rng(seed)
    tape = rng.standard_normal(2 * n_steps + 16)
    sig_u, sig_v = 1.9, 1.6
    g = pd.read_csv("data/reference_trajectory/ascent_controls/supersonic_state_action_ascent_control.csv")
    acts = g[["u0", "u1"]].values[:n_steps]
    out = {}
    for key, dt in (("f32", np.float32), ("f64", np.float64)):
        nt = NoiseTapeWind(tape, sig_u, sig_v)
        nt.install()
        try:
            env = quiet(rocket_environment_pre_wrap, type="rl", flight_phase=U_, enable_wind=True,
                        stochastic_wind=True, horiontal_wind_percentile=50, trajectory_length=1000,
                        discount_factor=0.99)
            nt.pos = 0
            quiet(env.reset)
            nt.pos = 0
            S, UG, R, FL = [], [], [], []
            for k in range(n_steps):
                s, r, d, t, info = quiet(env.step, acts[k].astype(dt))
                S.append([float(v) for v in s]); UG.append([float(info["ug"]), float(info["vg"])])
                R.append(float(r)); FL.append([float(d), float(t), float(env.truncation_id)])
                if d or t:
                    break
        finally:
            nt.uninstall()
        out[f"states_{key}"] = np.array(S); out[f"ug_vg_{key}"] = np.array(UG)
        out[f"rewards_{key}"] = np.array(R); out[f"flags_{key}"] = np.array(FL)
        out[f"tape_used_{key}"] = nt.pos
    np.savez_compressed(os.path.join(OUT, "wind_sequence_U.npz"), tape=tape, sigma_u=sig_u, sigma_v=sig_v,
                        actions=acts, percentile=50, **out)
    print("wind_sequence_U", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def classical():
    from src.classical_controls.landing_burn_pure_throttle import LandingBurn
    lb = quiet(LandingBurn, test_case="control")
    n = 0
    while lb.mass_propellant > 0 and lb.y > 1 and lb.dynamic_pressure < lb.max_q and n < 50000 \
            and lb.vy < 0 and lb.alpha_effective < math.degrees(5):
        quiet(lb.closed_loop_step)
        n += 1
    st = np.array([lb.x_vals, lb.y_vals, lb.vx_vals, lb.vy_vals, lb.theta_vals, lb.theta_dot_vals,
                   lb.gamma_vals, lb.alpha_vals, lb.mass_vals, lb.m_prop_vals, lb.time_vals]).T
    np.savez_compressed(os.path.join(OUT, "classical_rollout_P.npz"), steps=n, states=st.astype(float),
                        u0=np.array(lb.u0_vals, float))
    print("classical", n, "steps; final", st[-1])


def aero_probe(seed=5, n=400):
    """C_D / C_L through the reference's own compiled closures at random (Mach, alpha_eff)."""
    import src.envs.rockets_physics as rp
    rng = np.random.default_rng(seed)
    mach = rng.uniform(0.0, 6.0, n)
    mach[::7] = rng.uniform(0.0, 10.0, len(mach[::7]))
    alpha = rng.uniform(-4e-3, 4e-3, n)
    alpha[::5] = rng.uniform(-0.3, 0.3, len(alpha[::5]))
    CL_func = lambda M, a: rp.rocket_CL(M, math.degrees(a))
    CD_func = lambda M, a: rp.rocket_CD(M, math.degrees(a))
    cl = np.array([CL_func(m, a) for m, a in zip(mach, alpha)])
    cd = np.array([CD_func(m, a) for m, a in zip(mach, alpha)])
    from src.envs.utils.acs_model import Ca_func, Cn_func
    ca = np.array([float(Ca_func(m)) for m in mach])
    cn = np.array([float(Cn_func(m, a)) for m, a in zip(mach, alpha)])
    from src.envs.utils.atmosphere_dynamics import endo_atmospheric_model
    alt = rng.uniform(-100, 90000, n)
    atm = np.array([endo_atmospheric_model(h) for h in alt])
    np.savez_compressed(os.path.join(OUT, "aero_probe.npz"), mach=mach, alpha=alpha, cl=cl, cd=cd,
                        ca=ca, cn=cn, alt=alt, atm=atm)
    print("aero_probe", n)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    load_reference()
    which = sys.argv[1:] or ["tape", "ss", "pso", "best", "rl", "wind", "classical", "aero", "other", "info"]
    if "sup" in which:
        supervisory_steps()
    if "supw" in which:
        supervisory_wrapper_sequences()
    if "info" in which:
        info_full()
        stored_info_csv()
    if "other" in which:
        ascent_csv()
        pool = rl_sequence_other(S_, "S", 400, mode="csv")
        single_step_other(S_, "S", pool)
        pool = rl_sequence_other(U_, "U", 600, mode="csv")
        single_step_other(U_, "U", pool)
        pool = rl_sequence_other(B_, "B", 500, mode="random")
        single_step_other(B_, "B", pool)
        pool = rl_sequence_other(C_, "C", 400, mode="random")
        pool2 = rl_sequence_other(C_, "C2", 2500, mode="track")
        single_step_other(C_, "C", np.concatenate([pool, pool2]))
    if "flip" in which:
        flip_over()
    if "psorun" in which:
        pso_run()
    if "aero" in which:
        aero_probe()
    if "tape" in which:
        tape_replay()
    if "ss" in which:
        single_step(P, "P")
        single_step(G, "G")
    if "pso" in which:
        pso_fitness(P, "P", n=10)
        pso_fitness(G, "G", n=16)
    if "many" in which:         # not in the default list: ~10 min of Pool(all cores)
        pso_fitness_many(P, "P")
        pso_fitness_many(G, "G")
    if "batch" in which:        # not in the default list: ~5 min of Pool(all cores)
        batch_tape(P, "P")
        batch_tape(G, "G")
    if "best" in which:
        pso_best_actor()
    if "rl" in which:
        rl_sequence(P, "P", 220)
        rl_sequence(G, "G", 12)
    if "wind" in which:
        wind_sequence()
    if "windU" in which or "wind" in which:
        wind_sequence_ascent()
    if "classical" in which:
        classical()
