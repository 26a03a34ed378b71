"""Agreement of the fp32 production build with the fp64 parity build on one action tape.

The fp64 build is the one pinned to the reference (single steps 1e-12, tests/test_gpu_parity.py);
every throughput number is measured on the fp32 build.  north_star asks for termination flags
that match exactly, so this module measures - on BASELINE config 2's own workload, 65 536 envs x
1 000 steps of random float32 actions with auto-reset - how often they do:

    flag_match_frac          env-steps whose (done, truncated, truncation id) are identical
    episode_same_length_frac episodes (as the fp64 build sees them) that end at the same step with
                             the same flags in the fp32 build, with no disagreement before
    max_state_err            largest relative state difference (absolute floor per component)
                             over all env-steps of episodes that are still in agreement

Both handles run the same tape in lock step.  An env whose flags disagree is counted and then
re-synchronised (the fp64 handle's full state is copied into the fp32 handle), so one early
disagreement is not counted again on every later step and the two batches keep facing the same
episodes.  `bench.py` prints the result as its `parity` block, tests/test_gpu_parity.py asserts it.
"""
from __future__ import annotations

import numpy as np
import torch

from .envs import BatchedRocketEnv

# per-component absolute floors of the relative state error (tests/test_gpu_parity.py FLOOR)
FLOOR = {
    "landing_burn_pure_throttle": [1e3, 1e3, 1e2, 1e2, 1.0, 2.5, 1.0, 1.0, 1e5, 1e5, 1e2],
    "landing_burn": [1e3, 1e3, 1e2, 1e2, 2.0, 10.0, 1.0, 2.0, 1e5, 1e5, 1e2],
}


_ISA = ((-5.0e3, 320.65, -6.5e-3, 1.77687e5), (0.0e3, 288.15, -6.5e-3, 1.01325e5),
        (11.0e3, 216.65, 0.0, 2.26320e4), (20.0e3, 216.65, 1.0e-3, 5.47487e3),
        (32.0e3, 228.65, 2.8e-3, 8.68014e2), (47.0e3, 270.65, 0.0, 1.10906e2),
        (51.0e3, 270.65, -2.8e-3, 6.69384e1), (71.0e3, 214.65, -2.0e-3, 3.95639e0))


def isa_density(alt):
    """ICAO-1993 ISA density [kg/m^3] (the layer table of csrc/pd_api.cu), for the margin report."""
    alt = max(float(alt), 0.0)
    if not alt < 81020:
        return 0.0
    H = 6356766.0 * alt / (6356766.0 + alt)
    Hb, Tb, beta, pb = [l for l in _ISA if H >= l[0]][-1]
    T = Tb + beta * (H - Hb)
    if beta == 0.0:
        p = pb * np.exp(-9.80665 / (287.05287 * T) * (H - Hb))
    else:
        p = pb * (1.0 + (beta / Tb) * (H - Hb)) ** (-9.80665 / (beta * 287.05287))
    return float(p / (287.05287 * T))


def fp32_vs_fp64_tape(n_envs=65536, n_steps=1000, phase="landing_burn_pure_throttle", rtd="pso",
                      seed=0, device=None, max_records=32, action_tape=None, test="fp32"):
    """Returns a dict of agreement figures (see module docstring).  `action_tape`: optional cuda
    float32 tensor [n_steps, n_envs, A]; default = U(-1, 1) drawn step by step from `seed`.
    test = 'fp32': the production build against the fp64 build.  test = 'ulp': the CONDITIONING
    BASELINE - the fp64 build against itself with every action moved to the next float32 (what a
    different torch / BLAS rounding of the policy output does to the reference itself): whatever
    disagreement this shows is a property of the dynamics, not of a build."""
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    assert test in ("fp32", "ulp")
    e32 = BatchedRocketEnv(n_envs, rtd, phase, precision="fp32" if test == "fp32" else "fp64",
                           auto_reset=True, device=dev.index)
    e64 = BatchedRocketEnv(n_envs, rtd, phase, precision="fp64", auto_reset=True, device=dev.index)
    e32.reset(); e64.reset()
    A = e32.act_dim
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    floor = torch.tensor(FLOOR[phase], dtype=torch.float64, device=dev)
    clean = torch.ones(n_envs, dtype=torch.bool, device=dev)      # no disagreement yet in this episode
    n_flag_match = torch.zeros((), dtype=torch.int64, device=dev)
    n_episodes = torch.zeros((), dtype=torch.int64, device=dev)
    n_same_len = torch.zeros((), dtype=torch.int64, device=dev)
    n_resync = 0
    max_err = torch.zeros((), dtype=torch.float64, device=dev)
    q999_sum = 0.0
    err_by_comp = torch.zeros(11, dtype=torch.float64, device=dev)
    records = []
    ep_step = torch.zeros(n_envs, dtype=torch.int32, device=dev)
    for t in range(n_steps):
        a = action_tape[t] if action_tape is not None else \
            torch.rand(n_envs, A, device=dev, generator=gen, dtype=torch.float32) * 2 - 1
        # pre-step state of the reference build, kept for the mismatch records
        pre64 = e64.get_state() if len(records) < max_records else None
        a_test = a if test == "fp32" else torch.nextafter(a, torch.full_like(a, 2.0))
        _, _, d32, t32, i32 = e32.step(a_test)
        dbg64 = torch.zeros(n_envs, 16, dtype=torch.float64, device=dev) if len(records) < max_records else None
        _, _, d64, t64, i64 = e64.step(a, dbg=dbg64)
        same = (d32 == d64) & (t32 == t64) & (i32 == i64)
        n_flag_match += same.sum()
        ended64 = (d64 | t64).bool()
        ended32 = (d32 | t32).bool()
        n_episodes += ended64.sum()
        n_same_len += (ended64 & same & clean).sum()
        ep_step += 1
        # state agreement of the envs that are still inside a clean episode in both builds
        live = clean & same & ~ended64
        s32, s64 = e32.get_state(), e64.get_state()
        if t == 0 or pre64 is not None:
            post64 = s64
        rel = (s32 - s64).abs() / torch.maximum(s64.abs(), floor)
        rel = torch.where(live[:, None], rel, torch.zeros_like(rel))
        max_err = torch.maximum(max_err, rel.max())
        err_by_comp = torch.maximum(err_by_comp, rel.max(dim=0).values)
        q999_sum += float(torch.quantile(rel.max(dim=1).values, 0.999))
        bad = ~same
        nbad = int(bad.sum())
        if nbad:
            idx = torch.nonzero(bad).flatten()
            for j in idx[:max(0, max_records - len(records))].tolist():
                records.append(dict(step=t, env=j, episode_step=int(ep_step[j]),
                                    fp64=(int(d64[j]), int(t64[j]), int(i64[j])),
                                    fp32=(int(d32[j]), int(t32[j]), int(i32[j])),
                                    action=a[j].tolist(),
                                    state_before_fp64=pre64[j].tolist() if pre64 is not None else None))
                if pre64 is not None and dbg64 is not None and phase in FLOOR:
                    # margins from the pre-reset post-step state: only available for the build that did
                    # NOT reset; the fp64 state after a reset is the initial state, so use whichever
                    # build is still flying, and where both ended, the fp64 pre-state is reported only
                    flying = s32[j] if (d64[j] | t64[j]) else s64[j]
                    if not ((d64[j] | t64[j]) and (d32[j] | t32[j])):
                        st = flying.tolist()
                        m = threshold_margins(phase, rtd, st, float(dbg64[j, 12]), isa_density(st[1]))
                        k = min(m, key=m.get)
                        records[-1].update(nearest_threshold=k, margin=m[k])
            # re-synchronise: the fp32 handle continues from the fp64 handle's state
            f64 = e64.get_state(full=True)
            f32 = e32.get_state(full=True)
            merged = [torch.where(bad.reshape([-1] + [1] * (x64.dim() - 1)), x64, x32)
                      for x64, x32 in zip(f64, f32)]
            e32.set_state(*merged)
            n_resync += nbad
        # a new episode (either build ended) starts clean once the two handles are in sync again
        restart = ended64 | ended32 | bad
        clean = torch.where(ended64 | ended32, torch.ones_like(clean), clean & same)
        ep_step = torch.where(restart, torch.zeros_like(ep_step), ep_step)
    e32.check_status(); e64.check_status()
    total = n_envs * n_steps
    n_ep = int(n_episodes)
    return dict(phase=phase, rtd=rtd, test=test, n_envs=n_envs, n_steps=n_steps, env_steps=total,
                flag_match_frac=float(n_flag_match) / total,
                flag_mismatches=total - int(n_flag_match),
                episodes=n_ep,
                episode_same_length_frac=(int(n_same_len) / n_ep) if n_ep else None,
                max_state_err=float(max_err),
                max_state_err_by_component=[float(v) for v in err_by_comp.tolist()],
                max_translational_err=float(err_by_comp[:4].max()),
                mean_p999_state_err=q999_sum / n_steps,
                resynced_envs=n_resync,
                first_mismatches=records,
                what=("fp32 production build vs fp64 parity build (oracle-pinned), same float32 action tape"
                      if test == "fp32" else
                      "conditioning baseline: fp64 build vs fp64 build with every action moved by one float32 ulp")
                + ", lock step with auto-reset; state error = |d| / max(|fp64|, floor) over envs whose "
                  "episode is still in agreement (components x y vx vy theta theta_dot gamma alpha m mp t)")


def threshold_margins(phase, rtd, state, g1, rho):
    """Distance of every thresholded quantity of the pso / rl closures (rtd_pso.py:172-317,
    rtd_rl.py:194-240) from its threshold, relative to the quantity's scale, for one post-step state
    (numpy, float64).  `g1` = 1-s g-load window mean, `rho` = ISA density at the state's altitude.
    Used to show that a flag disagreement between two builds sits on a threshold."""
    x, y, vx, vy, theta, theta_dot, gamma, alpha, mass, m_prop, t = [float(v) for v in state]
    speed = float(np.hypot(vx, vy))
    q = 0.5 * rho * speed * speed
    m = {}
    fl_y, fl_v = 1e3, 1e2                   # the state floors: y is known to 1e-5 * max(|y|, 1e3)
    if rtd == "pso" and phase == "landing_burn_pure_throttle":
        m["y<0"] = abs(y) / fl_y
        m["y<1"] = abs(y - 1.0) / fl_y
        m["speed<5.5"] = abs(speed - 5.5) / fl_v
    elif rtd == "pso":
        dist = float(np.hypot(x, y))
        over = dist if (x < 0 and y < 0) else (-x if x < 0 else (-y if y < 0 else 0.0))
        m["overshoot>0.5"] = abs(over - 0.5) / fl_y
        m["dist<1"] = abs(dist - 1.0) / fl_y
        m["speed<2.5"] = abs(speed - 2.5) / fl_v
        a_eff = abs(gamma - theta - np.pi) if vy < 0 else abs(theta - gamma)
        m["alpha_eff>10deg"] = abs(a_eff - np.radians(10)) / 1.0
        m["y>1000"] = abs(y - 1000.0) / fl_y
        m["vx>0"] = abs(vx) / fl_v
    else:
        m["y<-10"] = abs(y + 10.0) / fl_y
        m["y<0"] = abs(y) / fl_y
        m["y<1"] = abs(y - 1.0) / fl_y
        m["speed<5"] = abs(speed - 5.0) / fl_v
        m["vx>0.01"] = abs(vx - 0.01) / fl_v
    m["m_prop<=0"] = abs(m_prop) / 1e5
    if phase == "landing_burn_pure_throttle" or rtd != "pso":
        m["theta>pi+2deg"] = abs(theta - (np.pi + np.radians(2))) / 1.0
    m["q>65000"] = abs(q - 65000.0) / 65000.0
    m["vy>0"] = abs(vy) / fl_v
    m["g>6"] = abs(g1 - 6.0) / 6.0
    return m
