"""Per-step latency of every rollout stage variant at a given number of identical long episodes:
the longest-lived particle of a random swarm is replicated k times and one traced rollout
(PD_ROLLOUT_TRACE) is run with the lane thresholds forced to 1, 8 and 32 lanes per episode.
    python tools/lone_episode_latency.py [wind]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs, _native as N

P = "landing_burn_pure_throttle"
wind = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
if len(sys.argv) > 2:
    P = sys.argv[2]
NP = 249 if P == "landing_burn_pure_throttle" else 372
STAGES = (128, 256) if NP == 249 else (8, 16)
model = envs.pso_wrapped_env(flight_phase=P, enable_wind=wind, stochastic_wind=wind, max_steps=4096, seed=99, precision=os.environ.get("PD_PRECISION", "fp32"))
b = model._b
w = torch.as_tensor(np.random.default_rng(7).uniform(-1.5, 1.5, (65536, NP)).astype(np.float32)).cuda()
fit, steps, tid = b.rollout_pso(w, n_seeds=1, max_steps=4096)
j = int(torch.argmax(steps))
print(f"longest particle {j}: {int(steps[j])} steps", file=sys.stderr, flush=True)
for k in (2, 148, 592, 2072, 8288):
    wk = w[j:j + 1].repeat(k, 1).contiguous()
    for lanes, name in (((66304, 66304), "32 lanes"), ((66304, 1), "8 lanes"), ((1, 1), "1 lane")):
        N.check(b.lib.pd_set_rollout_stages(b._h, *STAGES))
        N.check(b.lib.pd_set_rollout_lanes(b._h, *lanes))
        b.rollout_pso(wk, n_seeds=1, max_steps=4096)
        torch.cuda.synchronize()
        print(f"== {k} copies, record-fed stages with {name}", file=sys.stderr, flush=True)
        os.environ["PD_ROLLOUT_TRACE"] = "1"
        f2, s2, t2 = b.rollout_pso(wk, n_seeds=1, max_steps=4096)
        torch.cuda.synchronize()
        del os.environ["PD_ROLLOUT_TRACE"]
