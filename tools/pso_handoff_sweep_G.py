"""Rollout time of a random G-phase swarm (x 8 wind seeds) against the hand-off thresholds (one GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs, _native as N
m = envs.pso_wrapped_env(flight_phase="landing_burn", enable_wind=True, stochastic_wind=True, precision="fp32", max_steps=4096)
allpos = torch.as_tensor(np.random.default_rng(7).uniform(-1.5, 1.5, (65536, 372)).astype(np.float32)).cuda()
for n in (512, 8192, 65536):
    pos = allpos[:n].contiguous()
    for h1, h2 in ((0, 0), (128, 512), (8, 16), (12, 24), (16, 32), (24, 48), (16, 4096)):
        N.check(m._b.lib.pd_set_rollout_stages(m._b._h, h1, h2))
        best = 1e9
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            fit, steps, tid = m._b.rollout_pso(pos, n_seeds=8, max_steps=4096)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            if rep:
                best = min(best, dt)
        print(f"n {n} x 8 seeds handoff {h1}/{h2}: {best*1e3:.2f} ms, {n/best:.3e} evals/s, max steps {int(steps.max())}", flush=True)
