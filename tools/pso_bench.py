"""PSO rollout micro-benchmark (one GPU): python tools/pso_bench.py [particles] [phase] [seeds] [wind]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
phase = sys.argv[2] if len(sys.argv) > 2 else "landing_burn_pure_throttle"
seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 1
wind = len(sys.argv) > 4 and sys.argv[4] == "wind"
m = envs.pso_wrapped_env(flight_phase=phase, enable_wind=wind, stochastic_wind=wind, precision="fp32", max_steps=4096)
P = m.actor.number_of_network_parameters
pos = torch.as_tensor(np.random.default_rng(7).uniform(-1.5, 1.5, (n, P)).astype(np.float32)).cuda()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fit, steps, tid = m._b.rollout_pso(pos, n_seeds=seeds, max_steps=4096)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
m._b.check_status()
tot = float(steps.sum())
print(f"{phase} particles {n} seeds {seeds} wind {wind}: {dt*1e3:.1f} ms, {n/dt:.3e} evals/s, {tot/dt:.3e} env-steps/s, mean steps {tot/(n*seeds):.1f}, max {int(steps.max())}")
