"""Episode-length distribution of a random P-phase swarm and rollout time per swarm size (one GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs
m = envs.pso_wrapped_env(flight_phase="landing_burn_pure_throttle", precision="fp32", max_steps=4096)
allpos = torch.as_tensor(np.random.default_rng(7).uniform(-1.5, 1.5, (65536, 249)).astype(np.float32)).cuda()
for n in (64, 4096, 16384, 65536):
    pos = allpos[:n].contiguous()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fit, steps, tid = m._b.rollout_pso(pos, max_steps=4096)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    s = steps.cpu().numpy()
    print(f"n {n}: {dt*1e3:.1f} ms, max steps {s.max()}, us per step of the longest episode {dt*1e6/s.max():.1f}, "
          f"quantiles 50/90/99/99.9 {np.percentile(s, [50, 90, 99, 99.9])}, n>512 {(s>512).sum()}, n>1024 {(s>1024).sum()}, "
          f"capped {(tid.cpu().numpy()<0).sum()}")
