"""Wall time of pso_wrapped_env.evaluate (host list in, fitness out) per swarm size and phase (one GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs
for phase, P in (("landing_burn", 372), ("landing_burn_pure_throttle", 249)):
    m = envs.pso_wrapped_env(flight_phase=phase, precision="fp32", max_steps=4096)
    m.warn_on_cap = False
    for n in (4096, 65536):
        pos = np.random.default_rng(7).uniform(-1.5, 1.5, (n, P))
        for rep in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            w = np.ascontiguousarray(pos)
            t1 = time.perf_counter()
            out = m.evaluate(pos)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            wt = torch.as_tensor(pos.astype(np.float32)).cuda()
            torch.cuda.synchronize(); t3 = time.perf_counter()
            out2 = m._b.rollout_pso(wt, max_steps=4096)
            torch.cuda.synchronize(); t4 = time.perf_counter()
        print(f"{phase} n {n}: evaluate {1e3*(t2-t1):.2f} ms | numpy astype + upload {1e3*(t3-t2):.2f} ms, rollout only {1e3*(t4-t3):.2f} ms")
