"""Wall time of every generation of the device-resident optimiser over 65 generations (one GPU): shows the
one-off cost of the first migration / sharing steps and their steady-state cost."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psso_sac_for_powered_descent_b200 import envs, pso as pso_mod
G = "landing_burn"
for phase, n, seeds in ((G, 8192, 8), ("landing_burn_pure_throttle", 8192, 1)):
    model = envs.pso_wrapped_env(flight_phase=phase, enable_wind=True, stochastic_wind=True, max_steps=4096, seed=99, precision=os.environ.get("PD_PRECISION", "fp32"))
    params = dict(pso_mod.PSO_PARAMS[phase], pop_size=n)
    sw = pso_mod.DeviceSwarm(model, n, params, n_seeds=seeds, seed=5, max_steps=4096)
    ts = []
    for g in range(65):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sw.step()
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(phase, " ".join(f"{g}:{t:.1f}" for g, t in enumerate(ts)), flush=True)
