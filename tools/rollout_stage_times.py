"""Per-kernel device times of one pd_rollout_pso call, unserialised (torch.profiler / CUPTI):
python tools/rollout_stage_times.py [particles] [phase] [seeds] [wind]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from psso_sac_for_powered_descent_b200 import envs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
phase = sys.argv[2] if len(sys.argv) > 2 else "landing_burn_pure_throttle"
seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 1
wind = len(sys.argv) > 4 and sys.argv[4] == "wind"
m = envs.pso_wrapped_env(flight_phase=phase, enable_wind=wind, stochastic_wind=wind, precision="fp32", max_steps=4096)
P = m.actor.number_of_network_parameters
pos = torch.as_tensor(np.random.default_rng(7).uniform(-1.5, 1.5, (n, P)).astype(np.float32)).cuda()
for rep in range(2):
    m._b.rollout_pso(pos, n_seeds=seeds, max_steps=4096)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fit, steps, tid = m._b.rollout_pso(pos, n_seeds=seeds, max_steps=4096)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type.name == "CUDA"], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    if e.time_range.elapsed_us() > 20:
        print(f"{(e.time_range.start - t0) / 1e3:9.3f} ms  +{e.time_range.elapsed_us() / 1e3:8.3f} ms  {e.name[:110]}")
print("total", (ev[-1].time_range.end - t0) / 1e3, "ms")
