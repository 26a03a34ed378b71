"""One long hovering episode x 148 copies through the final 32-lane stage only (for ncu):
launch order of rollout_kernel: [0] 8 lanes reset->2048, [1] 1-lane final (no-op), [2] 8-lane final
(no-op), [3] 32-lane final 2048->4096."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs, _native as N
model = envs.pso_wrapped_env(flight_phase="landing_burn_pure_throttle", enable_wind=False, max_steps=4096, seed=99, precision=os.environ.get("PD_PRECISION", "fp32"))
b = model._b
w = torch.as_tensor(np.random.default_rng(7).uniform(-1.5, 1.5, (65536, 249)).astype(np.float32)).cuda()
wk = w[20791:20792].repeat(148, 1).contiguous()
N.check(b.lib.pd_set_rollout_stages(b._h, 2048, 4096))
N.check(b.lib.pd_set_rollout_lanes(b._h, 66304, 66304))
fit, steps, tid, term = b.rollout_pso(wk, n_seeds=1, max_steps=4096, terminal=True)
torch.cuda.synchronize()
print("steps", int(steps[0]), "terminal", term[0].cpu().numpy())
