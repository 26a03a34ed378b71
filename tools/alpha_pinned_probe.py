"""How often is the aerodynamic query pinned on the table edge?  C_L is looked up at
|alpha_eff| x (180/pi)^2 clamped to 10 and C_D at alpha_eff x 180/pi clamped to +-0.1745, so both
leave the edge lines only while |alpha_eff| < 0.003 rad.  Config-2 workload (random actions)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs
B = 65536
env = envs.BatchedRocketEnv(B, "pso", "landing_burn_pure_throttle", precision="fp32", auto_reset=True)
gen = torch.Generator(device="cuda").manual_seed(0)
tot = lanes = warps = 0
hist = np.zeros(6)
edges = [0.0, 0.003, 0.01, 0.03, 0.1, 0.3, 10.0]
for t in range(400):
    env.step(torch.rand(B, 1, device="cuda", generator=gen) * 2 - 1)
    s = env.get_state()
    a = torch.where(s[:, 3] < 0, s[:, 6] - s[:, 4] - math.pi, s[:, 7]).abs()
    free = a < 10.0 / (180 / math.pi) ** 2
    tot += B; lanes += int(free.sum()); warps += int(free.view(-1, 32).any(dim=1).sum())
    hist += np.histogram(a.cpu().numpy(), bins=edges)[0]
print(f"lanes off the edge lines: {lanes / tot:.4f}; warps with at least one such lane: {warps / (tot / 32):.4f}")
print("histogram of |alpha_eff| (rad) over bins", edges, ":", (hist / hist.sum()).round(4))
