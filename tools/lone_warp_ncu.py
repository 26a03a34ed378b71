"""One warp per SM through the fp32 step kernel (for ncu: the dependent chain of a step without contention)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psso_sac_for_powered_descent_b200 import envs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32 * 148
env = envs.BatchedRocketEnv(B, "pso", "landing_burn_pure_throttle", precision="fp32", auto_reset=True)
gen = torch.Generator(device="cuda").manual_seed(0)
acts = torch.rand(64, B, 1, device="cuda", generator=gen) * 2 - 1
for k in range(260): env.step(acts[k % 64])
torch.cuda.synchronize()
