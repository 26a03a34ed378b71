"""Small SAC-collection run for ncu launch lists: python tools/collect_probe.py [envs] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from psso_sac_for_powered_descent_b200 import envs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.manual_seed(0)
l1, l2, m, s = nn.Linear(2, 256), nn.Linear(256, 256), nn.Linear(256, 1), nn.Linear(256, 1)
actor = dict(w1=l1.weight, b1=l1.bias, w2=l2.weight, b2=l2.bias, wm=m.weight, bm=m.bias, ws=s.weight, bs=s.bias, max_action=1.0)
env = envs.BatchedRocketEnv(B, "rl", "landing_burn_pure_throttle", enable_wind=True, stochastic_wind=True,
                            precision="fp32", auto_reset=True, seed=3)
env.reset()
env.collect(actor, 2, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = env.collect(actor, T, seed=2); e1.record(); torch.cuda.synchronize()
env.check_status()
print(f"collect {T} steps x {B} envs: {e0.elapsed_time(e1)/T*1e3:.1f} us/step, mean reward {float(out['rewards'].mean()):.4f}")
