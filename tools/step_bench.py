"""Quick device-timed step-kernel benchmark (steady state): python tools/step_bench.py [envs] [precision] [lib]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 3:
    from psso_sac_for_powered_descent_b200 import _native
    _native.LIB_PATH = os.path.abspath(sys.argv[3])
import torch
from psso_sac_for_powered_descent_b200 import envs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
env = envs.BatchedRocketEnv(B, "pso", "landing_burn_pure_throttle", precision=prec, auto_reset=True)
K = 400
tape = torch.rand(K, B, 1, device="cuda") * 2 - 1
env.reset()
for k in range(200):
    env.step(tape[k])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(200, K):
    env.step(tape[k])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
env.check_status()
print(f"{os.path.basename(_native.LIB_PATH) if len(sys.argv)>3 else 'default'} envs {B} {prec}: {ms*1e3:.1f} us/step, {B/ms*1e3:.3e} env-steps/s")
