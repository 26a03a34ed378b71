"""One process, handles on two GPUs: every ABI entry point selects its handle's device (and restores
the caller's), the shared-memory opt-in is remembered per device.  python tools/two_device_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs
assert torch.cuda.device_count() >= 2
P = "landing_burn_pure_throttle"
rng = np.random.default_rng(0)
acts = torch.as_tensor(rng.uniform(-1, 1, (20, 4096, 1)).astype(np.float32))
pos = torch.as_tensor(rng.uniform(-1.5, 1.5, (512, 249)).astype(np.float32))
res = []
e = [envs.BatchedRocketEnv(4096, "pso", P, precision="fp32", auto_reset=True, device=d) for d in (0, 1)]
m = [envs.pso_wrapped_env(flight_phase=P, precision="fp32") for _ in range(2)]
m[1]._b = envs.BatchedRocketEnv(1, "pso", P, precision="fp32", device=1)
torch.cuda.set_device(0)                     # the caller's current device stays 0 throughout
for k in range(20):
    for d in (0, 1):
        with torch.cuda.device(d):
            e[d].step(acts[k].to(f"cuda:{d}"))
    assert torch.cuda.current_device() == 0
s0, s1 = e[0].get_state().cpu(), e[1].get_state().cpu()
f0 = m[0]._b.rollout_pso(pos.cuda(0))[0].cpu()
with torch.cuda.device(1):
    f1 = m[1]._b.rollout_pso(pos.cuda(1))[0].cpu()
# calling a device-1 handle while device 0 is current: the guard switches and restores
e[1].check_status()
st = e[1].get_state()
assert st.device.index == 1 and torch.cuda.current_device() == 0
print("states identical on both devices:", torch.equal(s0, s1), "| fitness identical:", torch.equal(f0, f1))
assert torch.equal(s0, s1) and torch.equal(f0, f1)
