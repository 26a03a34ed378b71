"""fp32 build with the aero patches against the exact sums: build statistics, single-step difference
on a random batch, step time.  python tools/patch_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs
P = "landing_burn_pure_throttle"
B = 65536
t0 = time.perf_counter()
a = envs.BatchedRocketEnv(B, "pso", P, precision="fp32", auto_reset=True)
torch.cuda.synchronize()
print(f"create with patches: {time.perf_counter() - t0:.2f} s", a.aero_patch_stats(), flush=True)
t0 = time.perf_counter()
a2 = envs.BatchedRocketEnv(B, "pso", P, precision="fp32", auto_reset=True)
print(f"second handle: {time.perf_counter() - t0:.2f} s")
b = envs.BatchedRocketEnv(B, "pso", P, precision="fp32", auto_reset=True, exact_aero=True)
gen = torch.Generator(device="cuda").manual_seed(0)
worst = 0.0
for t in range(300):
    act = torch.rand(B, 1, device="cuda", generator=gen) * 2 - 1
    # same state in both before every step: the difference of ONE step
    st, gw, nw, ap = b.get_state(full=True) if hasattr(b, "get_state_full") else (None, None, None, None)
    oa = a.step(act); ob = b.step(act)
    sa, sb = a.get_state(), b.get_state()
    d = ((sa - sb).abs() / sb.abs().clamp_min(1.0)).max().item()
    worst = max(worst, d)
    flags_same = all(torch.equal(x, y) for x, y in zip(oa[2:], ob[2:]))
    if t % 100 == 0 or not flags_same:
        print(f"step {t}: max rel state diff (trajectories, not resynchronised) {d:.2e}, flags equal {flags_same}", flush=True)
a.check_status(); b.check_status()
for name, env in (("patches", a), ("exact", b)):
    act = torch.rand(B, 1, device="cuda", generator=gen) * 2 - 1
    for _ in range(20): env.step(act)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): env.step(act)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per step of {B} envs", flush=True)
