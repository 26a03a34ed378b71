"""SAC collection smoke / timing run: shared tensor-core actor + fused env step, a few steps, prints rewards."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from psso_sac_for_powered_descent_b200 import envs
P = "landing_burn_pure_throttle"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
torch.manual_seed(0)
l1, l2, m, s = nn.Linear(2, 256), nn.Linear(256, 256), nn.Linear(256, 1), nn.Linear(256, 1)
actor = dict(w1=l1.weight, b1=l1.bias, w2=l2.weight, b2=l2.bias, wm=m.weight, bm=m.bias, ws=s.weight, bs=s.bias, max_action=1.0)
def timeit(name, fn, n=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"{name}: {dt*1e3:.3f} ms")
for wind in (False, True):
    env = envs.BatchedRocketEnv(B, "rl", P, enable_wind=wind, stochastic_wind=wind, precision="fp32", auto_reset=True)
    obs = torch.rand(B, 2, device="cuda") * 2 - 1
    a = torch.rand(B, 1, device="cuda") * 2 - 1
    timeit(f"wind={wind} step only", lambda: env.step(a))
    timeit(f"wind={wind} actor_forward tc", lambda: env.actor_forward(actor, obs, deterministic=False))
    timeit(f"wind={wind} actor_forward fp32", lambda: env.actor_forward(actor, obs, deterministic=False, fp32_path=True), n=3)
    timeit(f"wind={wind} collect 10", lambda: env.collect(actor, 10), n=5)
    timeit(f"wind={wind} step only again", lambda: env.step(a))

print("--- repeated collect(40) timings with other handles alive")
import numpy as np
env0 = envs.BatchedRocketEnv(65536, "pso", P, precision="fp32", auto_reset=True)
tape = torch.rand(1010, 65536, 1, device="cuda") * 2 - 1
for k in range(50): env0.step(tape[k])
model = envs.pso_wrapped_env(flight_phase=P, precision="fp32")
pos = torch.as_tensor(np.random.default_rng(7).uniform(-1.5, 1.5, (4096, 249)).astype(np.float32)).cuda()
model._b.rollout_pso(pos)
senv = envs.BatchedRocketEnv(B, "rl", P, enable_wind=True, stochastic_wind=True, precision="fp32", auto_reset=True)
senv.collect(actor, 3, seed=1)
for rep in range(6):
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s0.record(); out = senv.collect(actor, 40, seed=2); s1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"rep {rep}: device {s0.elapsed_time(s1):.2f} ms, host enqueue {1e3*(t1-t0):.2f} ms")
    del out
