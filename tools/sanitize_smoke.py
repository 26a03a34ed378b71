"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
from psso_sac_for_powered_descent_b200 import envs, pso
P, G = "landing_burn_pure_throttle", "landing_burn"
rng = np.random.default_rng(0)
for phase, A in ((P, 1), (G, 4)):
    for prec in ("fp32", "fp64"):
        for rtd in ("pso", "rl"):
            env = envs.BatchedRocketEnv(300, rtd, phase, enable_wind=True, stochastic_wind=True, precision=prec, auto_reset=True)
            for t in range(3):
                env.step(torch.as_tensor(rng.uniform(-1, 1, (300, A)).astype(np.float32)).cuda())
            env.check_status()
    m = envs.pso_wrapped_env(flight_phase=phase, precision="fp32", max_steps=64)
    n = m.actor.number_of_network_parameters
    m.evaluate(rng.uniform(-1.5, 1.5, (100, n)))                 # cooperative rollout
    m._b.rollout_pso(torch.as_tensor(rng.uniform(-1.5, 1.5, (40000, n)).astype(np.float32)).cuda(), max_steps=8)
    m._b.check_status()
env = envs.BatchedRocketEnv(1, "pso", P, precision="fp64")
env.rollout_classical(1, max_steps=64)
env.rollout_tape(torch.zeros(16, 4, 1, dtype=torch.float64).cuda(), record=True)
torch.manual_seed(0)
l1, l2, mm, ss = nn.Linear(2, 256), nn.Linear(256, 256), nn.Linear(256, 1), nn.Linear(256, 1)
actor = dict(w1=l1.weight, b1=l1.bias, w2=l2.weight, b2=l2.bias, wm=mm.weight, bm=mm.bias, ws=ss.weight, bs=ss.bias)
senv = envs.BatchedRocketEnv(1000, "rl", P, precision="fp32", auto_reset=True)
senv.collect(actor, 3)
senv.actor_forward(actor, torch.rand(77, 2).cuda(), fp32_path=True)
sw = pso.DeviceSwarm(envs.pso_wrapped_env(flight_phase=G, precision="fp32", max_steps=32), 256,
                     dict(pso.landing_burn_pso_params, pop_size=256), seed=1, max_steps=32)
sw.step(); sw.step()
torch.cuda.synchronize()
print("sanitize smoke done")
