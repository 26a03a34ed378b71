"""How long are the episodes of a swarm the optimiser has worked on for a while, and what does the
staged rollout make of them?  python tools/evolved_swarm_probe.py [particles seeds generations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs, pso as pso_mod, _native as N

P = "landing_burn_pure_throttle"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
gens = int(sys.argv[3]) if len(sys.argv) > 3 else 30
model = envs.pso_wrapped_env(flight_phase=P, enable_wind=True, stochastic_wind=True, max_steps=4096, seed=99, precision=os.environ.get("PD_PRECISION", "fp32"))
sw = pso_mod.DeviceSwarm(model, n, dict(pso_mod.PSO_PARAMS[P], pop_size=n), n_seeds=seeds, seed=5, max_steps=4096)
for g in range(gens):
    sw.step()
    if g in (0, 5, 10, 20, gens - 1):
        s = sw.last_steps.cpu().numpy()
        print(f"gen {g}: episodes {s.size} mean {s.mean():.0f} >128 {np.mean(s > 128):.3f} >512 {np.mean(s > 512):.3f} "
              f">1024 {np.mean(s > 1024):.3f} >2048 {np.mean(s > 2048):.3f} capped {int((s >= 4096).sum())}", flush=True)
b = model._b
rng_w = torch.as_tensor(np.random.default_rng(7).uniform(-1.5, 1.5, (n, 249)).astype(np.float32)).cuda()


def best_of(w, reps=3):
    best = 1e9
    for rep in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fit, steps, tid = b.rollout_pso(w, n_seeds=seeds, max_steps=4096, index0=0, generation=gens)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3, int(steps.max())


for name, w in (("random", rng_w), ("evolved", sw.weights.clone())):
    for stages in ((0, 0), (128, 256), (128, 512), (64, 128), (256, 512), (128, 4096)):
        N.check(b.lib.pd_set_rollout_stages(b._h, *stages))
        N.check(b.lib.pd_set_rollout_lanes(b._h, 0, 0))
        ms, longest = best_of(w)
        print(f"{name} stages {stages}: {ms:.1f} ms  (longest {longest} -> floor {longest * 7.6e-3:.1f} ms)", flush=True)
    N.check(b.lib.pd_set_rollout_stages(b._h, 128, 256))
    for lanes in ((66304, 8288), (33152, 2762), (16576, 4144), (16576, 2072), (16576, 1381), (8288, 2762), (24000, 2762)):
        N.check(b.lib.pd_set_rollout_lanes(b._h, *lanes))
        ms, longest = best_of(w)
        print(f"{name} stages (128, 256) lanes {lanes}: {ms:.1f} ms", flush=True)
