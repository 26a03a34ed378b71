"""Stage-by-stage times and survivor counts of one PSO rollout (PD_ROLLOUT_TRACE), for a random swarm
and for the swarm after `generations` of the optimiser:  python tools/rollout_trace.py [particles seeds generations wind]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs, pso as pso_mod, _native as N

P = "landing_burn_pure_throttle"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
gens = int(sys.argv[3]) if len(sys.argv) > 3 else 30
wind = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
model = envs.pso_wrapped_env(flight_phase=P, enable_wind=wind, stochastic_wind=wind, max_steps=4096, seed=99, precision=os.environ.get("PD_PRECISION", "fp32"))
sw = pso_mod.DeviceSwarm(model, n, dict(pso_mod.PSO_PARAMS[P], pop_size=n), n_seeds=seeds, seed=5, max_steps=4096)
w0 = sw.weights.clone()
for g in range(gens):
    sw.step()
b = model._b
for name, w in (("initial", w0), ("evolved", sw.weights.clone())):
    for stages, lanes in (((128, 256), (0, 0)), ((128, 512), (66304, 8288))):
        N.check(b.lib.pd_set_rollout_stages(b._h, *stages))
        N.check(b.lib.pd_set_rollout_lanes(b._h, *lanes))
        b.rollout_pso(w, n_seeds=seeds, max_steps=4096, index0=0, generation=gens)
        torch.cuda.synchronize()
        print(f"== {name} swarm, {n} particles x {seeds} seeds, wind {wind}, stages {stages} lanes {lanes}", file=sys.stderr, flush=True)
        os.environ["PD_ROLLOUT_TRACE"] = "1"
        b.rollout_pso(w, n_seeds=seeds, max_steps=4096, index0=0, generation=gens)
        torch.cuda.synchronize()
        del os.environ["PD_ROLLOUT_TRACE"]
