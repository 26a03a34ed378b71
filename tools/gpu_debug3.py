import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs
np.set_printoptions(linewidth=200, precision=9)
for tag, phase in (("P","landing_burn_pure_throttle"),("G","landing_burn")):
    g = np.load(f"tests/golden/pso_fitness_{tag}.npz")
    m = envs.pso_wrapped_env(flight_phase=phase, precision="fp64")
    w = torch.as_tensor(g["positions"].astype(np.float32)).cuda()
    out = m._b.rollout_pso(w, max_steps=512, trace=True)
    A = out["actions"].cpu().numpy(); steps = out["steps"].cpu().numpy(); fit = out["fitness"].cpu().numpy()
    for i in range(len(fit)):
        n = min(steps[i], g["steps"][i])
        ref = g["actions"][i][:n]; mine = A[:n, i]
        d = np.abs(mine-ref).max(axis=1)
        first = np.argmax(d > 2e-6) if (d > 2e-6).any() else -1
        print(tag, i, "steps", steps[i], g["steps"][i], "fit relerr", abs(fit[i]-g["fitness"][i])/abs(g["fitness"][i]), "cond", g["well_conditioned"][i],
              "max act diff", d.max(), "first>2e-6 at", first, "act0 diff", d[0], mine[0], ref[0])
