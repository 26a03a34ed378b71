"""Per-step divergence of a free-running RL-wrapper episode (fp64 build) from its reference fixture:
    python tools/sequence_divergence.py C landing_burn_pure_throttle_Pcontrol
Shows how fast a 1e-14 rounding difference grows through the pitch channel of each phase."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs
tag, phase = sys.argv[1], sys.argv[2]
g = np.load(f"tests/golden/rl_sequence_{tag}.npz")
env = envs.BatchedRocketEnv(1, "rl", phase, precision="fp64", trajectory_length=1000, discount_factor=0.99)
env.reset()
acts = torch.as_tensor(g["actions"]).cuda()
dbg = torch.zeros(1, 16, dtype=torch.float64, device="cuda")
for k in range(len(acts)):
    obs, rew, done, trunc, tid = env.step(acts[k].reshape(1, -1), dbg=dbg)
    st = env.get_state().cpu().numpy()[0]
    e = np.abs(st - g["states"][k]) / np.maximum(np.abs(g["states"][k]), 1e-3)
    if k < 6 or k % 20 == 0 or bool(trunc[0]) or bool(done[0]):
        print(k, "max rel err", f"{e.max():.2e}", "col", int(e.argmax()), "rew", float(rew[0]), g["rewards"][k],
              "flags", int(done[0]), int(trunc[0]), int(tid[0]), "ref", bool(g["done"][k]), bool(g["truncated"][k]),
              "g1", float(dbg[0, 12]), "q", float(dbg[0, 1]), "thr", float(dbg[0, 10]))
    if bool(trunc[0]) or bool(done[0]):
        break
