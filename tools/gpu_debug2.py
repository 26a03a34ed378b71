import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs
np.set_printoptions(linewidth=200, precision=6)
g = np.load("tests/golden/p_tape_replay.npz", allow_pickle=True)
env = envs.BatchedRocketEnv(1, "pso", "landing_burn_pure_throttle", precision="fp64")
env.reset()
dbg = torch.zeros(1,16,dtype=torch.float64,device="cuda")
cols = list(g["info_cols"])
first=None
for k in range(1281):
    obs, rew, done, trunc, tid = env.step(torch.tensor([[g["u0"][k]]],dtype=torch.float64,device="cuda"), dbg=dbg)
    d = dbg[0].cpu().numpy()
    st = env.get_state()[0].cpu().numpy()
    if d[15] != 0 and first is None:
        first = k
        print("first rbf status", d[15], "at step", k, "mach", d[0], "alpha_eff", d[11], "ref mach", g["info"][k][0], "ref alpha_eff", g["info"][k][11])
        print("  CL", d[2], g["info"][k][2], "CD", d[3], g["info"][k][3])
        print("  state", st); print("  ref  ", g["states"][k])
    if k % 160 == 0 or k > 1275:
        e = np.abs(st-g["states"][k])/np.maximum(np.abs(g["states"][k]),1e-3)
        print(k, "max rel err", e.max(), "CL", d[2], g["info"][k][2], "CD", d[3], g["info"][k][3], "status", d[15], bool(done[0]), bool(trunc[0]))
