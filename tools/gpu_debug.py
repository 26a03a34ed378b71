import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from psso_sac_for_powered_descent_b200 import envs
np.set_printoptions(linewidth=200, precision=3)
FLOOR = np.array([1e3, 1e3, 1e2, 1e2, 1.0, 1e-2, 1.0, 1.0, 1e5, 1e5, 1e2])
for tag, phase in (("P","landing_burn_pure_throttle"),("G","landing_burn")):
    g = np.load(f"tests/golden/single_step_{tag}.npz", allow_pickle=True)
    for key, akey in (("o64","act64"),("o32","act32")):
        n = len(g["state"])
        env = envs.BatchedRocketEnv(n, "pso", phase, precision="fp64")
        env.set_state(g["state"], g["win"], g["nwin"].astype(np.int32), g["aprev"])
        dbg = torch.zeros(n,16,dtype=torch.float64,device="cuda")
        obs, rew, done, trunc, tid = env.step(torch.as_tensor(g[akey]).cuda(), dbg=dbg)
        st, gw, nw, ap = env.get_state(full=True)
        ref = g[key]
        e = np.abs(st.cpu().numpy()-ref[:,:11])/np.maximum(np.abs(ref[:,:11]),FLOOR)
        print(tag,key,"state err per comp", e.max(0))
        i = e.max(1).argmax(); print("  worst row", i, "state", g["state"][i], "act", g[akey][i])
        d = dbg.cpu().numpy(); cols=list(g["out_cols"][18:])
        for j,name in enumerate(cols):
            jj = 12 if name=="g1" else j
            refv = ref[:,18+j]; er = np.abs(d[:,jj]-refv)/np.maximum(np.abs(refv),1e-3)
            print("   ",name, er.max(), "row", er.argmax(), d[er.argmax(),jj], refv[er.argmax()])
        print("  reward err", np.max(np.abs(rew.cpu().numpy()-ref[:,11])/np.maximum(np.abs(ref[:,11]),1.0)),
              "flags eq", np.array_equal(done.cpu().numpy(),ref[:,12]), np.array_equal(trunc.cpu().numpy(),ref[:,13]), np.array_equal(tid.cpu().numpy(),ref[:,14]))
        if tag=="G": print("  aprev err", np.max(np.abs(ap.cpu().numpy()-ref[:,15:18])))
        try: env.check_status()
        except Exception as ex: print("STATUS", ex)
