"""GPU diagnostic: where do the fp32 production build's termination flags differ from the
reference's (single-step fixtures) and from the fp64 build's (config-2 tape)?

    python tools/fp32_flag_probe.py [n_envs n_steps]

Prints every mismatching fixture row with the distance of each thresholded quantity from its
threshold (parity.threshold_margins), then the lock-step agreement figures of parity.fp32_vs_fp64_tape.
"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import pd_oracle as O  # noqa: E402  (diagnostic tool: the oracle's ISA only)
from psso_sac_for_powered_descent_b200 import envs, parity  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")
P, G = "landing_burn_pure_throttle", "landing_burn"


def fixture_rows(tag, phase):
    g = np.load(os.path.join(GOLD, f"single_step_{tag}.npz"), allow_pickle=True)
    n = len(g["state"])
    env = envs.BatchedRocketEnv(n, "pso", phase, precision="fp32")
    env.set_state(g["state"], g["win"], g["nwin"].astype(np.int32), g["aprev"])
    obs, rew, done, trunc, tid = env.step(torch.as_tensor(g["act32"]).cuda())
    ref = g["o32"]
    d, t, i = done.cpu().numpy(), trunc.cpu().numpy(), tid.cpu().numpy()
    bad = np.nonzero(~((d == ref[:, 12]) & (t == ref[:, 13]) & (i == ref[:, 14])))[0]
    cols = list(g["out_cols"])
    jg = cols.index("g1")
    print(f"== single_step_{tag}: {len(bad)} of {n} rows differ")
    st = env.get_state().cpu().numpy()
    for k in bad:
        rho = O.isa(ref[k, 1])[0]
        m = parity.threshold_margins(phase, "pso", ref[k, :11], ref[k, jg], rho)
        key = min(m, key=m.get)
        print(f" row {k}: ref flags {ref[k, 12:15]} fp32 {(d[k], t[k], i[k])}  nearest threshold {key} "
              f"margin {m[key]:.3e}  y {ref[k, 1]:.6g} vy {ref[k, 3]:.6g} g1 {ref[k, jg]:.9g} "
              f"fp32 y {st[k, 1]:.6g}")


if __name__ == "__main__":
    n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    fixture_rows("P", P)
    fixture_rows("G", G)
    for phase, T in ((P, n_steps), (G, min(n_steps, 200))):
        for test in ("fp32", "ulp"):
            r = parity.fp32_vs_fp64_tape(n_envs, T, phase=phase, test=test)
            recs = r.pop("first_mismatches")
            print(json.dumps(r))
            for rec in recs[:12]:
                rec.pop("state_before_fp64", None)
                print("   ", json.dumps(rec))
