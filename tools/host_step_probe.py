"""Times BatchedRocketEnv.step_host (the e2e leg of bench.py) for the PD_HOST_STEP variants."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from psso_sac_for_powered_descent_b200 import envs

B, K = 65536, 300
tape = (torch.rand(K + 8, B, 1) * 2 - 1).pin_memory()
for mode in sys.argv[1:] or ["copy", "zc_out", "zc_all"]:
    os.environ["PD_HOST_STEP"] = mode
    env = envs.BatchedRocketEnv(B, "pso", "landing_burn_pure_throttle", precision="fp32", auto_reset=True)
    best = 1e9
    for rep in range(3):
        env.reset()
        for w in range(4):
            env.step_host(tape[w])
        torch.cuda.synchronize()
        t0 = time.perf_counter(); chk = 0.0
        for k in range(K):
            obs, rew, done, trunc, tid = env.step_host(tape[4 + k])
            chk += float(rew[0])
        best = min(best, time.perf_counter() - t0)
    print(f"{mode}: {best / K * 1e6:.1f} us/step, {B * K / best:.3e} env-steps/s, chk {chk:.6f}")
    del env
