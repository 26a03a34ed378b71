"""Step-kernel time against the batch size (fp32 build, P, graph-free event timing over 200 launches):
one warp per SM shows the dependent-chain latency of a step, the full batch what contention adds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psso_sac_for_powered_descent_b200 import envs
for exact in (False, True):
    for B in (32 * 148, 4 * 32 * 148, 8 * 32 * 148, 65536, 131072, 262144):
        env = envs.BatchedRocketEnv(B, "pso", "landing_burn_pure_throttle", precision="fp32", auto_reset=True, exact_aero=exact)
        gen = torch.Generator(device="cuda").manual_seed(0)
        acts = torch.rand(64, B, 1, device="cuda", generator=gen) * 2 - 1
        for k in range(150): env.step(acts[k % 64])       # spread the episodes
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            env.step(acts[0]); s.synchronize()
            with torch.cuda.graph(g, stream=s):
                for k in range(100): env.step(acts[k % 64])
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 200 * 1e3
        print(f"exact_aero {exact} B {B}: {us:.1f} us/step, {B / us * 1e6:.3e} env-steps/s", flush=True)
        del env, g
