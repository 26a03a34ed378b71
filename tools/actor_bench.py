"""Shared-actor kernel alone: us per 131 072-env forward (python tools/actor_bench.py [lib])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    os.environ["PD_LIB_PATH"] = os.path.abspath(sys.argv[1])
import torch, torch.nn as nn
from psso_sac_for_powered_descent_b200 import envs
for phase, O, A in (("landing_burn_pure_throttle", 2, 1), ("landing_burn", 5, 4)):
    torch.manual_seed(0)
    l1, l2, mean_l, lstd_l = nn.Linear(O, 256), nn.Linear(256, 256), nn.Linear(256, A), nn.Linear(256, A)
    actor = dict(w1=l1.weight, b1=l1.bias, w2=l2.weight, b2=l2.bias, wm=mean_l.weight, bm=mean_l.bias,
                 ws=lstd_l.weight, bs=lstd_l.bias)
    actor = {k: v.detach().cuda().contiguous() for k, v in actor.items()}
    actor["max_action"] = 1.0
    env = envs.BatchedRocketEnv(1024, "rl", phase, precision="fp32", auto_reset=True)
    obs = torch.rand(131072, O, device="cuda") * 2 - 1
    for _ in range(5):
        env.actor_forward(actor, obs, deterministic=False, seed=1)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(20):
            env.actor_forward(actor, obs, deterministic=False, seed=1)
        torch.cuda.synchronize()
    ts = [e.time_range.elapsed_us() for e in prof.events() if e.device_type.name == "CUDA" and "actor_tc_kernel" in e.name]
    ts.sort()
    print(f"{os.path.basename(os.environ.get('PD_LIB_PATH', 'default'))} {phase}: actor_tc_kernel median {ts[len(ts) // 2]:.1f} us "
          f"per 131072 envs ({len(ts)} launches)")
