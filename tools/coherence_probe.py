"""Step time against de-synchronisation of the batch: python tools/coherence_probe.py
Prints the mean step time over windows of 100 steps for 3000 steps from a common reset
(episodes re-start at different times, so the batch loses coherence), then the same after
shuffling the env states across slots (worst case) and after sorting them by episode step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psso_sac_for_powered_descent_b200 import envs

B = 65536
env = envs.BatchedRocketEnv(B, "pso", "landing_burn_pure_throttle", precision="fp32", auto_reset=True)
g = torch.Generator(device="cuda").manual_seed(0)


def run(n, label):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for w in range(n // 100):
        tape = torch.rand(100, B, 1, device="cuda", generator=g) * 2 - 1
        e0.record()
        for k in range(100):
            env.step(tape[k])
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 10)
    print(label, " ".join(f"{v:.1f}" for v in out), "us/step per 100-step window")


def permute(order):
    full = env.get_state(full=True)
    env.set_state(*[t[order].contiguous() if t is not None else None for t in full])


env.reset()
run(3000, "from reset:")
try:
    st = env.get_state(full=True)
    t = st[0][:, 10]
    print("episode time spread:", float(t.min()), float(t.max()))
    order = torch.argsort(t)
    env.set_state(*[x[order].contiguous() for x in st])
    run(300, "sorted by time:")
    st = env.get_state(full=True)
    order = torch.randperm(B, device=st[0].device)
    env.set_state(*[x[order].contiguous() for x in st])
    run(300, "shuffled:")
except Exception as ex:
    print("permute failed:", repr(ex))
