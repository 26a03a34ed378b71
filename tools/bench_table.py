"""Table of the committed bench lines (profiles/r2_bench_{N}gpu.json.log or the paths given)."""
import json, sys, os
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
paths = sys.argv[1:] or [os.path.join(root, f"r2_bench_{n}gpu.json.log") for n in (1, 2, 4, 8)]
print("| GPUs | env-steps/s | µs/step | e2e env-steps/s | config3 P ms | config3 G ms | 65 536 P ms | config5 G ms/gen | config5 P ms/gen | device swarm P ms/gen | SAC collect steps/s |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for path in paths:
    if not os.path.exists(path):
        continue
    d = json.loads([x for x in open(path) if x.startswith("{")][-1])
    p = d["pso"]
    print(f"| {d['n_gpus']} | {d['value']:.3e} | {d['ms_per_step'] * 1e3:.1f} | {d['e2e']['value']:.3e} | {p['config3_P']['ms']:.2f} | "
          f"{p['config3_G']['ms']:.2f} | {p['swarm_65536_P']['ms']:.1f} | {p['config5_G']['ms_per_generation']:.2f} | "
          f"{p['config5_P']['ms_per_generation']:.1f} | {p['device_swarm_65536_P_nowind']['ms_per_generation']:.1f} | "
          f"{(d.get('sac_collect') or {}).get('env_steps_per_s', float('nan')):.3e} |")
