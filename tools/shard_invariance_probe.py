"""Does the device-resident optimiser follow the same trajectory for every world size?
torchrun --nproc-per-node 2 tools/shard_invariance_probe.py [phase] : every rank runs the swarm
sharded over the 2 ranks AND a solo copy (a process group of its own), and compares fitness
vectors and best-fitness histories generation by generation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from psso_sac_for_powered_descent_b200 import envs, pso

phase = sys.argv[1] if len(sys.argv) > 1 else "landing_burn"
n, seeds, gens = int(os.environ.get("N_PART", 8192)), 8, 12
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
solo_groups = [dist.new_group([r]) for r in range(world)]
params = dict(pso.PSO_PARAMS[phase], pop_size=n, communication_freq=4, migration_freq=3, re_initialise_generation=8,
              re_initialise_number_of_particles=n // 2)
mk = lambda: envs.pso_wrapped_env(flight_phase=phase, enable_wind=True, stochastic_wind=True, precision="fp32",
                                  max_steps=2048, seed=99)
sh = pso.DeviceSwarm(mk(), n, params, n_seeds=seeds, seed=5, max_steps=2048)
so = pso.DeviceSwarm(mk(), n, params, n_seeds=seeds, seed=5, max_steps=2048, group=solo_groups[rank])
assert sh.world == world and so.world == 1
for g in range(gens):
    fa = sh.step().clone()
    fb = so.step().clone()
    same = torch.equal(fa[:sh.N_total] if g != params["re_initialise_generation"] else fa[:0],
                       fb[:so.N_total] if g != params["re_initialise_generation"] else fb[:0])
    nd = int((fa[:min(len(fa), len(fb))] != fb[:min(len(fa), len(fb))]).sum()) if len(fa) == len(fb) else -1
    if rank == 0:
        print(f"gen {g}: fitness vectors identical {same} (differing entries {nd}), best sharded {sh.global_best_fitness!r} "
              f"solo {so.global_best_fitness!r}, N {sh.N_total}/{so.N_total}", flush=True)
if rank == 0:
    print("histories identical:", sh.global_best_fitness_array == so.global_best_fitness_array)
dist.destroy_process_group()
