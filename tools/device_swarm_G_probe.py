"""Per-generation time of the device-resident optimiser on landing_burn (8 wind seeds) against the
hand-off thresholds of the rollout, one GPU:  python tools/device_swarm_G_probe.py [particles ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psso_sac_for_powered_descent_b200 import envs, pso as pso_mod, _native as N

G = "landing_burn"
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:          # under torchrun: the particles are sharded over the ranks as in bench.py
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
sizes = [int(a) for a in sys.argv[1:]] or [8192, 16384, 32768, 65536]
for n in sizes:
    for stages in (None, (0, 0), (16, 4096), (8, 4096), (24, 4096)):
        model = envs.pso_wrapped_env(flight_phase=G, enable_wind=True, stochastic_wind=True, max_steps=4096, seed=99)
        if stages is not None:
            N.check(model._b.lib.pd_set_rollout_stages(model._b._h, *stages))
        params = dict(pso_mod.PSO_PARAMS[G], pop_size=n, re_initialise_generation=10 ** 9)
        sw = pso_mod.DeviceSwarm(model, n, params, n_seeds=8, seed=5, max_steps=4096)
        sw.step(); sw.step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(10):
            sw.step()
        e1.record()
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        if rank == 0:
          print(f"n {n} stages {stages}: {e0.elapsed_time(e1) / 10:.3f} ms/gen on the device, host enqueue {t_host * 100:.3f} ms/gen, "
              f"longest {int(sw.last_steps.max())}", flush=True)
        del sw, model
if world > 1:
    dist.destroy_process_group()
