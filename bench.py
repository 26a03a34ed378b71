#!/usr/bin/env python
"""bench.py - env-steps/s of the batched powered-descent step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # CUDA arm
    python bench.py --impl reference [--gpus N] [--steps K] ...     # CPU arm (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...               # one rank per GPU

Workload at every N: BASELINE config 2 per GPU - 65 536 envs of `landing_burn_pure_throttle`,
random U(-1,1) actions, ISA atmosphere, no wind, auto-reset, production fp32 build.  A bench
"step" is one fused step-kernel launch over the whole batch (65 536 env-steps per GPU);
envs are sharded across ranks with no data-path collective (weak scaling).

The JSON line carries: value (device-timed, actions resident in HBM, launches replayed from a
CUDA graph), e2e (same metric through the public step() call with pinned-host actions copied
in and obs/reward/flags copied out every step), roofline (HBM, algorithmic 160 B/env-step as
SURVEY.md 8d counts it), fp_roofline (the bound that actually applies: FP32/FP64 pipes),
cpu_baseline (the oracle port timed on this box's host cores), pso (fitness evals/s of the
persistent rollout kernel at the 4 096-particle swarm of config 3).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

P = "landing_burn_pure_throttle"
N_ENVS = 65536
ALGO_BYTES_PER_STEP = 160.0        # 40 words x 4 B, phase P, fp32
ALGO_FLOP_PER_STEP = 4800.0        # canonical flop per env-step (P, no wind)


def load_peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured"
    except Exception:
        return 6650.0, 1965.0, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    seed, n_steps, fast = args
    import numpy as np
    from oracle import pd_oracle as O
    env = O.OracleEnv(P, "pso", tables=O.Tables(fast_rbf=fast))
    rng = np.random.default_rng(seed)
    env.reset()
    t0 = time.perf_counter()
    done_steps = 0
    for _ in range(n_steps):
        a = rng.uniform(-1, 1, 1).astype(np.float32)
        s, r, d, tr, info = env.step(a)
        done_steps += 1
        if d or tr:
            env.reset()
    return done_steps, time.perf_counter() - t0


def _cpu_pso_worker(args):
    """evaluate_worker_function of the reference: a fresh model per particle, one episode."""
    seed, phase = args
    import numpy as np
    import torch
    torch.set_num_threads(1)          # one worker per core; avoids intra-op oversubscription
    from oracle import pd_oracle as O
    t0 = time.perf_counter()
    model = O.PsoModel(phase, tables=O.Tables(fast_rbf=False), max_steps=4096)
    pos = np.random.default_rng(seed).uniform(-1.5, 1.5, model.n_params)
    f = model.objective_function(pos)
    return model.steps, time.perf_counter() - t0, f


def cpu_pso_rate(phase, per_core=4, cores=None):
    """PSO fitness evals/s of the oracle port under multiprocessing.Pool(all cores) - the
    reference's parallel_evaluate structure (a new env per particle, one episode each)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(i, 2, False) for i in range(cores)])
        t0 = time.perf_counter()
        res = pool.map(_cpu_pso_worker, [(500 + i, phase) for i in range(cores * per_core)])
        dt = time.perf_counter() - t0
    return len(res) / dt, sum(r[0] for r in res) / dt, cores, len(res)


def cpu_rate(n_steps_per_core, cores=None, repeats=1):
    """The oracle port (same scipy RBFInterpolator-per-call cost structure as the reference's
    env.step) on `cores` processes; returns env-steps/s aggregate."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(i, 2, False) for i in range(cores)])     # import + table warm-up
        best = 0.0
        for rep in range(repeats):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [(1000 * rep + i, n_steps_per_core, False) for i in range(cores)])
            dt = time.perf_counter() - t0
            best = max(best, sum(r[0] for r in res) / dt)
    return best, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_core = 20
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(i, 2, False) for i in range(cores)])
        for w in range(args.warmup):
            pool.map(_cpu_worker, [(10_000 + 100 * w + i, per_core, False) for i in range(cores)])
        t0 = time.perf_counter()
        total = 0
        for k in range(args.steps):
            res = pool.map(_cpu_worker, [(100 * k + i, per_core, False) for i in range(cores)])
            total += sum(r[0] for r in res)
        dt = time.perf_counter() - t0
    val = total / dt
    sample = (f"{cores} processes x {per_core} env.step calls per bench step, oracle/pd_oracle.py "
              f"(scalar Python port; scipy RBFInterpolator per call as upstream), random U(-1,1) actions")
    line = {"impl": "reference", "metric": "env_steps_per_sec", "value": val, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config2: landing_burn_pure_throttle env.step, random actions, "
                                   "ISA, no wind (bounded CPU sample)"},
            "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from psso_sac_for_powered_descent_b200 import envs, _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 0)
    B = args.envs
    lib = _native.load_library()
    env = envs.BatchedRocketEnv(B, "pso", P, precision=args.precision, auto_reset=True, device=local,
                                seed=1234 + rank)
    gen = torch.Generator(device=dev)
    gen.manual_seed(rank)
    tape = torch.rand(K + W, B, 1, device=dev, generator=gen, dtype=torch.float32) * 2 - 1
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-timed region: K fused step launches replayed from one CUDA graph
    with torch.cuda.stream(stream):
        env.reset()
        for w in range(W):
            env.step(tape[w])
        stream.synchronize()
        env.check_status()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for k in range(K):
                env.step(tape[W + k])
        sampler = ClockSampler(local)
        barrier()
        if rank == 0:
            sampler.start()
        l0 = lib.pd_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        graph.replay()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        env.check_status()
    launches = K          # K step-kernel launches inside the replayed graph
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * K / (ms_max * 1e-3)

    # ---- per-launch duration of the dominant kernel, L2 flushed between launches
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    durs = []
    with torch.cuda.stream(stream):
        for k in range(min(20, K)):
            flush.fill_(k & 0xFF)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            env.step(tape[W + k])
            a1.record(stream)
            stream.synchronize()
            durs.append(a0.elapsed_time(a1))
    durs.sort()
    kern_ms = durs[len(durs) // 2]
    del flush

    # ---- end-to-end through the public host-facing call: numpy actions in, numpy results out,
    # every step (BatchedRocketEnv.step_host: the step kernel reads the actions from and stores its
    # results to mapped pinned host memory - the same bytes over PCIe, no staging copies)
    host_tape = tape.cpu().pin_memory()       # the host's actions live in pinned memory
    Ke = min(K, 300)
    env.reset()
    torch.cuda.synchronize(dev)
    for w in range(min(W, 3) + 1):
        env.step_host(host_tape[w])
    barrier()
    t0 = time.perf_counter()
    chk = 0.0
    for k in range(Ke):
        obs, rew, done, trunc, tid = env.step_host(host_tape[W + k])
        chk += float(rew[0])                      # the caller reads the result before the next action
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * B * Ke / float(t.item())
    h2d = B * 4
    d2h = int(env._host["n_copy"])

    # ---- PSO fitness evaluation (config 3 / 5): the swarm is block-sharded over the ranks, each
    # rank rolls out its particles (x wind seeds) in one persistent kernel, then one fp64
    # all-gather of the fitness vector and one broadcast of the best position (NCCL).
    pso = None
    if not args.no_pso:
        from psso_sac_for_powered_descent_b200 import pso as pso_mod
        phase = args.pso_phase
        n_par = 249 if phase == P else 372
        model = envs.pso_wrapped_env(flight_phase=phase, enable_wind=args.pso_wind,
                                     stochastic_wind=args.pso_wind, precision=args.precision,
                                     max_steps=4096, seed=99)
        rngp = np.random.default_rng(7)
        pos = rngp.uniform(-1.5, 1.5, (args.particles, n_par))        # identical on every rank
        stats = {}

        def local_eval(p):
            fit, steps, tid = model.evaluate(p, n_seeds=args.seeds)
            stats["steps"] = float(steps.sum())
            stats["capped"] = int((tid < 0).sum())
            return fit.reshape(len(p), args.seeds).mean(dim=1).cpu().numpy()
        ev = pso_mod.ShardedEvaluator(local_eval)
        lo, hi = pso_mod.shard_bounds(args.particles, world, rank)
        fw = ev(pos)                                                # warm-up (weights upload, caches,
        ev.broadcast_best(fw, pos[lo:hi], args.particles)           # NCCL channels of both collectives)
        barrier()
        t0 = time.perf_counter()
        fitness = ev(pos)
        idx, best, best_pos = ev.broadcast_best(fitness, pos[lo:hi], args.particles)
        barrier()
        dt = time.perf_counter() - t0
        # north_star's swarm size: 65 536 particles through the same sharded evaluation
        big = None
        if args.particles < 65536 and not args.no_pso_scale:
            pos_big = rngp.uniform(-1.5, 1.5, (65536, n_par))
            ev(pos_big)
            barrier()
            b0 = time.perf_counter()
            fb = ev(pos_big)
            barrier()
            bdt = time.perf_counter() - b0
            tb = torch.tensor([bdt, stats["steps"]], device=dev, dtype=torch.float64)
            if world > 1:
                mb = tb.clone(); dist.all_reduce(mb, op=dist.ReduceOp.MAX)
                sb = tb.clone(); dist.all_reduce(sb, op=dist.ReduceOp.SUM)
                bdt, bsteps = float(mb[0]), float(sb[1])
            else:
                bsteps = stats["steps"]
            big = {"particles": 65536, "wind_seeds": args.seeds, "fitness_evals_per_s": 65536 / bdt, "ms": bdt * 1e3,
                   "env_steps_per_s": bsteps / bdt, "mean_episode_steps": bsteps / (65536 * args.seeds),
                   "what": "host list of positions in, fitness list out (ShardedEvaluator): weights upload, "
                           "rollout with straggler hand-off, fitness all-gather"}
            del pos_big, fb
        # device-resident optimiser (swarm never leaves HBM): whole generations
        params = dict(pso_mod.PSO_PARAMS[phase], pop_size=args.particles)
        sw = pso_mod.DeviceSwarm(model, args.particles, params, n_seeds=args.seeds, seed=5, max_steps=4096)
        sw.step()
        barrier()
        g0 = time.perf_counter()
        n_gen = 5
        gsteps = 0.0
        for _ in range(n_gen):
            sw.step()
            gsteps += float(sw.last_steps.sum())
        barrier()
        gdt = (time.perf_counter() - g0) / n_gen
        tt = torch.tensor([dt, stats["steps"], stats["capped"], gdt, gsteps / n_gen], device=dev, dtype=torch.float64)
        if world > 1:
            mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            dt, tot_steps, capped, gdt, gsteps = float(mx[0]), float(sm[1]), int(sm[2]), float(mx[3]), float(sm[4])
        else:
            tot_steps, capped, gsteps = stats["steps"], stats["capped"], gsteps / n_gen
        episodes = args.particles * args.seeds
        pso = {"phase": phase, "particles": args.particles, "wind_seeds": args.seeds, "wind": bool(args.pso_wind),
               "fitness_evals_per_s": args.particles / dt, "episodes_per_s": episodes / dt,
               "env_steps_per_s": tot_steps / dt, "ms": dt * 1e3, "mean_episode_steps": tot_steps / episodes,
               "episodes_hitting_step_cap": capped, "best_fitness": best, "best_index": idx,
               "collectives": "1 all_gather(fp64 fitness) + 1 broadcast(best position) per generation",
               "timing": "wall clock between device-synchronised barriers, max over ranks",
               "at_65536_particles": big,
               "device_swarm": {"ms_per_generation": gdt * 1e3, "fitness_evals_per_s": args.particles / gdt,
                                "env_steps_per_s": gsteps / gdt,
                                "what": "evaluate (rollout kernel) + fitness all-gather + sub-swarm best "
                                        "broadcast + pd_pso_update, swarm resident in HBM"}}

    # ---- SAC data collection (config 4 shape): shared 2-256-256-(1,1) actor on the tensor cores
    # + fused env step with stochastic wind, 131 072 envs per GPU, auto-reset
    sac = None
    if not args.no_sac:
        import torch.nn as nn
        torch.manual_seed(0)
        lin1, lin2, mean_l, lstd_l = nn.Linear(2, 256), nn.Linear(256, 256), nn.Linear(256, 1), nn.Linear(256, 1)
        actor = dict(w1=lin1.weight, b1=lin1.bias, w2=lin2.weight, b2=lin2.bias, wm=mean_l.weight,
                     bm=mean_l.bias, ws=lstd_l.weight, bs=lstd_l.bias, max_action=1.0)
        Bs, Ts = args.sac_envs, args.sac_steps
        senv = envs.BatchedRocketEnv(Bs, "rl", P, enable_wind=True, stochastic_wind=True,
                                     horiontal_wind_percentile=50, precision="fp32", auto_reset=True,
                                     device=local, seed=77 + rank)
        senv.collect(actor, 3, seed=1)
        best_ms = float("inf")
        for rep in range(3):          # best of 3: the first pass on a fresh box pays lazy module loads
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            s0.record()
            out = senv.collect(actor, Ts, seed=2 + rep)
            s1.record()
            barrier()
            best_ms = min(best_ms, s0.elapsed_time(s1))
        senv.check_status()
        sms = torch.tensor([best_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        sac = {"envs_per_gpu": Bs, "steps": Ts, "env_steps_per_s": world * Bs * Ts / (sms.item() * 1e-3),
               "ms_per_step": sms.item() / Ts, "actor": "2-256-256-(1,1), bf16 tcgen05 UMMA + fp32 heads",
               "wind": "stochastic, percentile 50, Philox gusts", "rtd": "rl",
               "mean_reward": float(out["rewards"].mean()), "resets": int(out["truncated"].sum() + out["done"].sum())}
        del senv, out

    # ---- the four further flight phases, RL closures, same fused step kernel
    phases = None
    if not args.no_phases:
        phases = {}
        for ph in ("subsonic", "supersonic", "ballistic_arc_descent", "landing_burn_pure_throttle_Pcontrol"):
            penv = envs.BatchedRocketEnv(B, "rl", ph, precision=args.precision, auto_reset=True, device=local,
                                         trajectory_length=1000, discount_factor=0.99, seed=5 + rank)
            ptape = torch.rand(40, B, penv.act_dim, device=dev, generator=gen, dtype=torch.float32) * 2 - 1
            with torch.cuda.stream(stream):
                penv.reset()
                for k in range(10):
                    penv.step(ptape[k])
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                p0.record(stream)
                for k in range(10, 40):
                    penv.step(ptape[k])
                p1.record(stream)
                barrier()
                penv.check_status()
            pms = torch.tensor([p0.elapsed_time(p1)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(pms, op=dist.ReduceOp.MAX)
            phases[ph] = {"env_steps_per_s": world * B * 30 / (pms.item() * 1e-3), "us_per_step": pms.item() / 30 * 1e3}
            del penv, ptape

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, sm_max, peak_src = load_peaks()
    achieved = ALGO_BYTES_PER_STEP * B / (kern_ms * 1e-3) / 1e9
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    fp_ach = ALGO_FLOP_PER_STEP * B / (kern_ms * 1e-3) / 1e12
    cpu = None
    if not args.no_cpu:
        v, cores = cpu_rate(args.cpu_steps_per_core)
        cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
               "sample": f"{cores} processes x {args.cpu_steps_per_core} env.step calls of oracle/pd_oracle.py "
                         "(scalar Python port of the reference env, scipy RBFInterpolator per call), "
                         "same phase / rtd / random-action workload"}
        if pso is not None:
            ev_s, st_s, c2, n_ep = cpu_pso_rate(args.pso_phase)
            cpu["pso_fitness_evals_per_s"] = ev_s
            cpu["pso_env_steps_per_s"] = st_s
            cpu["pso_sample"] = (f"{n_ep} random particles, Pool({c2}) over the oracle's objective_function with a "
                                 "fresh model per particle (the reference's evaluate_worker_function structure)")
            pso["vs_cpu_port"] = pso["fitness_evals_per_s"] / ev_s
            if pso.get("at_65536_particles"):
                pso["at_65536_particles"]["vs_cpu_port"] = pso["at_65536_particles"]["fitness_evals_per_s"] / ev_s
    line = {
        "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
        "config": {"workload": f"config2: {B} envs/GPU x {K} steps, landing_burn_pure_throttle, pso rtd, "
                               "random U(-1,1) float32 actions, ISA, no wind, auto-reset",
                   "envs_per_gpu": B, "precision_build": args.precision,
                   "l2": "state (~19 MB/step at 65 536 envs) stays L2-resident between launches as in the "
                         "real rollout; the kernel is FP-pipe bound; roofline launch durations are "
                         "measured with a 256 MB L2 flush between launches",
                   "launch": "K step launches captured in one CUDA graph"},
        "e2e": {"value": e2e_val, "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": Ke,
                "transport": {"zc_all": "step kernel reads the pinned host actions and stores its results to "
                                        "pinned host memory directly (mapped, same bytes over PCIe)",
                              "zc_out": "H2D copy of the actions, results stored to mapped pinned memory",
                              "copy": "H2D copy, kernel, one D2H copy (CUDA graph)"}[env._host["mode"]]},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": None, "peak_source": peak_src,
                     "kernel": f"step_kernel<{args.precision}>", "kernel_ms": kern_ms,
                     "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_STEP,
                     "note": "HBM is not the limiter of this path (SURVEY 8d); see fp_roofline"},
        "fp_roofline": {"bound": "fp32_pipe", "achieved": fp_ach, "peak": fp32_peak, "unit": "TFLOP/s",
                        "frac": fp_ach / fp32_peak, "flop_per_env_step": ALGO_FLOP_PER_STEP,
                        "peak_def": f"148 SM x 128 lanes x 2 x {sm_max:.0f} MHz", "sm_mhz_under_load": sm_mhz},
        "cpu_baseline": cpu,
        "pso": pso,
        "sac_collect": sac,
        "other_phases": phases,
    }
    try:      # ncu figures of the same kernel (profiles/, captured by the builder, not live)
        with open(os.path.join(REPO, "profiles", "step_kernel_ncu.json")) as f:
            nc = json.load(f)
        line["roofline"]["traffic"] = nc.get("dram_bytes_per_launch")
        line["roofline"]["traffic_source"] = nc.get("source")
        line["fp_roofline"]["fp64_pipe_pct_ncu"] = nc.get("fp64_pipe_pct")
        line["fp_roofline"]["issue_active_pct_ncu"] = nc.get("issue_active_pct")
    except Exception:
        pass
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--envs", type=int, default=N_ENVS)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--particles", type=int, default=4096)
    ap.add_argument("--seeds", type=int, default=1)
    ap.add_argument("--pso-phase", default=P, choices=[P, "landing_burn"])
    ap.add_argument("--pso-wind", action="store_true")
    ap.add_argument("--cpu-steps-per-core", type=int, default=1500)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pso", action="store_true")
    ap.add_argument("--no-sac", action="store_true")
    ap.add_argument("--no-phases", action="store_true")
    ap.add_argument("--no-pso-scale", action="store_true")
    ap.add_argument("--sac-envs", type=int, default=131072)
    ap.add_argument("--sac-steps", type=int, default=40)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
