"""PSO fitness evaluation and sub-swarm optimisation on top of the persistent rollout kernel.

Mirrors src/particle_swarm_optimisation/particle_swarm_optimisation.py:

    evaluate_worker_function(args)          :18-26    -> one rollout (kept for API parity)
    ParticleSubswarmOptimisation            :285-848
        .parallel_evaluate(positions)       :334-350  -> ONE kernel launch for the whole list
        .run() / __call__                   :413-515  velocity/position update, sub-swarm bests,
        .share_information / .migrate_particles / .re_initialise_swarms   (:523-552, :375-386)
        .save / .load_swarms / .save_results                              (:647-714)

Sharding: with torch.distributed initialised, `ShardedEvaluator` gives rank r
the contiguous particle block [r*N/W, (r+1)*N/W) (all wind seeds of a particle stay on one
GPU), evaluates it locally and all-gathers the fp64 fitness vector - the only collective on
the data path.  Every rank then derives the same arg-min; `broadcast_best` ships the winning
position from its owner when positions themselves are sharded.  The collective layer is plain
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import json
import os
import pickle
import random
from datetime import datetime

import numpy as np

# configs/evolutionary_algorithms_config.py:69-101 (knobs kept unchanged)
landing_burn_pure_throttle_pso_params = {
    'pop_size': 150, 'generations': 400, 'c1': 1, 'c2': 1, 'w_start': 0.9, 'w_end': 0.4,
    'fitness_threshold': -100000000, 'num_sub_swarms': 2, 'communication_freq': 10,
    'migration_freq': 5, 'number_of_migrants': 1, 're_initialise_number_of_particles': 600,
    're_initialise_generation': 90,
}
landing_burn_pso_params = {
    'pop_size': 200, 'generations': 400, 'c1': 1, 'c2': 1, 'w_start': 0.9, 'w_end': 0.7,
    'fitness_threshold': -1000, 'num_sub_swarms': 2, 'communication_freq': 10,
    'migration_freq': 5, 'number_of_migrants': 1, 're_initialise_number_of_particles': 600,
    're_initialise_generation': 90,
}
PSO_PARAMS = {'landing_burn_pure_throttle': landing_burn_pure_throttle_pso_params,
              'landing_burn': landing_burn_pso_params}


def shard_bounds(n, world, rank):
    """Contiguous block partition; the first n % world ranks take one extra particle."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


# ---------------------------------------------------------------------------------------
# counter-based random streams shared by the host drop-in (rng='philox') and the device swarm
# ---------------------------------------------------------------------------------------
def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on numpy uint32 arrays (same rounds / constants as csrc/pd_pso.cu)."""
    c0, c1, c2, c3 = [np.asarray(v, dtype=np.uint64) & 0xFFFFFFFF for v in np.broadcast_arrays(c0, c1, c2, c3)]
    k0, k1 = np.uint64(int(k0) & 0xFFFFFFFF), np.uint64(int(k1) & 0xFFFFFFFF)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & MASK, (k1 + np.uint64(0xBB67AE85)) & MASK
    return c0, c1, c2, c3


def _u53(a, b):
    return ((((a << np.uint64(32)) | b) >> np.uint64(11)).astype(np.float64)) * (1.0 / 9007199254740992.0)


class PsoStreams:
    """Random numbers of the optimiser as pure functions of (seed, generation, index): the velocity
    update's r1, r2 per particle (the device kernel draws the same values), the sharing coin per
    sub-swarm and the migration draws - identical on every rank and for every sharding."""

    TAG_UPDATE, TAG_SHARE, TAG_MIGRATE = 0x50534F55, 0x53484152, 0x4D494752

    def __init__(self, seed):
        self.k0, self.k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF

    def r1r2(self, index, generation):
        r = philox4x32(np.asarray(index), generation, self.TAG_UPDATE, 0, self.k0, self.k1)
        return _u53(r[0], r[1]), _u53(r[2], r[3])

    def coin(self, swarm, generation):
        r = philox4x32(np.array([swarm]), generation, self.TAG_SHARE, 0, self.k0, self.k1)
        return float(_u53(r[0], r[1])[0])

    def migrant(self, swarm, generation, m):
        r = philox4x32(np.array([swarm]), generation, self.TAG_MIGRATE, m, self.k0, self.k1)
        return float(_u53(r[0], r[1])[0]), float(_u53(r[2], r[3])[0])


def _dist():
    try:
        import torch.distributed as dist
        return dist if dist.is_available() and dist.is_initialized() else None
    except Exception:
        return None


class ShardedEvaluator:
    """fitness = evaluate(positions) with the particle list block-sharded over the ranks of a
    torch.distributed process group.  `local_eval(np.ndarray[n_local, P]) -> np.ndarray[n_local]`
    is the per-rank evaluator (the CUDA rollout in production); if it accepts `index0` /
    `generation` keywords it is told the global index of its first particle and the PSO generation
    (the gust-noise stream of a windy evaluation must not depend on the sharding)."""

    def __init__(self, local_eval, group=None):
        import inspect
        self.local_eval = local_eval
        self.group = group
        try:
            names = inspect.signature(local_eval).parameters
        except (TypeError, ValueError):
            names = {}
        self._kw = {k for k in ("index0", "generation") if k in names}
        self.dist = _dist()
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0

    def _device(self):
        import torch
        if self.dist and self.dist.get_backend(self.group) == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def __call__(self, positions, index_base=0, generation=0):
        import torch
        pos = np.asarray(positions, dtype=np.float64)
        n = len(pos)
        lo, hi = shard_bounds(n, self.world, self.rank)
        kw = {}
        if "index0" in self._kw:
            kw["index0"] = index_base + lo
        if "generation" in self._kw:
            kw["generation"] = generation
        local = np.asarray(self.local_eval(pos[lo:hi], **kw), dtype=np.float64) if hi > lo else np.zeros(0)
        if self.world == 1:
            return local
        dev = self._device()
        width = -(-n // self.world)                    # padded so all_gather is regular
        buf = torch.full((width,), float("inf"), dtype=torch.float64, device=dev)
        buf[:hi - lo] = torch.as_tensor(local, dtype=torch.float64, device=dev)
        out = torch.empty(self.world * width, dtype=torch.float64, device=dev)
        self.dist.all_gather_into_tensor(out, buf, group=self.group)
        out = out.cpu().numpy().reshape(self.world, width)
        return np.concatenate([out[r, :shard_bounds(n, self.world, r)[1] - shard_bounds(n, self.world, r)[0]]
                               for r in range(self.world)])

    def broadcast_best(self, fitness, local_positions, n_total):
        """Global arg-min and its position, broadcast from the rank that owns it (used when the
        positions are held sharded).  Returns (index, fitness, position[P])."""
        import torch
        idx = int(np.argmin(fitness))
        if self.world == 1:
            return idx, float(fitness[idx]), np.array(local_positions[idx], dtype=np.float64)
        owner = next(r for r in range(self.world) if shard_bounds(n_total, self.world, r)[0] <= idx
                     < shard_bounds(n_total, self.world, r)[1])
        dev = self._device()
        P = local_positions.shape[1]
        t = torch.zeros(P, dtype=torch.float64, device=dev)
        if self.rank == owner:
            lo, _ = shard_bounds(n_total, self.world, owner)
            t.copy_(torch.as_tensor(local_positions[idx - lo], dtype=torch.float64))
        self.dist.broadcast(t, src=self.dist.get_global_rank(self.group, owner) if self.group else owner,
                            group=self.group)
        return idx, float(fitness[idx]), t.cpu().numpy()

    def agree(self, array, what="value"):
        """Raise on every rank unless `array` is bit-identical on all of them."""
        if self.world == 1:
            return
        import hashlib
        import torch
        h = hashlib.sha256(np.ascontiguousarray(array).tobytes()).digest()[:8]
        v = torch.tensor([int.from_bytes(h, "little") >> 1], dtype=torch.int64, device=self._device())
        lo, hi = v.clone(), v.clone()
        self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN, group=self.group)
        self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX, group=self.group)
        if int(lo) != int(hi):
            raise RuntimeError(f"ParticleSubswarmOptimisation: {what} differs between ranks - every rank must "
                               "build the same swarm (pass the same seed, or let rank 0 broadcast one)")

    def shared_seed(self):
        """A fresh seed drawn on rank 0 and broadcast (seed=None under torch.distributed)."""
        import torch
        s = int.from_bytes(os.urandom(4), "little")
        if self.world == 1:
            return s
        t = torch.tensor([s], dtype=torch.int64, device=self._device())
        src = self.dist.get_global_rank(self.group, 0) if self.group else 0
        self.dist.broadcast(t, src=src, group=self.group)
        return int(t)


def evaluate_worker_function(args):
    """Same tuple as the reference worker; evaluates one particle on the GPU."""
    from .envs import pso_wrapped_env
    position, flight_phase, enable_wind, stochastic_wind, horiontal_wind_percentile = args
    model = pso_wrapped_env(flight_phase=flight_phase, enable_wind=enable_wind,
                            stochastic_wind=stochastic_wind,
                            horiontal_wind_percentile=horiontal_wind_percentile)
    return model.objective_function(position)


def _csv_append(path, header, row, first):
    import csv
    with open(path, "w" if first else "a", newline="") as f:
        w = csv.writer(f)
        if first:
            w.writerow(header)
        w.writerow(row)


def write_generation_metrics(metrics_dir, generation, swarm_rows, global_best, global_avg):
    """save_generation_metrics (particle_swarm_optimisation.py:811-833): one appended row per
    generation in metrics/subswarm_<i>_metrics.csv and metrics/global_metrics.csv.
    swarm_rows[k] = (best_fitness, avg, min, max, std, num_particles)."""
    os.makedirs(metrics_dir, exist_ok=True)
    hdr = ["swarm_idx", "best_fitness", "avg_fitness", "min_fitness", "max_fitness", "std_fitness",
           "num_particles", "generation", "global_best_fitness", "global_avg_fitness"]
    for k, r in enumerate(swarm_rows):
        _csv_append(f"{metrics_dir}/subswarm_{k}_metrics.csv", hdr,
                    [k] + [repr(float(v)) for v in r[:5]] + [int(r[5]), generation, repr(float(global_best)),
                                                             repr(float(global_avg))], generation == 0)
    _csv_append(f"{metrics_dir}/global_metrics.csv", ["generation", "global_best_fitness", "global_avg_fitness"],
                [generation, repr(float(global_best)), repr(float(global_avg))], generation == 0)


def write_fitness_history(metrics_dir, global_best, average, sub_best, sub_avg):
    """save_fitness_history (:717-744): metrics/fitness_history.csv."""
    import csv
    os.makedirs(metrics_dir, exist_ok=True)
    n = len(global_best)
    cols, names = [list(range(n)), list(global_best), list(average)], ["Generation", "Global_Best_Fitness",
                                                                        "Average_Fitness"]
    for i, (b, a) in enumerate(zip(sub_best, sub_avg)):
        for label, arr in ((f"Subswarm_{i + 1}_Best", b), (f"Subswarm_{i + 1}_Average", a)):
            if len(arr) > 0:
                cols.append((list(arr) + [""] * n)[:n])
                names.append(label)
    with open(f"{metrics_dir}/fitness_history.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(names)
        for r in zip(*cols):
            w.writerow([v if isinstance(v, (int, str)) else repr(float(v)) for v in r])


def write_results_csv(base_save_dir, names, best_position, best_fitness):
    """particle_subswarm_optimisation_results.csv with the reference's column names
    (`<layer>_weight_<j>` / `<layer>_bias_<j>`, env_wrapped_ea.py:61-75, written :690-714) so that
    src/particle_swarm_optimisation/network_loader.py keeps working."""
    import csv
    os.makedirs(base_save_dir, exist_ok=True)
    with open(f'{base_save_dir}/particle_subswarm_optimisation_results.csv', 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['Algorithm'] + list(names) + ['Best Fitness'])
        w.writerow(['Particle Subswarm Optimisation'] + [repr(float(v)) for v in best_position]
                   + [repr(float(best_fitness))])


class ParticleSubswarmOptimisation:
    """Drop-in for the reference class; the swarm lives in flat arrays instead of a list of
    dicts (converted on save/load so `swarm.pkl` stays interchangeable), `members[k]` keeps the
    reference's per-sub-swarm list order.

    Extra keyword arguments: `model` (anything with .bounds and .evaluate(positions) or
    .objective_function), `pso_params` override, `seed`, `n_seeds` (wind seeds per particle;
    fitness = mean over seeds - the reference evaluates one stochastic draw), `base_save_dir`,
    `rng`: 'reference' = the reference's own generators and draw order (`random` for the initial
    positions, the sharing coin and the migration; `np.random.rand()` twice per particle in swarm
    order for the velocity update) - the global ones when seed is None, exactly as upstream; 'philox'
    = counter-based streams shared with `DeviceSwarm`, which then follows the identical trajectory.
    Under torch.distributed every rank must hold the same swarm: with seed=None rank 0 draws a seed
    and broadcasts it, and the initial positions are checked to agree across ranks.
    """

    def __init__(self, flight_phase, save_interval=5, enable_wind=False, stochastic_wind=False,
                 horiontal_wind_percentile=50, load_swarms=False, use_multiprocessing=True,
                 num_processes=None, model=None, pso_params=None, seed=None, n_seeds=1,
                 precision="fp32", base_save_dir=None, evaluator=None, rng="reference", write_metrics=True):
        assert flight_phase in ['subsonic', 'supersonic', 'flip_over_boostbackburn',
                                'ballistic_arc_descent', 'landing_burn', 'landing_burn_pure_throttle']
        assert rng in ("reference", "philox")
        self.flight_phase = flight_phase
        self.pso_params = dict(pso_params or PSO_PARAMS[flight_phase])
        self.enable_wind, self.stochastic_wind = enable_wind, stochastic_wind
        self.horiontal_wind_percentile = horiontal_wind_percentile
        self.use_multiprocessing, self.num_processes = use_multiprocessing, num_processes
        self.save_interval, self.n_seeds = save_interval, n_seeds
        if model is None:
            from .envs import pso_wrapped_env
            model = pso_wrapped_env(flight_phase, enable_wind=enable_wind, stochastic_wind=stochastic_wind,
                                    horiontal_wind_percentile=horiontal_wind_percentile,
                                    precision=precision)
        self.model = model
        import inspect
        try:
            self._eval_kw = {k for k in ("index0", "generation")
                             if k in inspect.signature(model.evaluate).parameters} if hasattr(model, "evaluate") else set()
        except (TypeError, ValueError):
            self._eval_kw = set()
        self.bounds = list(model.bounds)
        self.lower = np.array([b[0] for b in self.bounds], dtype=np.float64)
        self.upper = np.array([b[1] for b in self.bounds], dtype=np.float64)
        p = self.pso_params
        self.pop_size, self.generations = p['pop_size'], p['generations']
        self.w_start, self.w_end, self.c1, self.c2 = p['w_start'], p['w_end'], p['c1'], p['c2']
        self.num_sub_swarms = p['num_sub_swarms']
        self.communication_freq = p.get('communication_freq', 10)
        self.migration_freq = p.get('migration_freq', 20)
        self.number_of_migrants = p.get('number_of_migrants', 1)
        self.re_initialise_number_of_particles = p.get('re_initialise_number_of_particles', 500)
        self.re_initialise_generation = p.get('re_initialise_generation', 60)
        self.evaluator = evaluator or ShardedEvaluator(self._local_eval)
        if seed is None and getattr(self.evaluator, "world", 1) > 1:
            seed = self.evaluator.shared_seed()       # every rank must build the same swarm
        self.seed, self.rng_mode = seed, rng
        self.rng_py = random.Random(seed) if seed is not None else random
        self.rng_np = np.random.RandomState(seed) if seed is not None else np.random
        self.streams = PsoStreams(seed if seed is not None else 0)
        self.w = self.w_start
        self.generation = 0
        self.capped_episodes = 0
        self.write_metrics = write_metrics
        self.timestamp = datetime.now().strftime('%Y-%m-%d_%H-%M-%S')
        self.base_save_dir = base_save_dir or f'data/pso_saves/{flight_phase}/run_{self.timestamp}'
        self.metrics_dir = f'{self.base_save_dir}/metrics'
        self.initialize_swarms()
        if hasattr(self.evaluator, "agree"):
            self.evaluator.agree(self.position, "the initial swarm")
        if load_swarms:
            self.load_swarms()

    @property
    def _rank(self):
        return getattr(self.evaluator, "rank", 0)

    # ------------------------------------------------------------------ evaluation
    def _local_eval(self, positions, index0=0, generation=0):
        if hasattr(self.model, "evaluate"):
            kw = {}
            if "index0" in self._eval_kw:
                kw["index0"] = index0
            if "generation" in self._eval_kw:
                kw["generation"] = generation
            out = self.model.evaluate(positions, n_seeds=self.n_seeds, **kw)
            self.capped_episodes += int(getattr(self.model, "capped", 0))
            fit = out[0].detach().cpu().numpy() if hasattr(out[0], "detach") else np.asarray(out[0])
            fit = fit.reshape(len(positions), self.n_seeds)
            if self.n_seeds == 1:
                return fit[:, 0]
            tot = fit[:, 0].copy()                     # sequential sum, as pd_pso_seed_mean
            for k in range(1, self.n_seeds):
                tot = tot + fit[:, k]
            return tot / self.n_seeds
        return np.array([self.model.objective_function(p) for p in positions])

    def parallel_evaluate(self, positions, index_base=0):
        """list[np.ndarray(P)] -> list[float], order-preserving (reference :334-350)."""
        if len(positions) == 0:
            return []
        pos = np.asarray(positions, dtype=np.float64)
        try:
            fit = self.evaluator(pos, index_base=index_base, generation=self.generation)
        except TypeError:                              # a plain callable(positions)
            fit = self.evaluator(pos)
        return [float(f) for f in fit]

    # ------------------------------------------------------------------ swarm state
    def initialize_swarms(self):
        """Same draws as the reference: its base class first builds a plain `pop_size` swarm that the
        sub-swarm optimiser never uses (`ParticleSwarmOptimisation.__init__` -> `initialize_swarm`,
        :57, 76-90) - pop_size x P `random.uniform` draws are consumed - and only then come the
        sub-swarms (:389-411): sub-swarm, particle, bound.  With the same seed the swarm is therefore
        the one `random.seed(seed)` gives upstream (tests/golden/pso_run_reference.npz)."""
        n_sub = self.pop_size // self.num_sub_swarms
        N, P = n_sub * self.num_sub_swarms, len(self.bounds)
        for _ in range(self.pop_size):
            for b in self.bounds:
                self.rng_py.uniform(b[0], b[1])
        self.position = np.empty((N, P), dtype=np.float64)
        for i in range(N):
            for j, b in enumerate(self.bounds):
                self.position[i, j] = self.rng_py.uniform(b[0], b[1])
        self.velocity = np.zeros((N, P))
        self.best_position = np.full((N, P), np.nan)
        self.best_fitness = np.full(N, np.inf)
        self.members = [list(range(k * n_sub, (k + 1) * n_sub)) for k in range(self.num_sub_swarms)]
        self.subswarm_best_positions = [None] * self.num_sub_swarms
        self.subswarm_best_fitnesses = [float('inf')] * self.num_sub_swarms
        self.global_best_position, self.global_best_fitness = None, float('inf')
        self.global_best_fitness_array, self.global_best_position_array = [], []
        self.average_particle_fitness_array = []
        self.subswarm_best_fitness_array = [[] for _ in range(self.num_sub_swarms)]
        self.subswarm_avg_array = [[] for _ in range(self.num_sub_swarms)]

    @property
    def swarm_of(self):
        out = np.full(len(self.position), -1, dtype=np.int64)
        for k, m in enumerate(self.members):
            out[m] = k
        return out

    @property
    def swarms(self):
        """Reference layout: list[list[dict(position, velocity, best_position, best_fitness)]]."""
        out = []
        for m in self.members:
            sw = []
            for i in m:
                bp = None if np.isnan(self.best_position[i, 0]) else self.best_position[i].copy()
                sw.append({'position': self.position[i].copy(), 'velocity': self.velocity[i].copy(),
                           'best_position': bp, 'best_fitness': float(self.best_fitness[i])})
            out.append(sw)
        return out

    @swarms.setter
    def swarms(self, swarms):
        parts = [(k, p) for k, sw in enumerate(swarms) for p in sw]
        P = len(self.bounds)
        self.position = np.array([p['position'] for _, p in parts], dtype=np.float64).reshape(-1, P)
        self.velocity = np.array([p['velocity'] for _, p in parts], dtype=np.float64).reshape(-1, P)
        self.best_position = np.array([p['best_position'] if p['best_position'] is not None
                                       else np.full(P, np.nan) for _, p in parts]).reshape(-1, P)
        self.best_fitness = np.array([p['best_fitness'] for _, p in parts], dtype=np.float64)
        self.members, i = [], 0
        for sw in swarms:
            self.members.append(list(range(i, i + len(sw))))
            i += len(sw)

    def weight_linear_decrease(self, generation):
        return self.w_start - (self.w_start - self.w_end) * generation / self.generations

    # ------------------------------------------------------------------ one generation
    def step_generation(self, generation):
        self.generation = generation
        N = len(self.position)
        fitness = np.asarray(self.parallel_evaluate(list(self.position)))
        improved = fitness < self.best_fitness
        self.best_fitness[improved] = fitness[improved]
        self.best_position[improved] = self.position[improved]
        rows = []
        for k, m in enumerate(self.members):
            if len(m) == 0:
                rows.append((self.subswarm_best_fitnesses[k], np.nan, np.nan, np.nan, np.nan, 0))
                continue
            idx = np.asarray(m)
            j = idx[np.argmin(fitness[idx])]          # first occurrence in list order, as the sequential scan
            if fitness[j] < self.subswarm_best_fitnesses[k]:
                self.subswarm_best_fitnesses[k] = float(fitness[j])
                self.subswarm_best_positions[k] = self.position[j].copy()
            f = fitness[idx]
            rows.append((self.subswarm_best_fitnesses[k], float(np.mean(f)), float(np.min(f)), float(np.max(f)),
                         float(np.std(f)), len(idx)))
            self.subswarm_best_fitness_array[k].append(self.subswarm_best_fitnesses[k])
            self.subswarm_avg_array[k].append(float(np.mean(f)))
        for k, f in enumerate(self.subswarm_best_fitnesses):
            if f < self.global_best_fitness:
                self.global_best_fitness = f
                self.global_best_position = self.subswarm_best_positions[k].copy()
        order = np.concatenate([np.asarray(m, dtype=np.int64) for m in self.members])
        avg = float(np.mean(fitness[order]))
        self.average_particle_fitness_array.append(avg)
        if self.write_metrics and self._rank == 0:
            write_generation_metrics(self.metrics_dir, generation, rows, self.global_best_fitness, avg)
        self.w = self.weight_linear_decrease(generation)
        # velocity update with the *sub-swarm* best; one scalar r1, r2 per particle (:517-521)
        if self.rng_mode == "philox":
            r1, r2 = self.streams.r1r2(np.arange(N), generation)
        else:       # np.random.rand() for the cognitive then the social term, particle by particle in swarm order
            draws = self.rng_np.rand(2 * len(order))
            r1, r2 = np.empty(N), np.empty(N)
            r1[order], r2[order] = draws[0::2], draws[1::2]
        swarm_of = self.swarm_of
        local_best = np.stack([self.subswarm_best_positions[k] for k in swarm_of])
        t1 = self.w * self.velocity
        t2 = (self.c1 * r1)[:, None] * (self.best_position - self.position)
        t3 = (self.c2 * r2)[:, None] * (local_best - self.position)
        self.velocity = (t1 + t2) + t3
        self.position = np.clip(self.position + self.velocity, self.lower, self.upper)
        if generation % self.communication_freq == 0 and generation > 0:
            self.share_information()
        if generation % self.migration_freq == 0 and generation > 0:
            self.migrate_particles()
        self.global_best_fitness_array.append(self.global_best_fitness)
        self.global_best_position_array.append(self.global_best_position)
        if generation == self.re_initialise_generation:
            self.re_initialise_swarms()
        return fitness

    def run(self, generations=None):
        for generation in range(generations if generations is not None else self.generations):
            self.step_generation(generation)
            if self.save_interval and generation % self.save_interval == 0 and generation != 0 \
                    and self._rank == 0:
                self.save()
                self.save_results()
        if self.capped_episodes:
            import warnings
            warnings.warn(f"{self.capped_episodes} episode(s) hit the rollout step cap during this run and "
                          "were scored as truncated at their final state", RuntimeWarning)
        return self.global_best_position, self.global_best_fitness

    __call__ = run

    def share_information(self):
        """:523-542 - blend the other sub-swarms' bests towards the best one (p = 0.5); the blended
        position replaces the sub-swarm best even when its fitness is worse."""
        influence_factor, sharing_probability = 0.3, 0.5
        best = int(np.argmin(self.subswarm_best_fitnesses))
        best_position = self.subswarm_best_positions[best]
        N = len(self.position)
        for i in range(self.num_sub_swarms):
            if i == best:
                continue
            u = self.streams.coin(i, self.generation) if self.rng_mode == "philox" else self.rng_py.random()
            if u < sharing_probability:
                self.subswarm_best_positions[i] = ((1 - influence_factor) * self.subswarm_best_positions[i]
                                                   + influence_factor * best_position)
                new_fitness = self.parallel_evaluate([self.subswarm_best_positions[i]], index_base=N + i)[0]
                if new_fitness < self.subswarm_best_fitnesses[i]:
                    self.subswarm_best_fitnesses[i] = new_fitness

    def _migration_draws(self, i, m, n_members):
        others = [k for k in range(self.num_sub_swarms) if k != i]
        if self.rng_mode == "philox":
            u1, u2 = self.streams.migrant(i, self.generation, m)
            return min(int(u1 * n_members), n_members - 1), others[min(int(u2 * len(others)), len(others) - 1)]
        return self.rng_py.randrange(n_members), self.rng_py.choice(others)

    def migrate_particles(self):
        """:544-552 - the migrant moves to the END of the target list (index relabelling)."""
        for i in range(self.num_sub_swarms):
            if len(self.members[i]) > 1:
                for m in range(self.number_of_migrants):
                    pick, target = self._migration_draws(i, m, len(self.members[i]))
                    self.members[target].append(self.members[i].pop(pick))

    def re_initialise_swarms(self):
        """:375-386 - each sub-swarm keeps its `re_initialise_number_of_particles // num_sub_swarms`
        best particles (stable sort by personal best), in that order."""
        keep_n = self.re_initialise_number_of_particles // self.num_sub_swarms
        new_members, rows = [], []
        for m in self.members:
            kept = sorted(m, key=lambda i: self.best_fitness[i])[:keep_n]
            new_members.append(list(range(len(rows), len(rows) + len(kept))))
            rows.extend(kept)
        rows = np.asarray(rows, dtype=np.int64)
        for name in ("position", "velocity", "best_position", "best_fitness"):
            setattr(self, name, getattr(self, name)[rows])
        self.members = new_members

    # ------------------------------------------------------------------ persistence
    def save(self):
        os.makedirs(f'{self.base_save_dir}/saves', exist_ok=True)
        with open(f'{self.base_save_dir}/saves/swarm.pkl', 'wb') as f:
            pickle.dump(self.swarms, f)

    def load_swarms(self, file_path=None):
        file_path = file_path or f'data/pso_saves/{self.flight_phase}/saves/swarm.pkl'
        with open(file_path, 'rb') as f:
            self.swarms = pickle.load(f)
        self.subswarm_best_positions = [None] * self.num_sub_swarms
        self.subswarm_best_fitnesses = [float('inf')] * self.num_sub_swarms
        self.global_best_fitness, self.global_best_position = float('inf'), None
        for k, m in enumerate(self.members):
            for i in m:
                f = self.best_fitness[i]
                if f < self.subswarm_best_fitnesses[k]:
                    self.subswarm_best_fitnesses[k] = float(f)
                    self.subswarm_best_positions[k] = self.best_position[i].copy()
                if f < self.global_best_fitness:
                    self.global_best_fitness, self.global_best_position = float(f), self.best_position[i].copy()

    def save_results(self):
        names = list(getattr(self.model, "mock_dictionary_of_opt_params",
                             {f"p_{j}": 0 for j in range(len(self.bounds))}).keys())
        write_results_csv(self.base_save_dir, names, self.global_best_position, self.global_best_fitness)
        self.save_fitness_history()
        self.save_config_to_json()

    def save_fitness_history(self):
        write_fitness_history(self.metrics_dir, self.global_best_fitness_array, self.average_particle_fitness_array,
                              self.subswarm_best_fitness_array, self.subswarm_avg_array)

    def save_config_to_json(self):
        """:835-848 plus the knobs this implementation adds."""
        config = dict(self.pso_params)
        config.update({'flight_phase': self.flight_phase, 'enable_wind': self.enable_wind,
                       'stochastic_wind': self.stochastic_wind,
                       'horiontal_wind_percentile': self.horiontal_wind_percentile,
                       'use_multiprocessing': self.use_multiprocessing, 'num_processes': self.num_processes,
                       'save_interval': self.save_interval, 'timestamp': self.timestamp,
                       'n_seeds': self.n_seeds, 'seed': self.seed, 'rng': self.rng_mode})
        os.makedirs(self.base_save_dir, exist_ok=True)
        with open(f'{self.base_save_dir}/pso_config.json', 'w') as f:
            json.dump(config, f, indent=4)

    # ------------------------------------------------------------------ trajectory export
    def collect_trajectory_data(self, individual, max_steps=8192):
        """One episode of `individual` with everything the reference records per step
        (particle_swarm_optimisation.py:759-785): observations ('states'), actions, rewards and the
        complete `info` dict.  Runs closed-loop in a single fp64 diagnostic env lane
        (envs.rocket_environment_pre_wrap, pd_set_info_mode) with the per-particle MLP evaluated on
        the host in float32 (simple_actor.forward, env_wrapped_ea.py:18-44) - a one-episode export
        path, not the throughput path (that is pd_rollout_pso)."""
        from .envs import rocket_environment_pre_wrap
        layers = self._unpack_actor(np.asarray(individual, dtype=np.float64))
        env = rocket_environment_pre_wrap(type="pso", flight_phase=self.flight_phase, enable_wind=False,
                                          precision="fp64")
        data = {'states': [], 'actions': [], 'rewards': [], 'info': []}
        state = env.reset()
        for _ in range(max_steps):
            obs = self._augment_state(state)
            h = np.asarray(obs, dtype=np.float32)
            for i, (W, b) in enumerate(layers):
                h = W @ h + b
                h = np.tanh(h) if i == len(layers) - 1 else np.maximum(h, np.float32(0))
            action = h.astype(np.float32)
            state, r, d, t, info = env.step(action)
            data['states'].append(obs)
            data['actions'].append(action.tolist())
            data['rewards'].append(r)
            data['info'].append(info)
            if d or t:
                break
        return data

    def _unpack_actor(self, individual):
        """named_parameters() order: per layer weight (out x in, row-major) then bias
        (env_wrapped_ea.py:46-59)."""
        i, o, n = (2, 1, 3) if self.flight_phase == 'landing_burn_pure_throttle' else (5, 4, 4)
        shapes = [(8, i)] + [(8, 8)] * n + [(o, 8)]
        layers, k = [], 0
        for (a, b) in shapes:
            W = individual[k:k + a * b].astype(np.float32).reshape(a, b); k += a * b
            bias = individual[k:k + a].astype(np.float32); k += a
            layers.append((W, bias))
        assert k == len(individual), "individual does not match the phase's actor"
        return layers

    def _augment_state(self, state):
        """pso_wrapper.augment_state (env_wrapped_ea.py:97-123)."""
        import math
        nv = self.model._b.params.norm_vals if hasattr(self.model, "_b") else None
        if nv is None:
            from .params import RocketParams
            nv = RocketParams.default().norm_vals
        x, y, vx, vy, theta = state[:5]
        if self.flight_phase == 'landing_burn_pure_throttle':
            return [y / nv[0], vy / nv[1]]
        k = float(np.arctanh(0.75) / math.radians(25))
        return [x / nv[5], y / nv[0], vx / nv[6], vy / nv[1], math.tanh(k * (theta - math.pi / 2))]

    def save_trajectory_data(self, trajectory_data, trajectory_dir=None):
        """states.csv / actions.csv / rewards.csv / info_data.csv with the reference's layout
        (particle_swarm_optimisation.py:787-809: nested dicts flattened with '_'-joined keys)."""
        import pandas as pd
        d = trajectory_dir or f'{self.base_save_dir}/trajectory_data'
        os.makedirs(d, exist_ok=True)
        pd.DataFrame(trajectory_data['states']).to_csv(f'{d}/states.csv', index=False)
        pd.DataFrame(trajectory_data['actions']).to_csv(f'{d}/actions.csv', index=False)
        pd.DataFrame(trajectory_data['rewards'], columns=['reward']).to_csv(f'{d}/rewards.csv', index=False)
        if trajectory_data['info']:
            flat_data = []
            for info in trajectory_data['info']:
                flat = {}

                def flatten(dct, prefix=''):
                    for key, value in dct.items():
                        if isinstance(value, dict):
                            flatten(value, f"{prefix}{key}_")
                        else:
                            flat[f"{prefix}{key}"] = value
                flatten(info)
                flat_data.append(flat)
            pd.DataFrame(flat_data).to_csv(f'{d}/info_data.csv', index=False)
        return d


class DeviceSwarm:
    """Device-resident sub-swarm PSO for large swarms (BASELINE config 5): positions, velocities,
    personal bests, sub-swarm / global bests and the per-generation metrics stay in HBM; a
    generation is [rollout kernel -> seed mean -> all-gather of the fp64 fitness slices -> select
    (per-sub-swarm arg-min + metrics) -> gather of the improved rows -> all-reduce (= broadcast of
    each improved sub-swarm best from its owner) -> apply -> pd_pso_update] with NO host
    synchronisation; `share_information` (re-evaluation of the blended bests, replicated on every
    rank), `migrate_particles` (index relabelling) and `re_initialise_swarms` (one compaction, the
    only step that reads device data back) follow the reference
    (particle_swarm_optimisation.py:523-552, 375-386).

    Particles are block-sharded in equal blocks of ceil(N / world); r1, r2 / the sharing coin / the
    migration draws are Philox functions of (seed, generation, global index), so the optimisation
    trajectory is identical for every world size and identical to the host drop-in
    `ParticleSubswarmOptimisation(rng='philox')` started from the same positions.
    """

    def __init__(self, model, n_particles, pso_params, n_seeds=1, seed=0, max_steps=4096, group=None,
                 positions=None, base_save_dir=None):
        import ctypes as C
        import torch
        from . import _native as N
        self.torch, self.C, self.N = torch, C, N
        self.model, self.n_seeds, self.seed, self.max_steps = model, n_seeds, int(seed), max_steps
        self.env = model._b
        self.dev = self.env.device
        self.lib = N.load_library()
        p = dict(pso_params)
        self.params = p
        self.P = len(model.bounds)
        self.S = p['num_sub_swarms']
        sub = int(n_particles) // self.S
        self.N_total = sub * self.S                       # initialize_swarms: pop_size // num_sub_swarms each
        self.lo_b, self.hi_b = float(model.bounds[0][0]), float(model.bounds[0][1])
        self.dist = _dist()
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.streams = PsoStreams(self.seed)
        self.base_save_dir = base_save_dir
        self.generation = 0
        self.members = [list(range(k * sub, (k + 1) * sub)) for k in range(self.S)]
        self._layout(self.N_total)
        n = self.hi - self.lo
        if positions is not None:
            x0 = torch.as_tensor(np.asarray(positions, dtype=np.float64)[self.lo:self.hi])
        else:
            g = torch.Generator(device="cpu")
            if self.N_total * self.P <= 1 << 26:          # one stream for the whole swarm: sharding-independent
                g.manual_seed(self.seed)
                u = torch.rand(self.N_total, self.P, generator=g, dtype=torch.float64)[self.lo:self.hi]
            else:
                g.manual_seed(self.seed * 1000003 + self.rank)
                u = torch.rand(n, self.P, generator=g, dtype=torch.float64)
            x0 = self.lo_b + (self.hi_b - self.lo_b) * u
        self.x = x0.to(self.dev).contiguous()
        self.v = torch.zeros_like(self.x)
        self.best = self.x.clone()
        self.best_fit = torch.full((n,), float("inf"), dtype=torch.float64, device=self.dev)
        self.weights = self.x.to(torch.float32).contiguous()
        S, P, dev = self.S, self.P, self.dev
        self.swarm_best = torch.zeros(S, P, dtype=torch.float64, device=dev)
        self.swarm_best_fit = torch.full((S,), float("inf"), dtype=torch.float64, device=dev)
        self.gbest_pos = torch.zeros(P, dtype=torch.float64, device=dev)
        self.gbest_fit = torch.full((1,), float("inf"), dtype=torch.float64, device=dev)
        self.sel_idx = torch.zeros(S, dtype=torch.int32, device=dev)
        self.improved = torch.zeros(S, dtype=torch.int32, device=dev)
        self.cand = torch.zeros(S, P, dtype=torch.float64, device=dev)
        self._cap = max(int(p.get('generations', 400)), 16)
        self.hist_gb = torch.full((self._cap,), float("nan"), dtype=torch.float64, device=dev)
        self.hist_stats = torch.zeros(self._cap, S + 1, 6, dtype=torch.float64, device=dev)
        self.capped = torch.zeros((), dtype=torch.int64, device=dev)
        self.last_steps = None

    # ------------------------------------------------------------------ layout
    def _layout(self, n_total):
        """Equal blocks of ceil(N / world) particles per rank: a particle's slot in the gathered
        fitness vector is its global index, the tail of the last blocks is padding (swarm id -1)."""
        torch = self.torch
        self.N_total = int(n_total)
        self.width = max(-(-self.N_total // self.world), 1)
        self.lo = min(self.rank * self.width, self.N_total)
        self.hi = min(self.lo + self.width, self.N_total)
        npad = self.world * self.width
        self.allfit = torch.full((npad,), float("inf"), dtype=torch.float64, device=self.dev)
        self._send = torch.full((self.width,), float("inf"), dtype=torch.float64, device=self.dev)
        sw = np.full(npad, -1, dtype=np.int32)
        for k, m in enumerate(self.members):
            sw[np.asarray(m, dtype=np.int64)] = k
        self.swarm_of_all = torch.as_tensor(sw).to(self.dev)

    def _ptr(self, t, offset_elems=0):
        return self.C.c_void_p(t.data_ptr() + offset_elems * t.element_size())

    def _stream(self):
        return self.C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _evaluate(self, weights, index0):
        """rollout + seed mean; returns the per-particle fitness (device) of `weights`."""
        fit, steps, tid = self.env.rollout_pso(weights, n_seeds=self.n_seeds, max_steps=self.max_steps,
                                               index0=index0, generation=self.generation)
        self.capped += (tid < 0).sum()
        self.last_steps = steps
        return fit

    # ------------------------------------------------------------------ one generation
    def step(self):
        """One generation; returns the gathered fitness vector (device tensor, padded slots = +inf)."""
        torch, N, lib, p = self.torch, self.N, self.lib, self.params
        g, n, st = self.generation, self.hi - self.lo, self._stream()
        if g >= self._cap:                                # grow the history buffers
            self.hist_gb = torch.cat([self.hist_gb, torch.full_like(self.hist_gb, float("nan"))])
            self.hist_stats = torch.cat([self.hist_stats, torch.zeros_like(self.hist_stats)])
            self._cap *= 2
        if n > 0:
            fit = self._evaluate(self.weights, self.lo)
            dst = self._send if self.world > 1 else self.allfit
            N.check(lib.pd_pso_seed_mean(self._ptr(fit), n, self.n_seeds,
                                         self._ptr(dst, 0 if self.world > 1 else self.lo), st))
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.allfit, self._send, group=self.group)
        N.check(lib.pd_pso_select(self._ptr(self.allfit), self._ptr(self.swarm_of_all), self.allfit.numel(), self.S,
                                  self._ptr(self.swarm_best_fit), self._ptr(self.sel_idx), self._ptr(self.improved),
                                  self._ptr(self.hist_stats, g * (self.S + 1) * 6), st))
        N.check(lib.pd_pso_gather(self._ptr(self.x), self.lo, n, self.P, self._ptr(self.sel_idx),
                                  self._ptr(self.improved), self.S, self._ptr(self.cand), st))
        if self.world > 1:          # exactly one rank holds each improved row, the others contribute zeros
            self.dist.all_reduce(self.cand, op=self.dist.ReduceOp.SUM, group=self.group)
        N.check(lib.pd_pso_apply(self._ptr(self.cand), self._ptr(self.improved), self.S, self.P,
                                 self._ptr(self.swarm_best), self._ptr(self.swarm_best_fit), self._ptr(self.gbest_pos),
                                 self._ptr(self.gbest_fit), self._ptr(self.hist_gb, g), st))
        w = p['w_start'] - (p['w_start'] - p['w_end']) * g / p['generations']
        if n > 0:
            N.check(lib.pd_pso_update(
                self._ptr(self.x), self._ptr(self.v), self._ptr(self.best), self._ptr(self.best_fit),
                self._ptr(self.allfit, self.lo), self._ptr(self.swarm_of_all, self.lo), self._ptr(self.swarm_best),
                self._ptr(self.weights), n, self.P, self.lo, float(w), float(p['c1']), float(p['c2']),
                self.lo_b, self.hi_b, self.seed, g, st))
        if g % p.get('communication_freq', 10) == 0 and g > 0:
            self.share_information()
        if g % p.get('migration_freq', 20) == 0 and g > 0:
            self.migrate_particles()
        if g == p.get('re_initialise_generation', 60):
            self.re_initialise_swarms()
        self.generation += 1
        return self.allfit

    def run(self, generations=None):
        for _ in range(generations if generations is not None else self.params['generations']):
            self.step()
        return self.global_best_position, self.global_best_fitness

    # ------------------------------------------------------------------ reference operators
    def share_information(self):
        """:523-542 on the device: the coin is a host-side Philox draw per sub-swarm, the best
        sub-swarm index stays a device scalar; every rank evaluates the (at most S) blended
        positions itself - they are replicated - so no communication is needed."""
        torch = self.torch
        coins = [self.streams.coin(i, self.generation) < 0.5 for i in range(self.S)]
        if not any(coins):
            return
        kbest = torch.argmin(self.swarm_best_fit)
        coin_t = torch.tensor(coins, device=self.dev)
        mask = coin_t & (torch.arange(self.S, device=self.dev) != kbest)
        blended = (1 - 0.3) * self.swarm_best + 0.3 * self.swarm_best.index_select(0, kbest.reshape(1))
        self.swarm_best.copy_(torch.where(mask[:, None], blended, self.swarm_best))
        fit = self._evaluate(self.swarm_best.to(torch.float32).contiguous(), self.N_total)
        new = torch.empty(self.S, dtype=torch.float64, device=self.dev)
        self.N.check(self.lib.pd_pso_seed_mean(self._ptr(fit), self.S, self.n_seeds, self._ptr(new), self._stream()))
        better = mask & (new < self.swarm_best_fit)
        self.swarm_best_fit.copy_(torch.where(better, new, self.swarm_best_fit))

    def migrate_particles(self):
        """:544-552 as index relabelling: host-side membership lists (the draw needs only their
        lengths) and one int32 store into the device membership array per migrant."""
        S = self.S
        for i in range(S):
            if len(self.members[i]) > 1:
                for m in range(self.params.get('number_of_migrants', 1)):
                    u1, u2 = self.streams.migrant(i, self.generation, m)
                    others = [k for k in range(S) if k != i]
                    nm = len(self.members[i])
                    pick = min(int(u1 * nm), nm - 1)
                    target = others[min(int(u2 * len(others)), len(others) - 1)]
                    j = self.members[i].pop(pick)
                    self.members[target].append(j)
                    self.swarm_of_all[j] = target

    def _gather_rows(self, local, rows):
        """local [n_local, ...] block-sharded -> the global rows `rows` on every rank."""
        torch = self.torch
        rows_t = torch.as_tensor(rows, device=self.dev)
        mine = (rows_t >= self.lo) & (rows_t < self.hi)
        src = (rows_t - self.lo).clamp(0, max(self.hi - self.lo - 1, 0))
        out = local.index_select(0, src) if self.hi > self.lo else \
            torch.zeros((len(rows),) + tuple(local.shape[1:]), dtype=local.dtype, device=self.dev)
        out = torch.where(mine.reshape([-1] + [1] * (local.dim() - 1)), out, torch.zeros_like(out))
        if self.world > 1:
            self.dist.all_reduce(out, op=self.dist.ReduceOp.SUM, group=self.group)
        return out

    def re_initialise_swarms(self):
        """:375-386 - each sub-swarm keeps its best `re_initialise_number_of_particles //
        num_sub_swarms` particles (stable sort by personal best) in that order; the kept rows are
        compacted and re-sharded.  The one step of a run that reads device data back."""
        torch = self.torch
        keep_n = self.params.get('re_initialise_number_of_particles', 500) // self.S
        bf = self._gather_rows(self.best_fit, np.arange(self.N_total)).cpu().numpy()
        rows, new_members = [], []
        for m in self.members:
            kept = sorted(m, key=lambda i: bf[i])[:keep_n]
            new_members.append(list(range(len(rows), len(rows) + len(kept))))
            rows.extend(kept)
        rows = np.asarray(rows, dtype=np.int64)
        full = {name: self._gather_rows(getattr(self, name), rows) for name in ("x", "v", "best", "best_fit")}
        self.members = new_members
        self._layout(len(rows))
        for name, t in full.items():
            setattr(self, name, t[self.lo:self.hi].contiguous())
        self.weights = self.x.to(torch.float32).contiguous()

    # ------------------------------------------------------------------ results
    @property
    def global_best_fitness(self):
        return float(self.gbest_fit[0])

    @property
    def global_best_position(self):
        return self.gbest_pos.clone()

    @property
    def global_best_fitness_array(self):
        return self.hist_gb[:self.generation].cpu().numpy().tolist()

    @property
    def capped_episodes(self):
        return int(self.capped)

    def save_metrics(self, base_save_dir=None):
        """The reference's metrics files (metrics/subswarm_<i>_metrics.csv, global_metrics.csv,
        fitness_history.csv, :717-744, 811-833) from the device-side history, written once."""
        d = f"{base_save_dir or self.base_save_dir}/metrics"
        gb = self.hist_gb[:self.generation].cpu().numpy()
        stats = self.hist_stats[:self.generation].cpu().numpy()
        for g in range(self.generation):
            write_generation_metrics(d, g, [tuple(stats[g, k]) for k in range(self.S)], gb[g], stats[g, self.S, 1])
        write_fitness_history(d, gb, stats[:, self.S, 1], [stats[:, k, 0] for k in range(self.S)],
                              [stats[:, k, 1] for k in range(self.S)])
        return d

    def save_results(self, base_save_dir=None):
        names = list(getattr(self.model, "mock_dictionary_of_opt_params", {}).keys()) or \
            [f"p_{j}" for j in range(self.P)]
        write_results_csv(base_save_dir or self.base_save_dir, names, self.gbest_pos.cpu().numpy(),
                          self.global_best_fitness)

    def swarms(self):
        """The reference's `swarm.pkl` layout (list of sub-swarms of particle dicts) on the host."""
        n = self.N_total
        rows = np.arange(n)
        x, v, b = [self._gather_rows(getattr(self, k), rows).cpu().numpy() for k in ("x", "v", "best")]
        bf = self._gather_rows(self.best_fit, rows).cpu().numpy()
        return [[{'position': x[i].copy(), 'velocity': v[i].copy(),
                  'best_position': b[i].copy() if np.isfinite(bf[i]) else None, 'best_fitness': float(bf[i])}
                 for i in m] for m in self.members]

    def load_swarms(self, file_path):
        """Resume from a `swarm.pkl` in the reference's layout (load_swarms, :651-688): positions,
        velocities and personal bests are re-sharded over the ranks; sub-swarm and global bests are
        rebuilt from the personal bests, as upstream does."""
        torch = self.torch
        with open(file_path, "rb") as f:
            swarms = pickle.load(f)
        parts = [p for sw in swarms for p in sw]
        P = self.P
        x = np.array([p["position"] for p in parts], dtype=np.float64).reshape(-1, P)
        v = np.array([p["velocity"] for p in parts], dtype=np.float64).reshape(-1, P)
        bf = np.array([p["best_fitness"] for p in parts], dtype=np.float64)
        b = np.array([p["best_position"] if p["best_position"] is not None else p["position"] for p in parts],
                     dtype=np.float64).reshape(-1, P)
        self.members, i = [], 0
        for sw in swarms:
            self.members.append(list(range(i, i + len(sw))))
            i += len(sw)
        self.S = len(swarms)
        self._layout(len(parts))
        sl = slice(self.lo, self.hi)
        self.x = torch.as_tensor(x[sl]).to(self.dev).contiguous()
        self.v = torch.as_tensor(v[sl]).to(self.dev).contiguous()
        self.best = torch.as_tensor(b[sl]).to(self.dev).contiguous()
        self.best_fit = torch.as_tensor(bf[sl]).to(self.dev).contiguous()
        self.weights = self.x.to(torch.float32).contiguous()
        sb, sbf = np.zeros((self.S, P)), np.full(self.S, np.inf)
        for k, m in enumerate(self.members):
            if len(m):
                j = m[int(np.argmin(bf[m]))]            # first occurrence, as the sequential scan
                if np.isfinite(bf[j]):
                    sb[k], sbf[k] = b[j], bf[j]
        self.swarm_best = torch.as_tensor(sb).to(self.dev).contiguous()
        self.swarm_best_fit = torch.as_tensor(sbf).to(self.dev).contiguous()
        k = int(np.argmin(sbf))
        self.gbest_fit = torch.as_tensor(sbf[k:k + 1].copy()).to(self.dev)
        self.gbest_pos = torch.as_tensor(sb[k].copy()).to(self.dev)
        S, dev = self.S, self.dev
        self.sel_idx = torch.zeros(S, dtype=torch.int32, device=dev)
        self.improved = torch.zeros(S, dtype=torch.int32, device=dev)
        self.cand = torch.zeros(S, P, dtype=torch.float64, device=dev)
        self.hist_stats = torch.zeros(self._cap, S + 1, 6, dtype=torch.float64, device=dev)

    def save(self, base_save_dir=None):
        d = f"{base_save_dir or self.base_save_dir}/saves"
        sw = self.swarms()
        if self.rank == 0:
            os.makedirs(d, exist_ok=True)
            with open(f"{d}/swarm.pkl", "wb") as f:
                pickle.dump(sw, f)
