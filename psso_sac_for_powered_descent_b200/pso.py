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


class ShardedEvaluator:
    """fitness = evaluate(positions) with the particle list block-sharded over the ranks of a
    torch.distributed process group.  `local_eval(np.ndarray[n_local, P]) -> np.ndarray[n_local]`
    is the per-rank evaluator (the CUDA rollout in production)."""

    def __init__(self, local_eval, group=None):
        self.local_eval = local_eval
        self.group = group
        try:
            import torch.distributed as dist
            self.dist = dist if dist.is_available() and dist.is_initialized() else None
        except Exception:
            self.dist = None
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0

    def _device(self):
        import torch
        if self.dist and self.dist.get_backend(self.group) == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def __call__(self, positions):
        import torch
        pos = np.asarray(positions, dtype=np.float64)
        n = len(pos)
        lo, hi = shard_bounds(n, self.world, self.rank)
        local = np.asarray(self.local_eval(pos[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
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


def evaluate_worker_function(args):
    """Same tuple as the reference worker; evaluates one particle on the GPU."""
    from .envs import pso_wrapped_env
    position, flight_phase, enable_wind, stochastic_wind, horiontal_wind_percentile = args
    model = pso_wrapped_env(flight_phase=flight_phase, enable_wind=enable_wind,
                            stochastic_wind=stochastic_wind,
                            horiontal_wind_percentile=horiontal_wind_percentile)
    return model.objective_function(position)


class ParticleSubswarmOptimisation:
    """Drop-in for the reference class; the swarm lives in flat arrays instead of a list of
    dicts (converted on save/load so `swarm.pkl` stays interchangeable).

    Extra keyword arguments: `model` (anything with .bounds and .evaluate(positions) or
    .objective_function), `pso_params` override, `seed`, `n_seeds` (wind seeds per particle;
    fitness = mean over seeds - the reference evaluates one stochastic draw), `base_save_dir`.
    """

    def __init__(self, flight_phase, save_interval=5, enable_wind=False, stochastic_wind=False,
                 horiontal_wind_percentile=50, load_swarms=False, use_multiprocessing=True,
                 num_processes=None, model=None, pso_params=None, seed=None, n_seeds=1,
                 precision="fp32", base_save_dir=None, evaluator=None):
        assert flight_phase in ['subsonic', 'supersonic', 'flip_over_boostbackburn',
                                'ballistic_arc_descent', 'landing_burn', 'landing_burn_pure_throttle']
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
        self.rng_py = random.Random(seed) if seed is not None else random
        self.rng_np = np.random.default_rng(seed)
        self.evaluator = evaluator or ShardedEvaluator(self._local_eval)
        self.w = self.w_start
        self.timestamp = datetime.now().strftime('%Y-%m-%d_%H-%M-%S')
        self.base_save_dir = base_save_dir or f'data/pso_saves/{flight_phase}/run_{self.timestamp}'
        self.initialize_swarms()
        if load_swarms:
            self.load_swarms()

    # ------------------------------------------------------------------ evaluation
    def _local_eval(self, positions):
        if hasattr(self.model, "evaluate"):
            out = self.model.evaluate(positions, n_seeds=self.n_seeds)
            fit = out[0].detach().cpu().numpy() if hasattr(out[0], "detach") else np.asarray(out[0])
            return fit.reshape(len(positions), self.n_seeds).mean(axis=1)
        return np.array([self.model.objective_function(p) for p in positions])

    def parallel_evaluate(self, positions):
        """list[np.ndarray(P)] -> list[float], order-preserving (reference :334-350)."""
        if len(positions) == 0:
            return []
        return [float(f) for f in self.evaluator(np.asarray(positions, dtype=np.float64))]

    # ------------------------------------------------------------------ swarm state
    def initialize_swarms(self):
        """Same draw order as the reference (:389-411): sub-swarm, particle, bound."""
        n_sub = self.pop_size // self.num_sub_swarms
        N, P = n_sub * self.num_sub_swarms, len(self.bounds)
        self.position = np.empty((N, P), dtype=np.float64)
        for i in range(N):
            for j, b in enumerate(self.bounds):
                self.position[i, j] = self.rng_py.uniform(b[0], b[1])
        self.velocity = np.zeros((N, P))
        self.best_position = np.full((N, P), np.nan)
        self.best_fitness = np.full(N, np.inf)
        self.swarm_of = np.repeat(np.arange(self.num_sub_swarms), n_sub)
        self.subswarm_best_positions = [None] * self.num_sub_swarms
        self.subswarm_best_fitnesses = [float('inf')] * self.num_sub_swarms
        self.global_best_position, self.global_best_fitness = None, float('inf')
        self.global_best_fitness_array, self.global_best_position_array = [], []
        self.average_particle_fitness_array = []
        self.subswarm_best_fitness_array = [[] for _ in range(self.num_sub_swarms)]
        self.subswarm_avg_array = [[] for _ in range(self.num_sub_swarms)]

    @property
    def swarms(self):
        """Reference layout: list[list[dict(position, velocity, best_position, best_fitness)]]."""
        out = [[] for _ in range(self.num_sub_swarms)]
        for i in range(len(self.position)):
            bp = None if np.isnan(self.best_position[i, 0]) else self.best_position[i].copy()
            out[self.swarm_of[i]].append({'position': self.position[i].copy(),
                                          'velocity': self.velocity[i].copy(),
                                          'best_position': bp, 'best_fitness': float(self.best_fitness[i])})
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
        self.swarm_of = np.array([k for k, _ in parts])

    def weight_linear_decrease(self, generation):
        return self.w_start - (self.w_start - self.w_end) * generation / self.generations

    # ------------------------------------------------------------------ one generation
    def step_generation(self, generation):
        fitness = np.asarray(self.parallel_evaluate(list(self.position)))
        improved = fitness < self.best_fitness
        self.best_fitness[improved] = fitness[improved]
        self.best_position[improved] = self.position[improved]
        for k in range(self.num_sub_swarms):
            idx = np.nonzero(self.swarm_of == k)[0]
            if len(idx) == 0:
                continue
            j = idx[np.argmin(fitness[idx])]
            if fitness[j] < self.subswarm_best_fitnesses[k]:
                self.subswarm_best_fitnesses[k] = float(fitness[j])
                self.subswarm_best_positions[k] = self.position[j].copy()
            self.subswarm_best_fitness_array[k].append(self.subswarm_best_fitnesses[k])
            self.subswarm_avg_array[k].append(float(np.mean(fitness[idx])))
        for k, f in enumerate(self.subswarm_best_fitnesses):
            if f < self.global_best_fitness:
                self.global_best_fitness = f
                self.global_best_position = self.subswarm_best_positions[k].copy()
        self.average_particle_fitness_array.append(float(np.mean(fitness)))
        self.w = self.weight_linear_decrease(generation)
        # velocity update with the *sub-swarm* best; one scalar r1, r2 per particle (:517-521)
        r = self.rng_np.random((len(self.position), 2))
        local_best = np.stack([self.subswarm_best_positions[k] for k in self.swarm_of])
        self.velocity = (self.w * self.velocity
                         + self.c1 * r[:, :1] * (self.best_position - self.position)
                         + self.c2 * r[:, 1:] * (local_best - self.position))
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
                    and getattr(self.evaluator, "rank", 0) == 0:
                self.save()
                self.save_results()
        return self.global_best_position, self.global_best_fitness

    __call__ = run

    def share_information(self):
        """:523-542 - blend the other sub-swarms' bests towards the best one (p = 0.5)."""
        influence_factor, sharing_probability = 0.3, 0.5
        best = int(np.argmin(self.subswarm_best_fitnesses))
        for i in range(self.num_sub_swarms):
            if i != best and self.rng_py.random() < sharing_probability:
                self.subswarm_best_positions[i] = ((1 - influence_factor) * self.subswarm_best_positions[i]
                                                   + influence_factor * self.subswarm_best_positions[best])
                new_fitness = self.parallel_evaluate([self.subswarm_best_positions[i]])[0]
                if new_fitness < self.subswarm_best_fitnesses[i]:
                    self.subswarm_best_fitnesses[i] = new_fitness

    def migrate_particles(self):
        """:544-552 - index relabelling instead of moving dicts between lists."""
        for i in range(self.num_sub_swarms):
            members = np.nonzero(self.swarm_of == i)[0]
            if len(members) > 1:
                for _ in range(self.number_of_migrants):
                    members = np.nonzero(self.swarm_of == i)[0]
                    j = members[self.rng_py.randrange(len(members))]
                    self.swarm_of[j] = self.rng_py.choice([k for k in range(self.num_sub_swarms) if k != i])

    def re_initialise_swarms(self):
        """:375-386 - keep the best `re_initialise_number_of_particles // num_sub_swarms` per swarm."""
        keep_n = self.re_initialise_number_of_particles // self.num_sub_swarms
        keep = []
        for k in range(self.num_sub_swarms):
            idx = np.nonzero(self.swarm_of == k)[0]
            keep.extend(idx[np.argsort(self.best_fitness[idx], kind="stable")[:keep_n]])
        keep = np.array(sorted(keep))
        for name in ("position", "velocity", "best_position", "best_fitness", "swarm_of"):
            setattr(self, name, getattr(self, name)[keep])

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
        for i in range(len(self.position)):
            k, f = self.swarm_of[i], self.best_fitness[i]
            if f < self.subswarm_best_fitnesses[k]:
                self.subswarm_best_fitnesses[k] = float(f)
                self.subswarm_best_positions[k] = self.best_position[i].copy()
            if f < self.global_best_fitness:
                self.global_best_fitness, self.global_best_position = float(f), self.best_position[i].copy()

    def save_results(self):
        """particle_subswarm_optimisation_results.csv with the reference's column names
        (`<layer>_weight_<j>` / `<layer>_bias_<j>`, env_wrapped_ea.py:61-75) so that
        src/particle_swarm_optimisation/network_loader.py keeps working."""
        import csv
        os.makedirs(self.base_save_dir, exist_ok=True)
        names = list(getattr(self.model, "mock_dictionary_of_opt_params",
                             {f"p_{j}": 0 for j in range(len(self.bounds))}).keys())
        with open(f'{self.base_save_dir}/particle_subswarm_optimisation_results.csv', 'w', newline='') as f:
            w = csv.writer(f)
            w.writerow(['Algorithm'] + names + ['Best Fitness'])
            w.writerow(['Particle Subswarm Optimisation'] + [repr(float(v)) for v in self.global_best_position]
                       + [repr(float(self.global_best_fitness))])
        with open(f'{self.base_save_dir}/pso_config.json', 'w') as f:
            json.dump({'flight_phase': self.flight_phase, 'pso_params': self.pso_params,
                       'enable_wind': self.enable_wind, 'stochastic_wind': self.stochastic_wind,
                       'horiontal_wind_percentile': self.horiontal_wind_percentile,
                       'n_seeds': self.n_seeds}, f, indent=1)


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
    """Device-resident sub-swarm PSO for large swarms (BASELINE config 5): positions,
    velocities and personal bests stay in HBM as fp64, the fitness evaluation is the persistent
    rollout kernel, the update is `pd_pso_update`, and the only per-generation communication is
    one all-gather of the fp64 fitness slice plus one broadcast of a sub-swarm best position
    (only in generations where that sub-swarm improved).

    Particles are block-sharded: rank r owns [lo, hi) = shard_bounds(N, world, r); sub-swarms are
    index ranges of the global swarm as in `initialize_swarms` (:389-411).  r1, r2 come from
    Philox keyed by the global particle index, so the optimisation trajectory is identical for
    every world size.
    """

    def __init__(self, model, n_particles, pso_params, n_seeds=1, seed=0, max_steps=4096, group=None):
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
        self.N_total, self.P = int(n_particles), len(model.bounds)
        self.S = p['num_sub_swarms']
        self.lo_b, self.hi_b = float(model.bounds[0][0]), float(model.bounds[0][1])
        try:
            import torch.distributed as dist
            self.dist = dist if dist.is_available() and dist.is_initialized() else None
        except Exception:
            self.dist = None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.lo, self.hi = shard_bounds(self.N_total, self.world, self.rank)
        n = self.hi - self.lo
        # initial positions: one uniform draw per (particle, parameter), keyed by global index
        g = torch.Generator(device="cpu")
        g.manual_seed(self.seed)
        full = torch.rand(self.N_total, self.P, generator=g, dtype=torch.float64) \
            if self.N_total * self.P <= 1 << 26 else None
        if full is not None:
            x0 = full[self.lo:self.hi]
        else:       # very large swarms: per-rank stream (still deterministic per world size)
            g.manual_seed(self.seed * 1000003 + self.rank)
            x0 = torch.rand(n, self.P, generator=g, dtype=torch.float64)
        self.x = (self.lo_b + (self.hi_b - self.lo_b) * x0).to(self.dev).contiguous()
        self.v = torch.zeros_like(self.x)
        self.best = self.x.clone()
        self.best_fit = torch.full((n,), float("inf"), dtype=torch.float64, device=self.dev)
        sub = self.N_total // self.S
        gidx = torch.arange(self.lo, self.hi, device=self.dev)
        self.swarm_of = torch.clamp(gidx // max(sub, 1), max=self.S - 1).to(torch.int32).contiguous()
        self.swarm_of_all = torch.clamp(torch.arange(self.N_total, device=self.dev) // max(sub, 1), max=self.S - 1)
        self.swarm_best = torch.zeros(self.S, self.P, dtype=torch.float64, device=self.dev)
        self.swarm_best_fit = torch.full((self.S,), float("inf"), dtype=torch.float64, device=self.dev)
        self.weights = self.x.to(torch.float32).contiguous()
        self.generation = 0
        self.global_best_fitness = float("inf")
        self.global_best_position = None
        width = -(-self.N_total // self.world)
        self._width = width
        self._send = torch.full((width,), float("inf"), dtype=torch.float64, device=self.dev)
        self._recv = torch.empty(self.world * width, dtype=torch.float64, device=self.dev)

    def _gather_fitness(self, local):
        torch = self.torch
        if self.world == 1:
            return local
        self._send.fill_(float("inf"))
        self._send[:local.numel()] = local
        self.dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        parts = [self._recv[r * self._width: r * self._width + (shard_bounds(self.N_total, self.world, r)[1]
                                                               - shard_bounds(self.N_total, self.world, r)[0])]
                 for r in range(self.world)]
        return torch.cat(parts)

    def step(self):
        """One generation; returns the full fitness vector (device tensor)."""
        torch, C = self.torch, self.C
        p = self.params
        fit, steps, tid = self.env.rollout_pso(self.weights, n_seeds=self.n_seeds, max_steps=self.max_steps)
        local = fit.reshape(-1, self.n_seeds).mean(dim=1)
        allfit = self._gather_fitness(local)
        # sub-swarm bests: arg-min per index range, position broadcast from the owner
        for k in range(self.S):
            mask = self.swarm_of_all == k
            masked = torch.where(mask, allfit, torch.full_like(allfit, float("inf")))
            j = int(torch.argmin(masked))
            fj = float(masked[j])
            if fj < float(self.swarm_best_fit[k]):
                self.swarm_best_fit[k] = fj
                row = self.swarm_best[k]
                if self.lo <= j < self.hi:
                    row.copy_(self.x[j - self.lo])
                if self.world > 1:
                    owner = next(r for r in range(self.world)
                                 if shard_bounds(self.N_total, self.world, r)[0] <= j
                                 < shard_bounds(self.N_total, self.world, r)[1])
                    src = self.dist.get_global_rank(self.group, owner) if self.group else owner
                    self.dist.broadcast(row, src=src, group=self.group)
        kbest = int(torch.argmin(self.swarm_best_fit))
        if float(self.swarm_best_fit[kbest]) < self.global_best_fitness:
            self.global_best_fitness = float(self.swarm_best_fit[kbest])
            self.global_best_position = self.swarm_best[kbest].clone()
        w = p['w_start'] - (p['w_start'] - p['w_end']) * self.generation / p['generations']
        ptr = lambda t: C.c_void_p(t.data_ptr())
        self.N.check(self.lib.pd_pso_update(
            ptr(self.x), ptr(self.v), ptr(self.best), ptr(self.best_fit), ptr(local.contiguous()),
            ptr(self.swarm_of), ptr(self.swarm_best), ptr(self.weights), self.hi - self.lo, self.P,
            self.lo, float(w), float(p['c1']), float(p['c2']), self.lo_b, self.hi_b, self.seed,
            self.generation, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.generation += 1
        self.last_steps = steps
        return allfit
