"""Host-side mirror of the reference's env interfaces on top of the CUDA library.

Batched API
    BatchedRocketEnv          n_envs copies of rocket_environment_pre_wrap
                              (src/envs/base_environment.py:12-154) living on one GPU.

Drop-in, scalar-compatible mirrors (same constructor kwargs - including the misspelt
`horiontal_wind_percentile` - same method names, return shapes and error behaviour):
    rocket_environment_pre_wrap   src/envs/base_environment.py
    pso_wrapped_env               src/envs/pso/env_wrapped_ea.py:137-230
    rl_wrapped_env_pytorch        src/envs/rl/env_wrapped_rl_pytorch.py:68-205

All arithmetic of reset/step/objective_function happens in the kernels behind
include/pd_b200.h; torch is used for device memory and streams only.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _native as N
from .params import RocketParams

# landing_burn_pure_throttle / landing_burn: type 'pso' and 'rl'.  subsonic, supersonic,
# ballistic_arc_descent, landing_burn_pure_throttle_Pcontrol: type 'rl' only - upstream their pso
# closures have the wrong arity (rtd_pso.py:38-157 vs base_environment.py:150-152 -> TypeError).
# flip_over_boostbackburn: type 'supervisory' only - its rl and pso truncated_func take one argument
# (rtd_rl.py:132, rtd_pso.py:107 -> TypeError in step), the supervisory closures
# (rtd_supervisory_mock.py:34-38, 57-61) and the classical controller run it.
# landing_burn_ACS (broken in compile_physics, rockets_physics.py:867-889) does not run upstream and is
# not offered.
WORKING_PHASES = tuple(N.PHASES)
ALL_PHASES = ["subsonic", "supersonic", "flip_over_boostbackburn", "ballistic_arc_descent",
              "landing_burn", "landing_burn_ACS", "landing_burn_pure_throttle",
              "landing_burn_pure_throttle_Pcontrol"]


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class BatchedRocketEnv:
    """A batch of landing-burn environments resident on one GPU.

    type / flight_phase / enable_wind / stochastic_wind / horiontal_wind_percentile /
    trajectory_length / discount_factor have the reference's meaning
    (base_environment.py:12-20).  precision = 'fp64' (parity build) | 'fp32' (production).
    """

    def __init__(self, n_envs, type="pso", flight_phase="landing_burn_pure_throttle",
                 enable_wind=False, stochastic_wind=False, horiontal_wind_percentile=50,
                 trajectory_length=1, discount_factor=0.99, precision="fp32", auto_reset=False,
                 device=None, seed=0, params: RocketParams | None = None, raw_actions=False,
                 exact_aero=None):
        """raw_actions (type 'rl'): False = `step` takes the policy's action and applies
        rl_wrapped_env_pytorch.augment_action inside the kernel (landing_burn log-compression,
        P-control reference-speed scaling); True = actions are what
        rocket_environment_pre_wrap.step expects."""
        assert flight_phase in ALL_PHASES
        if flight_phase not in WORKING_PHASES:
            raise NotImplementedError(
                f"flight phase {flight_phase!r} does not run in the reference either "
                "(landing_burn_ACS: rockets_physics.py:867-889); see DESIGN.md")
        assert type in ("rl", "pso", "supervisory")
        if flight_phase in N.SUPERVISORY_ONLY_PHASES and type != "supervisory":
            raise TypeError(f"{flight_phase}: only type='supervisory' works upstream (the rl and pso "
                            "truncated_func of this phase take one argument, rtd_rl.py:132 / rtd_pso.py:107)")
        if flight_phase in N.RL_ONLY_PHASES and type == "pso":
            raise TypeError(f"{flight_phase}: only type='rl' works upstream (the pso closures of this "
                            "phase have the wrong arity, rtd_pso.py:38-157)")
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedRocketEnv needs a CUDA device (no CPU fallback)")
        if enable_wind:
            assert 50 <= horiontal_wind_percentile <= 99, \
                "Given percentile must be between 50 and 99"
        self.lib = N.load_library()
        self.params = params or RocketParams.default()
        self.flight_phase, self.type = flight_phase, type
        self.phase_id = N.PHASES[flight_phase]
        self.n_envs = int(n_envs)
        self.precision = precision
        self.dtype = torch.float64 if precision == "fp64" else torch.float32
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else device.index or 0)
        self.obs_dim, self.act_dim = N.OBS_DIM[self.phase_id], N.ACT_DIM[self.phase_id]
        self.n_actor_params = N.N_PARAMS.get(self.phase_id)
        self.dt = 0.1
        self.enable_wind = bool(enable_wind)
        cfg = N.PdConfig()
        cfg.phase, cfg.rtd, cfg.precision = self.phase_id, N.RTD[type], N.PRECISION[precision]
        cfg.enable_wind, cfg.stochastic_wind = int(bool(enable_wind)), int(bool(stochastic_wind))
        cfg.auto_reset, cfg.n_envs, cfg.device = int(bool(auto_reset)), self.n_envs, self.device.index
        cfg.seed = int(seed)
        g, L = discount_factor, trajectory_length
        cfg.rl_reward_scale = (1 - g) / (1 - g ** L) if (g is not None and L) else 1.0
        cfg.discount_factor = g if g is not None else 0.99
        cfg.raw_actions = int(bool(raw_actions))
        # fp32 build: C_L / C_D from the bicubic patches of the thin-plate sums (csrc/pd_patch.h);
        # exact_aero=True (or PD_EXACT_AERO=1) keeps the 50-term sums, as the fp64 build always does
        if exact_aero is None:
            exact_aero = os.environ.get("PD_EXACT_AERO", "0") == "1"
        cfg.exact_aero = int(bool(exact_aero))
        self._cparams, self._keep = N.make_params(self.params, horiontal_wind_percentile)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self.lib.pd_create(C.byref(cfg), C.byref(self._cparams), C.byref(self._h)))
        B, dev = self.n_envs, self.device
        # step outputs live in ONE device allocation so that a host caller needs a single
        # device->host copy per step: [obs | reward | trunc_id | done | truncated]
        esz = 8 if precision == "fp64" else 4
        self._out_layout = []
        off = 0
        for name, nbytes in (("obs", B * self.obs_dim * esz), ("reward", B * esz),
                             ("done", B), ("truncated", B), ("trunc_id", B * 4)):
            self._out_layout.append((name, off, nbytes))
            off += (nbytes + 15) // 16 * 16
        self._out = torch.empty(off, dtype=torch.uint8, device=dev)

        def view(buf, name, dt, shape):
            _, o, nb = next(x for x in self._out_layout if x[0] == name)
            return buf[o:o + nb].view(dt).reshape(shape)
        self._view = view
        self.obs = view(self._out, "obs", self.dtype, (B, self.obs_dim))
        self.reward = view(self._out, "reward", self.dtype, (B,))
        self.trunc_id = view(self._out, "trunc_id", torch.int32, (B,))
        self.done = view(self._out, "done", torch.uint8, (B,))
        self.truncated = view(self._out, "truncated", torch.uint8, (B,))
        self.next_obs = torch.empty(B, self.obs_dim, dtype=self.dtype, device=dev)
        self._tape = self._sigma = None
        self._host = None
        self._dirty = True      # device-side calls since the last step_host (stream ordering)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.pd_destroy(self._h)
                self._h = None
        except Exception:
            pass

    close = __del__

    # ------------------------------------------------------------------ reset / step
    def reset(self, mask: torch.Tensor | None = None):
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        N.check(self.lib.pd_reset(self._h, _ptr(mask), _stream()))
        self._dirty = True
        return self.get_state()

    def step(self, actions: torch.Tensor, dbg: torch.Tensor | None = None, _out=None):
        """actions: cuda tensor [n_envs, A] float64 or float32 (its dtype selects the
        reference's pure-fp64 or float32-contaminated arithmetic).  Returns the handle's
        (obs, reward, done, truncated, trunc_id) tensors, overwritten in place every step."""
        if actions.dtype not in (torch.float64, torch.float32):
            raise TypeError("actions must be float64 or float32")
        a = actions.reshape(self.n_envs, self.act_dim).contiguous()
        if a.device != self.device and not a.is_pinned():
            raise ValueError("actions must live on the env's device")
        obs, reward, done, truncated = _out or (self.obs, self.reward, self.done, self.truncated)
        self._dirty = True
        N.check(self.lib.pd_step(self._h, _ptr(a), 1 if a.dtype == torch.float32 else 0,
                                 _ptr(obs), _ptr(reward), _ptr(done),
                                 _ptr(truncated), _ptr(self.trunc_id), _ptr(self.next_obs),
                                 _ptr(dbg), _stream()))
        return obs, reward, done, truncated, self.trunc_id

    def step_host(self, actions):
        """Host-facing step: `actions` is a float32/float64 numpy array (or CPU tensor)
        [n_envs, A]; returns numpy views (obs, reward, done, truncated, trunc_id) of a pinned
        host buffer, valid until the next call (trunc_id is only refreshed by
        `self.trunc_id.cpu()`).  The step kernel reads the actions from, and stores its results to,
        mapped pinned host memory (the same bytes cross PCIe, without staging copies or extra
        launches, one ctypes call with cached pointers: 90 us against 125 us per 65 536-env step on
        B200); a CPU tensor that already
        lives in pinned memory is read where it is.  PD_HOST_STEP=copy selects explicit copies
        (H2D copy, then kernel + one D2H copy replayed from a CUDA graph), PD_HOST_STEP=zc_out
        keeps the H2D copy only."""
        a = actions if isinstance(actions, torch.Tensor) else torch.as_tensor(actions)
        if a.dtype not in (torch.float32, torch.float64):
            a = a.to(torch.float64)
        a = a.reshape(self.n_envs, self.act_dim)
        if self._host is None or self._host["dtype"] != a.dtype:
            h = dict(dtype=a.dtype)
            h["act_pin"] = torch.empty(self.n_envs, self.act_dim, dtype=a.dtype).pin_memory()
            h["act_dev"] = torch.empty(self.n_envs, self.act_dim, dtype=a.dtype, device=self.device)
            # trunc_id (diagnostic) is last in the layout and stays on the device unless asked for
            n_copy = next(o for nm, o, nb in self._out_layout if nm == "trunc_id")
            h["n_copy"] = n_copy
            h["out_pin"] = torch.empty(self._out.numel(), dtype=torch.uint8).pin_memory()
            h["stream"] = torch.cuda.Stream(device=self.device)
            v = lambda n, dt, sh: self._view(h["out_pin"], n, dt, sh).numpy()
            B = self.n_envs
            h["views"] = (v("obs", self.dtype, (B, self.obs_dim)), v("reward", self.dtype, (B,)),
                          v("done", torch.uint8, (B,)), v("truncated", torch.uint8, (B,)),
                          v("trunc_id", torch.int32, (B,)))
            B = self.n_envs
            mode = os.environ.get("PD_HOST_STEP", "zc_all")
            assert mode in ("zc_all", "zc_out", "copy")
            pv = lambda n, dt, sh: self._view(h["out_pin"], n, dt, sh)
            pinned_out = (pv("obs", self.dtype, (B, self.obs_dim)), pv("reward", self.dtype, (B,)),
                          pv("done", torch.uint8, (B,)), pv("truncated", torch.uint8, (B,)))
            h["mode"], h["pinned_out"] = mode, pinned_out
            h["act_code"] = 1 if a.dtype == torch.float32 else 0
            h["ptrs"] = tuple(_ptr(t) for t in pinned_out) + (_ptr(self.trunc_id), _ptr(self.next_obs), None,
                                                              h["stream"].cuda_stream)
            with torch.cuda.stream(h["stream"]) if mode != "zc_all" else contextlib.nullcontext():
                if mode == "zc_all":
                    self._host = h
                    return self.step_host(actions)

                def body():
                    if mode == "copy":
                        self.step(h["act_dev"])
                        h["out_pin"][:n_copy].copy_(self._out[:n_copy], non_blocking=True)
                    else:       # the kernel stores its results straight into mapped pinned memory
                        self.step(h["act_dev"], _out=pinned_out)
                h["act_dev"].zero_()
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=h["stream"]):
                    body()
                h["graph"] = g
            self._host = h
        h = self._host
        if self._dirty:
            # reset() / step() / set_state() / collect() run on torch's current stream and return
            # without synchronising; the host step runs on the handle's own stream: order it after them
            h["stream"].wait_stream(torch.cuda.current_stream())
        if h["mode"] == "zc_all":
            # lean path: one ctypes call with cached pointers on the handle's own stream
            if not (a.is_pinned() and a.is_contiguous()):
                h["act_pin"].copy_(a)
                a = h["act_pin"]
            if torch.cuda.current_device() != self.device.index:
                with torch.cuda.device(self.device):
                    N.check(self.lib.pd_step(self._h, a.data_ptr(), h["act_code"], *h["ptrs"]))
            else:
                N.check(self.lib.pd_step(self._h, a.data_ptr(), h["act_code"], *h["ptrs"]))
            h["stream"].synchronize()
            self._dirty = False
            return h["views"]
        if a.is_pinned() and a.is_contiguous():
            src = a
        else:
            h["act_pin"].copy_(a)
            src = h["act_pin"]
        with torch.cuda.stream(h["stream"]):
            h["act_dev"].copy_(src, non_blocking=True)
            h["graph"].replay()
        h["stream"].synchronize()
        self._dirty = False
        return h["views"]

    # ------------------------------------------------------------------ state access
    def get_state(self, full=False):
        B = self.n_envs
        st = torch.empty(B, 11, dtype=torch.float64, device=self.device)
        if not full:
            N.check(self.lib.pd_get_state(self._h, _ptr(st), None, None, None, _stream()))
            return st
        gw = torch.empty(B, 10, dtype=torch.float64, device=self.device)
        nw = torch.empty(B, dtype=torch.int32, device=self.device)
        ap = torch.empty(B, 3, dtype=torch.float64, device=self.device)
        N.check(self.lib.pd_get_state(self._h, _ptr(st), _ptr(gw), _ptr(nw), _ptr(ap), _stream()))
        return st, gw, nw, ap

    def set_state(self, state, g_window=None, n_window=None, act_prev=None):
        def prep(x, dt, shape):
            if x is None:
                return None
            return torch.as_tensor(x, dtype=dt).to(self.device).reshape(shape).contiguous()
        B = self.n_envs
        st = prep(state, torch.float64, (B, 11))
        gw = prep(g_window, torch.float64, (B, 10))
        nw = prep(n_window, torch.int32, (B,))
        ap = prep(act_prev, torch.float64, (B, 3))
        N.check(self.lib.pd_set_state(self._h, _ptr(st), _ptr(gw), _ptr(nw), _ptr(ap), _stream()))
        self._dirty = True
        torch.cuda.current_stream().synchronize()

    def set_wind_tape(self, tape, sigma_uv):
        """Parity hook: N(0,1) tape [n_envs, T] and (sigma_u, sigma_v) [n_envs, 2]."""
        if tape is None:
            self._tape = self._sigma = None
            N.check(self.lib.pd_set_wind_tape(self._h, None, 0, None))
            return
        self._tape = torch.as_tensor(tape, dtype=torch.float64).to(self.device).reshape(self.n_envs, -1).contiguous()
        self._sigma = torch.as_tensor(sigma_uv, dtype=torch.float64).to(self.device).reshape(-1, 2).contiguous()
        N.check(self.lib.pd_set_wind_tape(self._h, _ptr(self._tape), self._tape.shape[1], _ptr(self._sigma)))

    def check_status(self):
        st = C.c_int32(0)
        N.check(self.lib.pd_check_status(self._h, C.byref(st)))

    # ------------------------------------------------------------------ rollouts
    def rollout_pso(self, weights: torch.Tensor, n_seeds=1, max_steps=4096, terminal=False,
                    trace=False, index0=0, generation=0):
        """weights: cuda float32 [n_particles, n_params].  Returns fitness (float64
        [n_particles*n_seeds]), steps, trunc_id (, terminal_state).  trace=True returns a dict
        with the per-step states / actions / rewards as well.  index0 / generation select the
        gust-noise stream (pd_set_rollout_stream): global index of the first particle of this
        call and the PSO generation.  An episode that reaches max_steps has trunc_id -1 and is
        scored as a truncation at its final state."""
        w = weights.to(device=self.device, dtype=torch.float32).contiguous()
        N.check(self.lib.pd_set_rollout_stream(self._h, int(index0), int(generation) & 0xFFFFFFFF))
        n, p = w.shape
        E = n * n_seeds
        fit = torch.empty(E, dtype=torch.float64, device=self.device)
        steps = torch.empty(E, dtype=torch.int32, device=self.device)
        tid = torch.empty(E, dtype=torch.int32, device=self.device)
        term = torch.empty(E, 11, dtype=torch.float64, device=self.device) if terminal else None
        traj = acts = rews = None
        if trace:
            term = torch.empty(E, 11, dtype=torch.float64, device=self.device)
            traj = torch.zeros(max_steps, E, 11, dtype=torch.float64, device=self.device)
            acts = torch.zeros(max_steps, E, self.act_dim, dtype=torch.float32, device=self.device)
            rews = torch.zeros(max_steps, E, dtype=torch.float64, device=self.device)
        N.check(self.lib.pd_rollout_pso(self._h, _ptr(w), n, p, n_seeds, max_steps, _ptr(fit),
                                        _ptr(steps), _ptr(tid), _ptr(term), _ptr(traj), _ptr(acts),
                                        _ptr(rews), _stream()))
        if trace:
            return dict(fitness=fit, steps=steps, trunc_id=tid, terminal=term, traj=traj,
                        actions=acts, rewards=rews)
        return (fit, steps, tid, term) if terminal else (fit, steps, tid)

    def aero_patch_stats(self):
        """fp32 build: how many bicubic aero patches were built / rejected (csrc/pd_patch.h)."""
        counts = (C.c_int64 * 4)()
        err = C.c_double(0.0)
        N.check(self.lib.pd_aero_patch_stats(self._h, counts, C.byref(err)))
        return {"cd_patches": counts[0], "cd_rejected": counts[1], "cl_patches": counts[2],
                "cl_rejected": counts[3], "max_abs_error_in_use": err.value}

    def rollout_tape(self, actions: torch.Tensor, record=False):
        """actions [T, n_episodes, A] (float64 or float32): an env.reset() + env.step loop per
        episode, stopping at done/truncated."""
        T, E = actions.shape[0], actions.shape[1]
        a = actions.to(self.device).contiguous()
        ret = torch.empty(E, dtype=torch.float64, device=self.device)
        steps = torch.empty(E, dtype=torch.int32, device=self.device)
        tid = torch.empty(E, dtype=torch.int32, device=self.device)
        term = torch.empty(E, 11, dtype=torch.float64, device=self.device)
        traj = torch.zeros(T, E, 11, dtype=torch.float64, device=self.device) if record else None
        rew = torch.zeros(T, E, dtype=torch.float64, device=self.device) if record else None
        N.check(self.lib.pd_rollout_policy(self._h, 1, _ptr(a), 1 if a.dtype == torch.float32 else 0,
                                           E, T, _ptr(ret), _ptr(steps), _ptr(tid), _ptr(term),
                                           _ptr(traj), _ptr(rew), _stream()))
        return dict(ret=ret, steps=steps, trunc_id=tid, terminal=term, traj=traj, rewards=rew)

    def rollout_classical(self, n_episodes=1, max_steps=50000, record=False):
        """LandingBurn(test_case='control').run_closed_loop()
        (src/classical_controls/landing_burn_pure_throttle.py:332-339)."""
        E, T = n_episodes, max_steps
        steps = torch.empty(E, dtype=torch.int32, device=self.device)
        term = torch.empty(E, 11, dtype=torch.float64, device=self.device)
        traj = torch.zeros(T, E, 11, dtype=torch.float64, device=self.device) if record else None
        N.check(self.lib.pd_rollout_policy(self._h, 2, None, 0, E, T, None, _ptr(steps), None,
                                           _ptr(term), _ptr(traj), None, _stream()))
        return dict(steps=steps, terminal=term, traj=traj)


    # ------------------------------------------------------------------ shared SAC actor
    def _actor_struct(self, actor, deterministic, seed, fp32_path):
        """actor: src/agents/sac_pytorch.Actor-like module (layers Sequential of Linear+ReLU,
        mean, log_std, max_action) or a dict with w1,b1,w2,b2,wm,bm,ws,bs."""
        if isinstance(actor, dict):
            t = actor
        else:
            lin = [m for m in actor.layers if isinstance(m, torch.nn.Linear)]
            if len(lin) != 2:
                raise NotImplementedError("shared-actor kernel supports number_of_hidden_layers = 2")
            t = dict(w1=lin[0].weight, b1=lin[0].bias, w2=lin[1].weight, b2=lin[1].bias,
                     wm=actor.mean.weight, bm=actor.mean.bias, ws=actor.log_std.weight, bs=actor.log_std.bias)
            t["max_action"] = float(getattr(actor, "max_action", 1.0))
        keep = {k: v.detach().to(device=self.device, dtype=torch.float32).contiguous()
                for k, v in t.items() if k != "max_action"}
        a = N.PdSharedActor()
        a.hidden, a.deterministic = keep["w1"].shape[0], int(bool(deterministic))
        a.max_action, a.fp32_path, a.seed = float(t.get("max_action", 1.0)), int(bool(fp32_path)), int(seed)
        for k in ("w1", "b1", "w2", "b2", "wm", "bm", "ws", "bs"):
            setattr(a, k, keep[k].data_ptr())
        return a, keep

    def actor_forward(self, actor, obs, deterministic=True, seed=0, fp32_path=False, want_mean=False):
        a, keep = self._actor_struct(actor, deterministic, seed, fp32_path)
        obs = obs.to(device=self.device, dtype=torch.float32).contiguous()
        n = obs.shape[0]
        act = torch.empty(n, self.act_dim, dtype=torch.float32, device=self.device)
        mean = torch.empty(n, self.act_dim, dtype=torch.float32, device=self.device) if want_mean else None
        N.check(self.lib.pd_actor_forward(self._h, C.byref(a), _ptr(obs), n, _ptr(act), _ptr(mean), _stream()))
        return (act, mean) if want_mean else act

    def collect(self, actor, n_steps, deterministic=False, seed=0, fp32_path=False, next_obs=True,
                into=None):
        """n_steps of [shared-actor inference -> fused env step with auto-reset] on the whole
        batch (the loop body of sac_pytorch_powered_descent.py:160-183).  Returns step-major
        tensors obs, actions, rewards, done, truncated (, next_obs).  `into`: a
        replay.DeviceReplayBuffer - the kernels then write the transitions straight into its
        storage and the returned tensors are views of it."""
        a, keep = self._actor_struct(actor, deterministic, seed, fp32_path)
        B, T, dev = self.n_envs, n_steps, self.device
        if into is not None:
            if (into.state_dim, into.action_dim) != (self.obs_dim, self.act_dim) or into.device != dev:
                raise ValueError("replay buffer does not match this env's dims / device")
            start, v = into.reserve(T * B)
            out = dict(obs=v["obs"].view(T, B, self.obs_dim), actions=v["actions"].view(T, B, self.act_dim),
                       rewards=v["rewards"].view(T, B), done=v["done"].view(T, B),
                       truncated=torch.empty(T, B, dtype=torch.uint8, device=dev),
                       next_obs=v["next_obs"].view(T, B, self.obs_dim))
        else:
            out = dict(obs=torch.empty(T, B, self.obs_dim, dtype=torch.float32, device=dev),
                       actions=torch.empty(T, B, self.act_dim, dtype=torch.float32, device=dev),
                       rewards=torch.empty(T, B, dtype=torch.float32, device=dev),
                       done=torch.empty(T, B, dtype=torch.uint8, device=dev),
                       truncated=torch.empty(T, B, dtype=torch.uint8, device=dev))
            if next_obs:
                out["next_obs"] = torch.empty(T, B, self.obs_dim, dtype=torch.float32, device=dev)
        self._dirty = True
        N.check(self.lib.pd_collect_shared_actor(self._h, C.byref(a), T, _ptr(out["obs"]), _ptr(out["actions"]),
                                                 _ptr(out["rewards"]), _ptr(out["done"]), _ptr(out["truncated"]),
                                                 _ptr(out.get("next_obs")), _stream()))
        if into is not None:
            into.commit(start, T * B)
        return out


# =======================================================================================
# scalar-compatible drop-ins
# =======================================================================================
def _action_tensor(actions, act_dim, device):
    """Accept what the reference's decomposers accept (tuple, list, 1-D / 2-D ndarray, torch
    tensor; rockets_physics.py:199-219, 360-370) and keep its dtype semantics: float32
    ndarrays stay float32, everything else is float64."""
    if isinstance(actions, torch.Tensor):
        a = actions.detach()
        a = a.to(torch.float32 if a.dtype == torch.float32 else torch.float64)
        return a.reshape(1, act_dim).to(device)
    if isinstance(actions, np.ndarray):
        dt = torch.float32 if actions.dtype == np.float32 else torch.float64
        return torch.as_tensor(np.ascontiguousarray(actions).reshape(1, act_dim), dtype=dt).to(device)
    if isinstance(actions, (tuple, list)):
        return torch.tensor([float(v) for v in actions], dtype=torch.float64).reshape(1, act_dim).to(device)
    return torch.tensor([[float(actions)]], dtype=torch.float64).to(device)


class rocket_environment_pre_wrap:
    """One reference env (base_environment.py:12-154) backed by a 1-env GPU batch."""

    def __init__(self, type="rl", flight_phase="subsonic", enable_wind=True, stochastic_wind=True,
                 horiontal_wind_percentile=50, trajectory_length=100, discount_factor=0.99,
                 precision="fp64", seed=0, _wrapped=False):
        assert flight_phase in ALL_PHASES
        assert type in ["rl", "pso", "supervisory"]
        self.flight_phase, self.type, self.dt = flight_phase, type, 0.1
        self.enable_wind = enable_wind
        # the base env takes its actions as they are; the RL wrapper (_wrapped) hands over the
        # policy's action and lets the kernel apply augment_action
        self._b = BatchedRocketEnv(1, type, flight_phase, enable_wind, stochastic_wind,
                                   horiontal_wind_percentile, trajectory_length, discount_factor,
                                   precision=precision, seed=seed, raw_actions=not _wrapped)
        op = self._b.params.other_phases.get("initial_states", {}) if self._b.params.other_phases else {}
        self.state_initial = list(op.get(flight_phase, self._b.params.initial_state))
        self.wind_generator = self._b if enable_wind else None
        # the complete `info` dict of the reference (rockets_physics.py:649-702) needs the fp64
        # diagnostic kernel (no wind); otherwise the 16-value subset
        self._full_info = precision == "fp64" and not enable_wind
        if self._full_info:
            N.check(self._b.lib.pd_set_info_mode(self._b._h, 1))
        self._dbg = torch.zeros(1, 48 if self._full_info else 16, dtype=torch.float64,
                                device=self._b.device)
        self.truncation_id = 0
        self.reset()

    def reset(self):
        self.state = self._b.reset()[0].tolist()
        self.previous_state = self.state
        self.truncation_id = 0
        return self.state

    def step(self, actions):
        a = _action_tensor(actions, self._b.act_dim, self._b.device)
        obs, rew, done, trunc, tid = self._b.step(a, dbg=self._dbg)
        self.previous_state = self.state
        self.state = self._b.get_state()[0].tolist()
        d = self._dbg[0].tolist()
        self.truncation_id = int(tid[0])
        info = dict(mach_number=d[0], dynamic_pressure=d[1], CL=d[2], CD=d[3], air_density=d[4],
                    atmospheric_pressure=d[5], speed_of_sound=d[6], x_cog=d[7], inertia=d[8],
                    mass_flow=d[9], action_info=dict(throttle=d[10]), alpha_effective=d[11],
                    g_load_1_sec_window=d[12], ug=d[13], vg=d[14], state=self.state, actions=actions)
        if self._full_info:
            info.update(_full_info_dict(self.flight_phase, d, self.state, self.previous_state, actions,
                                        self._b.params))
        self._obs = obs[0]
        return self.state, float(rew[0]), bool(done[0]), bool(trunc[0]), info


def _full_info_dict(phase, d, state, prev_state, actions, p):
    """The rest of the reference's `info` dict (rockets_physics.py:649-702) from the primitives
    the diagnostic kernel exports (include/pd_b200.h: pd_set_info_mode)."""
    x = d[16:]
    (drag, lift, d_cp_cg, d_thrust_cg, fuel, c_par, c_perp, c_x, c_y, aero_x, aero_y, g, c_mz, aero_mz,
     mz, tdd, vx_dot, vy_dot, f_wind_x, gimbal_deg, dl_cmd, dr_cmd, mach_max) = x[:23]
    x = list(x) + [0.0] * 8
    gamma, mass = state[6], state[8]
    acc = {
        "acceleration_x_component_control": c_x / mass,
        "acceleration_y_component_control": c_y / mass,
        "acceleration_x_component_drag": -drag * math.cos(gamma) / mass,
        "acceleration_y_component_drag": -drag / mass * math.sin(gamma) / mass,     # sic, :654
        "acceleration_x_component_lift": -lift * math.cos(math.pi - gamma) / mass,
        "acceleration_y_component_lift": lift * math.sin(math.pi - gamma) / mass,
        "acceleration_x_component_gravity": 0,
        "acceleration_y_component_gravity": -g,
        "acceleration_x_component": vx_dot,
        "acceleration_y_component": vy_dot,
        "acceleration_x_component_wind": f_wind_x / mass,
        "acceleration_y_component_wind": 0.0,
    }
    mom = {"control_moment_z": c_mz, "aero_moment_z": aero_mz, "moments_z": mz, "theta_dot_dot": tdd,
           "M_wind_z": 0.0}
    throttle = d[10]
    acs_info = None
    if phase in ("landing_burn", "landing_burn_pure_throttle", "landing_burn_pure_throttle_Pcontrol"):
        # acs_model.py:39-84 from the exported C_a, C_n_alpha, filtered deflections and pitch angle
        theta0, Ca, cna, dl, dr = x[23:28]
        a_eff, q, x_cog = d[11], d[1], d[7]
        qS = q * p.grid_fin_area
        al, ar = a_eff - dl, a_eff - dr
        Cn_L, Cn_R = cna * math.degrees(al), cna * math.degrees(ar)
        f_perp = qS * (Cn_R * math.cos(dr) - Cn_L * math.cos(dl) - Ca * (math.sin(dl) - math.sin(dr)))
        f_par = qS * (Ca * (2 + math.cos(dl) + math.cos(dr)) - Cn_L * math.sin(dl) + Cn_R * math.sin(dr))
        m_z = -(p.d_base_grid_fin - x_cog) * f_perp + p.rocket_radius * qS * (
            Ca * (math.sin(dr) - math.sin(dl)) - Cn_L * math.cos(dl) + Cn_R * math.cos(dr))
        acs_info = {
            "alpha_local_left_rad": al, "alpha_local_right_rad": ar, "C_n_L": Cn_L, "C_a_L": Ca,
            "C_n_R": Cn_R, "C_a_R": Ca, "F_n_L": Cn_L * qS, "F_a_L": Ca * qS, "F_n_R": Cn_R * qS,
            "F_a_R": Ca * qS,
            "F_perpendicular_L": qS * (Cn_L * math.cos(dl) - Ca * math.sin(dl)),
            "F_perpendicular_R": qS * (Cn_R * math.cos(dr) - Ca * math.sin(dr)),
            "F_perpendicular": f_perp,
            "F_parallel_L": qS * (Ca * math.cos(dl) + Cn_L * math.sin(dl)),
            "F_parallel_R": qS * (Ca * math.cos(dr) + Cn_R * math.sin(dr)),
            "F_parallel": f_par,
            "Fx": f_par * math.cos(theta0) + f_perp * math.sin(theta0),
            "Fy": f_par * math.sin(theta0) - f_perp * math.cos(theta0),
            "Mz": m_z, "d_fin_cg": p.d_base_grid_fin - x_cog, "delta_left_rad": dl, "delta_right_rad": dr}
    if phase in ("subsonic", "supersonic"):
        action_info = {"gimbal_angle_deg": gimbal_deg, "throttle": throttle}
    elif phase == "ballistic_arc_descent":
        action_info = {"RCS_throttle": actions}
    elif phase == "landing_burn":
        action_info = {"throttle": throttle, "delta_command_left_rad": dl_cmd,
                       "delta_command_right_rad": dr_cmd, "gimbal_angle_deg": gimbal_deg,
                       "acs_info": acs_info}
    else:
        action_info = {"throttle": throttle, "acs_info": acs_info}
    return dict(acceleration_dict=acc, moment_dict=mom, mach_number_max=mach_max, drag=drag, lift=lift,
                d_cp_cg=d_cp_cg, d_thrust_cg=d_thrust_cg, fuel_percentage_consumed=fuel,
                control_force_parallel=c_par, control_force_perpendicular=c_perp, control_force_x=c_x,
                control_force_y=c_y, aero_force_x=aero_x, aero_force_y=aero_y, gravity_force_y=-g * mass,
                action_info=action_info)


class simple_actor_spec:
    """Shape bookkeeping of env_wrapped_ea.simple_actor (env_wrapped_ea.py:18-75): parameter
    names/order of nn.Sequential(Linear, ReLU, n x Sequential(Linear, ReLU), Linear, Tanh)."""

    def __init__(self, input_dim, output_dim, number_of_hidden_layers, hidden_dim):
        self.shapes = [("0", (hidden_dim, input_dim))]
        for l in range(number_of_hidden_layers):
            self.shapes.append((f"{2 + l}.0", (hidden_dim, hidden_dim)))
        self.shapes.append((f"{2 + number_of_hidden_layers}", (output_dim, hidden_dim)))
        self.number_of_network_parameters = sum(o * i + o for _, (o, i) in self.shapes)

    def return_setup_vals(self):
        names, bounds = {}, []
        for name, (o, i) in self.shapes:
            for kind, n in (("weight", o * i), ("bias", o)):
                for j in range(n):
                    names[f'{name.replace(".", "_")}_{kind}_{j}'] = 0.0
                    bounds.append((-1.5, 1.5))
        return names, bounds


class pso_wrapped_env:
    """PSO model (env_wrapped_ea.py:137-230): .objective_function(individual) -> float,
    .bounds, .mock_dictionary_of_opt_params, .individual_update_model, .reset,
    .env.truncation_id().  The whole episode (actor MLP in the loop) is one kernel launch;
    `evaluate(positions)` is the batched form used by parallel_evaluate."""

    def __init__(self, flight_phase="subsonic", enable_wind=False, stochastic_wind=False,
                 horiontal_wind_percentile=50, precision="fp64", max_steps=8192, seed=0):
        assert flight_phase in ["subsonic", "supersonic", "flip_over_boostbackburn",
                                "ballistic_arc_descent", "landing_burn_pure_throttle", "landing_burn"]
        self.flight_phase, self.enable_wind = flight_phase, enable_wind
        self._b = BatchedRocketEnv(1, "pso", flight_phase, enable_wind, stochastic_wind,
                                   horiontal_wind_percentile, precision=precision, seed=seed)
        if flight_phase == "landing_burn_pure_throttle":
            self.actor = simple_actor_spec(2, 1, 3, 8)
        else:
            self.actor = simple_actor_spec(5, 4, 4, 8)
        self.mock_dictionary_of_opt_params, self.bounds = self.actor.return_setup_vals()
        self.max_steps = max_steps
        self.capped, self.warn_on_cap = 0, True
        self.experience_buffer = []
        self.episode_idx = 0
        self._individual = None
        self._last_tid = 0
        self.env = self          # model.env.truncation_id() as in the reference

    def truncation_id(self):
        return self._last_tid

    def individual_update_model(self, individual):
        self._individual = np.asarray(individual, dtype=np.float64)

    def reset(self):
        self.experience_buffer = []

    def evaluate(self, positions, n_seeds=1, terminal=False, index0=0, generation=0):
        """Batched objective_function: positions [n, P] (numpy / list, or a cuda float32 tensor that is
        used where it is) -> (fitness[n*n_seeds], steps, trunc_id [, terminal_state]) device tensors.
        `self.capped` = number of episodes of this call that hit max_steps (trunc_id -1)."""
        if isinstance(positions, torch.Tensor) and positions.is_cuda:
            wt = positions.reshape(-1, self.actor.number_of_network_parameters)
        elif not isinstance(positions, np.ndarray) and len(positions) * self.actor.number_of_network_parameters < (1 << 21):
            # a list of per-particle arrays (what ParticleSubswarmOptimisation passes): one conversion
            # straight to float32 - the same rounding as float64 -> float32, half the time
            wt = torch.as_tensor(np.asarray(positions, dtype=np.float32).reshape(
                -1, self.actor.number_of_network_parameters)).to(self._b.device)
        else:
            w = np.ascontiguousarray(np.asarray(positions, dtype=np.float64).reshape(
                -1, self.actor.number_of_network_parameters))
            # float64 -> float32 as update_individiual does (env_wrapped_ea.py:46-59), with torch's
            # multi-threaded cast (numpy's astype is single-threaded: 30 ms for 65 536 x 249), through a
            # pinned staging buffer so that the upload is one DMA
            n = w.shape[0] * w.shape[1]
            if n < (1 << 21):       # small swarms: waking torch's thread pool costs more than the cast
                wt = torch.as_tensor(w.astype(np.float32)).to(self._b.device)
            else:
                if getattr(self, "_stage", None) is None or self._stage.numel() < n:
                    self._stage = torch.empty(n, dtype=torch.float32).pin_memory()
                st = self._stage[:n].view(w.shape)
                st.copy_(torch.from_numpy(w))
                wt = st.to(self._b.device, non_blocking=True)
        out = self._b.rollout_pso(wt, n_seeds=n_seeds, max_steps=self.max_steps, terminal=terminal,
                                  index0=index0, generation=generation)
        self._b.check_status()
        self.capped = int((out[2] < 0).sum())
        if self.capped and self.warn_on_cap:
            import warnings
            warnings.warn(f"{self.capped} episode(s) reached max_steps={self.max_steps} and were scored as "
                          "truncated at their final state (the reference has no step cap)", RuntimeWarning)
        return out

    def objective_function(self, individual):
        self.individual_update_model(individual)
        fit, steps, tid = self.evaluate(self._individual[None, :])
        self._last_tid = int(tid[0])
        self.last_steps = int(steps[0])
        self.episode_idx += 1
        return float(fit[0])


class rl_wrapped_env_pytorch:
    """rl_wrapped_env_pytorch (env_wrapped_rl_pytorch.py:68-205): float32-rounded observation,
    G action log-compression, P-control reference-speed scaling (both inside the step kernel),
    state_dim/action_dim."""

    def __init__(self, flight_phase="subsonic", enable_wind=False, stochastic_wind=True,
                 horiontal_wind_percentile=50, trajectory_length=None, discount_factor=None,
                 precision="fp64", seed=0):
        assert flight_phase in ALL_PHASES
        self.flight_phase = flight_phase
        self.env = rocket_environment_pre_wrap("rl", flight_phase, enable_wind, stochastic_wind,
                                               horiontal_wind_percentile, trajectory_length,
                                               discount_factor, precision=precision, seed=seed,
                                               _wrapped=True)
        self.enable_wind = enable_wind
        self.state_dim, self.action_dim = self.env._b.obs_dim, self.env._b.act_dim

    def truncation_id(self):
        return self.env.truncation_id

    def reset(self):
        self.env.reset()
        # observation of the reset state: one kernel-side observe() via a zero-length trick is
        # not exposed, so evaluate augment_state's closed form on the fp32-rounded state
        return self._obs_from_state(self.env.state)

    def _obs_from_state(self, state):
        s = np.asarray(state, dtype=np.float32)
        nv = np.array(self.env._b.params.norm_vals)     # np.float64 scalars, as upstream
        if self.flight_phase in ("subsonic", "supersonic", "ballistic_arc_descent"):
            idx = [4, 5, 6, 7] if self.flight_phase == "ballistic_arc_descent" else [0, 1, 2, 3, 4, 5, 7, 8]
            o = s[idx].copy()
            o /= np.array(self.env._b.params.other_phases["norm_vals"][self.flight_phase])
            return o
        if self.flight_phase == "landing_burn_pure_throttle_Pcontrol":
            return np.array([(1 - s[1] / nv[0]) * 2 - 1])
        if self.flight_phase == "landing_burn_pure_throttle":
            return np.array([(1 - s[1] / nv[0]) * 2 - 1, (1 - s[3] / nv[1]) * 2 - 1])
        k = float(np.arctanh(0.75) / math.radians(5))
        kd = float(np.arctanh(0.75) / 0.01)
        return np.array([s[1] / nv[0], s[3] / nv[1], math.tanh(k * (s[4] - math.pi / 2)),
                         math.tanh(kd * s[5]), math.tanh(k * (s[6] - 3 / 2 * math.pi))])

    def step(self, action):
        if isinstance(action, np.ndarray):
            a = action
        else:
            try:
                a = action.cpu().numpy()
            except Exception:
                a = np.array(action)
        if a.ndim == 2:
            a = a[0]
        state, reward, done, truncated, info = self.env.step(a)
        obs = self.env._obs.to(torch.float64).cpu().numpy()
        if self.flight_phase in ("subsonic", "supersonic", "ballistic_arc_descent"):
            obs = obs.astype(np.float32)        # upstream hands back a float32 array for these phases
        return obs, float(reward), bool(done), bool(truncated), info

    def __getattr__(self, name):
        return getattr(self.env, name)

    def render(self):
        pass

    def close(self):
        pass


class supervisory_wrapper:
    """supervisory_wrapper (src/envs/supervisory/env_wrapped_supervisory.py:6-116): the base env with
    type='supervisory' (reward 0, rtd_supervisory_mock.py verdicts), float64 observation scaled by
    the caller's `input_normalisation_values`, action shaping on the host in float64 (the base env
    then takes the action as it is).  As upstream, the wind arguments are accepted and ignored."""

    _SLICES = {"subsonic": [0, 1, 2, 3, 4, 5, 7, 8], "supersonic": [0, 1, 2, 3, 4, 5, 7, 8],
               "landing_burn": [0, 1, 2, 3, 4, 5, 7, 8], "ballistic_arc_descent": [4, 5, 6, 7],
               "flip_over_boostbackburn": [4, 5]}

    def __init__(self, input_normalisation_values, flight_phase="subsonic", enable_wind=False,
                 stochastic_wind=False, horiontal_wind_percentile=95, precision="fp64", seed=0):
        assert flight_phase in ["subsonic", "supersonic", "flip_over_boostbackburn", "ballistic_arc_descent",
                                "landing_burn", "landing_burn_pure_throttle",
                                "landing_burn_pure_throttle_Pcontrol"]
        self.flight_phase = flight_phase
        self.input_normalisation_values = input_normalisation_values
        if flight_phase == "landing_burn_pure_throttle":
            self.input_normalisation_values = input_normalisation_values[:2]
        self.enable_wind, self.stochastic_wind, self.horiontal_wind_percentile = False, False, 95
        self.env = rocket_environment_pre_wrap("supervisory", flight_phase, False, False, 95,
                                               precision=precision, seed=seed)
        self.initial_mass = self.env.reset()[-2]
        if flight_phase == "landing_burn_pure_throttle_Pcontrol":
            vx0, vy0 = self.env._b.params.initial_state[2:4]
            self.speed0 = math.sqrt(vx0 ** 2 + vy0 ** 2)

    def truncation_id(self):
        return self.env.truncation_id

    def augment_state(self, state):
        nv = self.input_normalisation_values
        if self.flight_phase in self._SLICES:
            return np.array([state[i] for i in self._SLICES[self.flight_phase]]) / nv
        if self.flight_phase == "landing_burn_pure_throttle":
            return np.array([(1 - state[1] / nv[0]) * 2 - 1, (1 - state[3] / nv[1]) * 2 - 1])
        return np.array([(1 - state[1] / nv[0]) * 2 - 1])

    def augment_action(self, actions):
        if self.flight_phase == "landing_burn":
            u = actions[0] if actions.ndim == 2 else actions

            def squash(v, c):
                return math.copysign(math.log(1 + c * abs(v)) / math.log(1 + c), v)
            shaped = [squash(u[0], 10), u[1], squash(u[2], 5), squash(u[3], 5)]
            actions = np.array([shaped]) if actions.ndim == 2 else np.array(shaped)
        if self.flight_phase == "landing_burn_pure_throttle_Pcontrol":
            u0 = actions[0] if actions.ndim == 2 else actions
            actions = np.array([(u0 + 1) / 2 * self.speed0])
        return actions

    def step(self, action):
        state, reward, done, truncated, info = self.env.step(self.augment_action(np.array(action)))
        return self.augment_state(state), reward, done, truncated, info

    def reset(self):
        return self.augment_state(self.env.reset())
