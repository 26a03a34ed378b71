"""Import harness for the *unmodified* upstream reference (container-only tool).

TEST/FIXTURE INFRASTRUCTURE - never imported by the product package.

The upstream reference (/root/reference, read-only) cannot be imported as-is in
this image: `ambiance`, `matplotlib`, `gymnasium`, `pyswarm`,
`lmdb` are absent and the two dill pickles carry Python <= 3.10 bytecode.  This
module makes the reference's own hot-path modules importable *without touching
them*:

  1. MagicMock stubs for the plotting / gym / lmdb imports (never executed on
     the hot path).
  2. A stand-in `ambiance.Atmosphere` restating the ICAO-1993 ISA (the published
     algorithm of the third-party package; version unpinned upstream).
  3. `dill.load` wrapped so the pickled closures are re-bound to the reference's
     own *source* functions from their closure cells (stage_inertia,
     d_cg_thrusters, cop_func).
  4. cwd = reference root (all its data paths are relative).

It exists only to (a) validate oracle/ against the real reference and (b)
generate the committed fixtures under tests/golden/ (tools/make_golden.py).
/root/reference does not exist on the GPU box, so nothing under tests/ -m gpu,
bench.py or __graft_entry__.py may import this file.
"""
from __future__ import annotations

import math
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

REFERENCE_ROOT = os.environ.get("PD_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.patches",
    "matplotlib.lines", "matplotlib.cm", "matplotlib.colors", "matplotlib.ticker",
    "matplotlib.animation", "matplotlib.collections", "mpl_toolkits",
    "mpl_toolkits.mplot3d", "pyswarm", "gymnasium", "lmdb",
]


# --------------------------------------------------------------------------
# ambiance stand-in (ICAO 1993 ISA, geopotential formulation)
# --------------------------------------------------------------------------
_ISA_LAYERS = (
    # H_b [m], T_b [K], beta [K/m], p_b [Pa]
    (-5.0e3, 320.65, -6.5e-3, 1.77687e5),
    (0.0e3, 288.15, -6.5e-3, 1.01325e5),
    (11.0e3, 216.65, 0.0, 2.26320e4),
    (20.0e3, 216.65, 1.0e-3, 5.47487e3),
    (32.0e3, 228.65, 2.8e-3, 8.68014e2),
    (47.0e3, 270.65, 0.0, 1.10906e2),
    (51.0e3, 270.65, -2.8e-3, 6.69384e1),
    (71.0e3, 214.65, -2.0e-3, 3.95639e0),
)
_G0, _R, _KAPPA, _REARTH = 9.80665, 287.05287, 1.4, 6_356_766.0


class _Atmosphere:
    def __init__(self, h):
        h = np.atleast_1d(np.asarray(h, dtype=float))
        H = _REARTH * h / (_REARTH + h)
        T = np.empty_like(H)
        p = np.empty_like(H)
        for i, Hi in enumerate(H):
            k = 0
            for j, lay in enumerate(_ISA_LAYERS):
                if Hi >= lay[0]:
                    k = j
            Hb, Tb, beta, pb = _ISA_LAYERS[k]
            Ti = Tb + beta * (Hi - Hb)
            if beta == 0.0:
                pi = pb * math.exp(-_G0 / (_R * Ti) * (Hi - Hb))
            else:
                pi = pb * (1.0 + (beta / Tb) * (Hi - Hb)) ** (-_G0 / (beta * _R))
            T[i] = Ti
            p[i] = pi
        self.temperature = T
        self.pressure = p
        self.density = p / (_R * T)
        self.speed_of_sound = np.sqrt(_KAPPA * _R * T)


def _install_stubs():
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    amb = types.ModuleType("ambiance")
    amb.Atmosphere = _Atmosphere
    sys.modules["ambiance"] = amb


def _patch_dill():
    import dill

    if getattr(dill, "_pd_patched", False):
        return
    real_load = dill.load

    def _cells(fn):
        return {n: c.cell_contents for n, c in zip(fn.__code__.co_freevars, fn.__closure__ or ())}

    def load(f, *a, **k):
        obj = real_load(f, *a, **k)
        if not isinstance(obj, dict):
            # velocity-profile pickle (landing_burn_pure_throttle.py:111-116):
            # v_opt = lambda y: a_opt * y**2 + b_opt * y
            if callable(obj) and hasattr(obj, "__code__") and \
                    set(obj.__code__.co_freevars) == {"a_opt", "b_opt"}:
                c = _cells(obj)
                a_opt, b_opt = c["a_opt"], c["b_opt"]
                return lambda y: a_opt * y**2 + b_opt * y
            return obj
        from src.RocketSizing.functions.rocket_dimensions import (
            stage_inertia, full_rocket_inertia, d_cg_thrusters)
        from src.RocketSizing.functions.cop_estimation import cop_func
        out = {}
        for key, fn in obj.items():
            if not callable(fn) or not hasattr(fn, "__code__"):
                out[key] = fn
                continue
            cells = _cells(fn)
            if key.startswith("x_cog_inertia"):
                if "m_s_1" in cells:
                    out[key] = full_rocket_inertia(**cells)
                else:
                    out[key] = stage_inertia(**cells)
            elif key.startswith("d_cg_thrusters"):
                eh = cells["self"].engine_height
                out[key] = (lambda eh: (lambda x_cog: d_cg_thrusters(x_cog, eh)))(eh)
            elif key.startswith("cop_subrocket"):
                idx = int(key.split("_")[2])
                L = cells["self"].lengths[idx]
                d0 = (0.25, 0.25, 0.75)[idx]
                out[key] = (lambda L, d0: (lambda alpha, M: cop_func(L, alpha, M, d_0=d0)))(L, d0)
            else:
                out[key] = fn
        return out

    dill.load = load
    dill._pd_patched = True


_LOADED = False


def load_reference():
    """Make `import src.envs...` resolve to the unmodified reference; chdir to its root."""
    global _LOADED
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT} (container-only tool)")
    if _LOADED:
        os.chdir(REFERENCE_ROOT)
        return
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    os.chdir(REFERENCE_ROOT)
    _patch_dill()
    _LOADED = True


class NoiseTapeWind:
    """Monkey-patch helper: make the reference's VonKarmanFilter consume an external
    N(0,1) tape and a fixed (sigma_u, sigma_v) instead of the global numpy RNG /
    `random.uniform`.  The tape is consumed in the
    reference's own draw order: u-filter then v-filter, once per physics sub-step
    while y < 15 km."""

    def __init__(self, tape, sigma_u, sigma_v):
        self.tape = np.asarray(tape, dtype=float).ravel()
        self.pos = 0
        self.sigma_u = float(sigma_u)
        self.sigma_v = float(sigma_v)

    def install(self):
        import src.envs.wind.vonkarman as vk
        outer = self

        def step(filt):
            w = outer.tape[outer.pos]
            outer.pos += 1
            filt.state = filt.Ad @ filt.state + filt.Bd * w
            return float((filt.Cd @ filt.state)[0])

        def _new_filters(gen):
            u = vk.VonKarmanFilter(gen.L_u, outer.sigma_u, gen.V, gen.dt)
            v = vk.VonKarmanFilter(gen.L_v, outer.sigma_v, gen.V, gen.dt)
            gen.sigma_u, gen.sigma_v = outer.sigma_u, outer.sigma_v
            return u, v

        self._orig = (vk.VonKarmanFilter.step, vk.VKDisturbanceGenerator._new_filters)
        vk.VonKarmanFilter.step = step
        vk.VKDisturbanceGenerator._new_filters = _new_filters

    def uninstall(self):
        import src.envs.wind.vonkarman as vk
        vk.VonKarmanFilter.step, vk.VKDisturbanceGenerator._new_filters = self._orig
