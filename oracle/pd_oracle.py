"""CPU oracle for the powered-descent hot path  --  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does (it fails loudly when
its CUDA extension is missing instead of falling back to this).

It is a scalar, pure-Python/numpy/scipy restatement of the reference's algorithm
for the two landing phases that work in PSO and RL mode

    P = 'landing_burn_pure_throttle'   (1 action, dt_phys 0.025 x 4)
    G = 'landing_burn'                 (4 actions, dt_phys 0.1 x 4)

and for the four further phases that work upstream in RL mode only (their PSO closures
have the wrong arity, rtd_pso.py:38-157; 'flip_over_boostbackburn' raises TypeError in
RL mode too, rtd_rl.py:132, and 'landing_burn_ACS' is broken in compile_physics)

    S = 'subsonic', U = 'supersonic'   (2 actions: gimbal, throttle; dt 0.1 x 1)
    B = 'ballistic_arc_descent'        (1 action: RCS; dt 0.1 x 1)
    C = 'landing_burn_pure_throttle_Pcontrol'  (1 action: reference speed; dt 0.1 x 1)

following, function by function (paths relative to /root/reference):

    isa()                 src/envs/utils/atmosphere_dynamics.py:5-27 + the third-party
                          `ambiance` package (ICAO-1993 ISA; NOT under /root/reference,
                          version unpinned upstream).  PARITY UNPINNED at the ambiance
                          boundary: its published algorithm is restated here and pinned only
                          indirectly, through the reference's golden trajectories
    gravity()             src/envs/utils/atmosphere_dynamics.py:29-33
    cd()/cl()             src/envs/utils/aerodynamic_coefficients.py:57-66,105-132 and
                          src/envs/rockets_physics.py:711-712 (degrees passed twice)
    grid_fin_ca()/cn()    src/envs/utils/grid_fin_aerodynamics.py:7-46
    acs()                 src/envs/utils/acs_model.py:13-86
    cog_inertia()         src/RocketSizing/functions/rocket_dimensions.py:167-196
    control_P()/control_G()  src/envs/rockets_physics.py:340-400 / 168-269
    control_ascent()/control_rcs()/control_C()  rockets_physics.py:17-56 / 149-166 / 402-451
    control_flip()        src/envs/rockets_physics.py:63-92 (+ :542-560: no aerodynamic forces)
    cog_inertia_full()    src/RocketSizing/functions/rocket_dimensions.py:199-241
    ascent / ballistic / P-control rtd   src/envs/rl/rtd_rl.py:11-114, 147-188, 353-534
    substep()             src/envs/rockets_physics.py:455-646
    wind                  src/envs/wind/{full_wind_model.py:35-49, vonkarman.py:9-96,
                          HorizontalWindSpeed.py:44-114}
    step()/reset()        src/envs/base_environment.py:80-154
    rtd closures          src/envs/pso/rtd_pso.py:172-317, src/envs/rl/rtd_rl.py:190-336
    PsoModel              src/envs/pso/env_wrapped_ea.py:18-222
    RlEnv                 src/envs/rl/env_wrapped_rl_pytorch.py:25-47,120-202
    classical_rollout()   src/classical_controls/landing_burn_pure_throttle.py:261-339

Numeric-type discipline: the reference's results depend on NumPy's NEP-50 scalar
promotion (a float32 action makes throttle / thrust / mass-flow float32, SURVEY.md
section 8a "dtype rule").  This restatement therefore keeps the same scalar *types*
the reference has at every point (np.float64 state scalars as pandas yields them,
Python floats for atmosphere values and constants, the caller's dtype for the
action) and lets NumPy promote, rather than emulating the rule by hand.

Parity pin: tests/test_oracle_golden.py checks this file against fixtures generated
by running the unmodified reference in the build container (tools/make_golden.py)
and against the reference's own committed CSVs.
"""
from __future__ import annotations

import json
import math
import os

import numpy as np
from scipy.interpolate import RBFInterpolator, interp1d
from scipy.signal import cont2discrete

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SNAPSHOT = os.path.join(
    os.path.dirname(_HERE), "psso_sac_for_powered_descent_b200", "data",
    "rocket_parameters_snapshot.json")

PHASE_P = "landing_burn_pure_throttle"
PHASE_G = "landing_burn"
PHASE_S = "subsonic"
PHASE_U = "supersonic"
PHASE_B = "ballistic_arc_descent"
PHASE_C = "landing_burn_pure_throttle_Pcontrol"
PHASE_F = "flip_over_boostbackburn"
RL_ONLY_PHASES = (PHASE_S, PHASE_U, PHASE_B, PHASE_C)

# Mach-scheduled truncation thresholds / reward weights of the ascent phases
# (rtd_rl.py:544-575): [mach, max_x_error, max_vy_error, max_vx_error, max_alpha_deg,
#  alpha_w, x_w, vy_w, vx_w]
SUBSONIC_HYPER = [
    [0.0, 50, 10, 10, 0.5, 100, 100, 100, 100], [0.1, 50, 15, 10, 10, 100, 100, 100, 100],
    [0.2, 50, 20, 5, 2, 100, 100, 100, 100], [0.3, 50, 20, 5, 2, 100, 100, 100, 100],
    [0.4, 50, 20, 5, 2, 100, 100, 100, 100], [0.5, 50, 20, 5, 2, 100, 100, 100, 100],
    [0.6, 50, 20, 5, 1.75, 100, 100, 100, 100], [0.7, 50, 20, 5, 1.75, 100, 100, 100, 100],
    [0.8, 50, 20, 5, 1.75, 100, 100, 100, 100], [0.9, 50, 20, 5, 1.75, 100, 100, 100, 100],
    [1.0, 50, 20, 5, 1.75, 100, 100, 100, 100], [1.1, 50, 20, 5, 1.75, 100, 100, 100, 100]]
SUPERSONIC_HYPER = [
    [1.0, 100, 50, 9, 8, 100, 100, 100, 100], [1.1, 100, 60, 20, 8, 100, 100, 100, 100],
    [1.5, 100, 60, 20, 8, 100, 100, 100, 100], [1.75, 100, 60, 30, 8, 100, 100, 100, 100],
    [2.0, 100, 60, 40, 8, 100, 100, 100, 100], [2.25, 100, 60, 50, 8, 100, 100, 100, 100],
    [2.5, 100, 60, 60, 8, 100, 100, 100, 100], [2.75, 100, 60, 70, 8, 100, 100, 100, 100],
    [3.0, 100, 60, 80, 8, 100, 100, 100, 100], [3.25, 100, 60, 90, 8, 100, 100, 100, 100],
    [3.5, 100, 60, 100, 8, 100, 100, 100, 100], [3.75, 100, 60, 100, 8, 100, 100, 100, 100]]

# ---------------------------------------------------------------------------
# ISA (ambiance restatement)
# ---------------------------------------------------------------------------
ISA_LAYERS = (
    (-5.0e3, 320.65, -6.5e-3, 1.77687e5),
    (0.0e3, 288.15, -6.5e-3, 1.01325e5),
    (11.0e3, 216.65, 0.0, 2.26320e4),
    (20.0e3, 216.65, 1.0e-3, 5.47487e3),
    (32.0e3, 228.65, 2.8e-3, 8.68014e2),
    (47.0e3, 270.65, 0.0, 1.10906e2),
    (51.0e3, 270.65, -2.8e-3, 6.69384e1),
    (71.0e3, 214.65, -2.0e-3, 3.95639e0),
)
ISA_G0, ISA_R, ISA_KAPPA, ISA_REARTH = 9.80665, 287.05287, 1.4, 6_356_766.0


def isa(alt):
    """(rho, p, a) as Python floats; alt<0 -> 0; >= 81020 m -> zeros."""
    if alt < 0:
        alt = 0
    if not alt < 81020:
        return 0.0, 0.0, 0.0
    h = float(alt)
    H = ISA_REARTH * h / (ISA_REARTH + h)
    k = 0
    for j, lay in enumerate(ISA_LAYERS):
        if H >= lay[0]:
            k = j
    Hb, Tb, beta, pb = ISA_LAYERS[k]
    T = Tb + beta * (H - Hb)
    if beta == 0.0:
        p = pb * math.exp(-ISA_G0 / (ISA_R * T) * (H - Hb))
    else:
        p = pb * (1.0 + (beta / Tb) * (H - Hb)) ** (-ISA_G0 / (beta * ISA_R))
    rho = p / (ISA_R * T)
    a = math.sqrt(ISA_KAPPA * ISA_R * T)
    return float(rho), float(p), float(a)


def gravity(alt):
    R = 6371000
    return 9.80665 * (R / (R + alt)) ** 2


# ---------------------------------------------------------------------------
class _LocalRBF:
    """scipy RBFInterpolator(kernel='thin_plate_spline', neighbors=50) as the reference
    builds it.  `fast=True` memoises the per-neighbourhood solve and skips np.unique;
    it calls scipy's own tree / solve / evaluation routines in the same order, so the
    value is bit-identical to `interp(pts)` (checked in tests)."""

    def __init__(self, mach, aoa, val, fast=False):
        pts = np.column_stack((np.asarray(mach, float), np.asarray(aoa, float)))
        self.interp = RBFInterpolator(pts, np.asarray(val, float),
                                      kernel="thin_plate_spline", neighbors=50)
        self.fast = fast
        self._cache = {}
        if fast:
            from scipy.interpolate import _rbfinterp_np as _np_backend
            self._backend = _np_backend

    def __call__(self, mach_val, aoa_val):
        pts = np.array([[mach_val, aoa_val]])
        if not self.fast:
            return float(self.interp(pts)[0])
        it = self.interp
        x = np.asarray(pts, dtype=np.float64, order="C")
        _, yidx = it._tree.query(x, it.neighbors)
        yidx = np.sort(yidx, axis=1)[0]
        key = yidx.tobytes()
        ent = self._cache.get(key)
        if ent is None:
            shift, scale, coeffs = self._backend._build_and_solve_system(
                it.y[yidx], it.d[yidx], it.smoothing[yidx], it.kernel, it.epsilon,
                it.powers, np)
            ent = (it.y[yidx], shift, scale, coeffs)
            self._cache[key] = ent
        ynbr, shift, scale, coeffs = ent
        out = self._backend.compute_interpolation(
            x, ynbr, it.kernel, it.epsilon, it.powers, shift, scale, coeffs, np)
        return float(out[0, 0])

    def neighbourhood(self, mach_val, aoa_val):
        x = np.array([[mach_val, aoa_val]], dtype=np.float64)
        _, yidx = self.interp._tree.query(x, self.interp.neighbors)
        return np.sort(yidx, axis=1)[0]


def _vk_discrete(L, sigma, V, dt):
    omega0 = V / L
    zeta = 1.0 / math.sqrt(2.0)
    scale = math.sqrt(math.pi / (2.0 * omega0 ** 3))
    A = np.array([[0.0, 1.0], [-omega0 ** 2, -2.0 * zeta * omega0]])
    B = np.array([[0.0], [sigma * scale]])
    C = np.array([[0.0, 1.0]])
    D = np.zeros((1, 1))
    Ad, Bd, Cd, _, _ = cont2discrete((A, B, C, D), dt)
    return Ad, Bd.flatten(), Cd


def wind_profile(wind_table, percentile):
    """Sorted (altitude_km, wind_speed) for a requested percentile.  The reference
    passes a number, which never matches the string keys, so it always blends two
    tabulated percentiles (HorizontalWindSpeed.py:47-51, 72-114)."""
    if percentile in wind_table:
        speed = np.array(wind_table[percentile]["wind_speed"])
        alt = np.array(wind_table[percentile]["altitude_km"])
    else:
        req = float(percentile)
        names = list(wind_table.keys())
        vals = [float(n.split("_")[0]) for n in names]
        idx = np.searchsorted(vals, req)
        if idx == 0:
            lo = hi = names[0]
            w = 1.0
        elif idx == len(vals):
            lo = hi = names[-1]
            w = 0.0
        else:
            lo, hi = names[idx - 1], names[idx]
            w = (req - vals[idx - 1]) / (vals[idx] - vals[idx - 1])
        lo_alt = np.array(wind_table[lo]["altitude_km"])
        hi_alt = np.array(wind_table[hi]["altitude_km"])
        alt = np.unique(np.concatenate([lo_alt, hi_alt]))
        lo_f = interp1d(lo_alt, np.array(wind_table[lo]["wind_speed"]), kind="linear",
                        bounds_error=False, fill_value="extrapolate")
        hi_f = interp1d(hi_alt, np.array(wind_table[hi]["wind_speed"]), kind="linear",
                        bounds_error=False, fill_value="extrapolate")
        speed = lo_f(alt) * (1 - w) + hi_f(alt) * w
    order = np.argsort(alt)
    return alt[order], speed[order]


class OracleWind:
    """WindModel + VKDisturbanceGenerator with an injectable noise tape.

    noise = dict(sigma_u, sigma_v, tape) consumes tape[k] in the reference's draw
    order (u then v per sub-step while y < 15 km).  Without it the reference's
    own unseedable RNG calls are used (np.random.randn / random.uniform)."""

    def __init__(self, wind_table, dt, stochastic, percentile, noise=None):
        self.dt = dt
        self.V = 100
        self.y_thr = 15000
        self.stochastic = stochastic
        self.noise = noise
        self.alt_km, self.speed = wind_profile(wind_table, percentile)
        self._f = interp1d(self.alt_km, self.speed, kind="linear", bounds_error=False,
                           fill_value=(self.speed[0], self.speed[-1]))
        self.reset()

    def reset(self):
        import random
        if self.noise is not None:
            self.sigma_u, self.sigma_v = float(self.noise["sigma_u"]), float(self.noise["sigma_v"])
            self.tape = np.asarray(self.noise["tape"], float).ravel()
            self.pos = 0
        else:
            np.random.seed(None)
            self.sigma_u = random.uniform(0.5, (0.5 + 4.0) / 2)
            self.sigma_v = random.uniform((0.5 + 2.0) / 2, 2.0)
            self.tape = None
        self.Adu, self.Bdu, self.Cdu = _vk_discrete(100.0, self.sigma_u, self.V, self.dt)
        self.Adv, self.Bdv, self.Cdv = _vk_discrete(30.0, self.sigma_v, self.V, self.dt)
        self.xu = np.zeros(2)
        self.xv = np.zeros(2)

    def _draw(self):
        if self.tape is not None:
            w = self.tape[self.pos]
            self.pos += 1
            return w
        return np.random.randn()

    def __call__(self, y):
        fixed = self._f(y / 1000.0)
        if y < self.y_thr and self.stochastic:
            self.xu = self.Adu @ self.xu + self.Bdu * self._draw()
            gu = float((self.Cdu @ self.xu)[0])
            self.xv = self.Adv @ self.xv + self.Bdv * self._draw()
            gv = float((self.Cdv @ self.xv)[0])
        else:
            gu, gv = 0, 0
        return fixed + gu, gv


# ---------------------------------------------------------------------------
class Tables:
    """Everything compile_physics / module import side effects set up."""

    def __init__(self, snapshot=None, fast_rbf=False):
        with open(snapshot or DEFAULT_SNAPSHOT) as f:
            p = json.load(f)
        self.p = p
        self.cd_rbf = _LocalRBF(p["cd_mach"], p["cd_aoa"], p["cd_val"], fast=fast_rbf)
        self.cl_rbf = _LocalRBF(p["cl_mach"], p["cl_aoa"], p["cl_val"], fast=fast_rbf)
        ca_m, ca_v = np.array(p["gf_ca_mach"]), np.array(p["gf_ca_val"])
        self.ca_min_mach = np.min(ca_m)
        self.ca_f = interp1d(ca_m, ca_v, kind="linear", fill_value="extrapolate")
        self.ca_min = ca_v[np.argmin(ca_m)]
        cn_m, cn_v = np.array(p["gf_cn_mach"]), np.array(p["gf_cn_val"])
        self.cn_min_mach, self.cn_max_mach = np.min(cn_m), np.max(cn_m)
        self.cn_f = interp1d(cn_m, cn_v, kind="linear")
        self.cn_min = cn_v[np.argmin(cn_m)]
        srt = np.argsort(cn_m)
        self.cn_max = cn_v[srt[-1]]
        self.cn_slope = (cn_v[srt[-1]] - cn_v[srt[-2]]) / (cn_m[srt[-1]] - cn_m[srt[-2]])
        self.m_prop0 = p["propellant_mass_stage1_ton"] * 1000
        burnout = (p["stage1_mass_ton"] - p["propellant_mass_stage1_ton"]) * 1000.0
        self.c_gust_x = 2 * burnout * (0.5 * 9.81) / (1.225 * (10 + 6.0) ** 2 * p["frontal_area"])
        self.c_gust_y = 0.0
        self.cop = p["cop_d0"] * p["cop_length"]
        # pandas hands the reference np.float64 scalars
        self.initial_state = [np.float64(v) for v in p["initial_state"]]
        self.norm_vals = np.array(p["norm_vals"])
        self.inertia = {k: np.float64(v) for k, v in p["inertia"].items()}
        o = p.get("other_phases") or {}
        self.other = o
        if o:
            self.inertia_full = {k: np.float64(v) for k, v in o["inertia_full"].items()}
            self.cop_full = o["cop_d0_full"] * o["cop_length_full"]
            self.initial_states = {k: [np.float64(v) for v in st] for k, st in o["initial_states"].items()}
            rt = o["ref_traj_ascent"]
            # reference_trajectory_interpolation.py:14-16 (only x, vx, vy are consumed)
            self.ref_x = interp1d(rt["y"], rt["x"], kind="linear", fill_value="extrapolate")
            self.ref_vx = interp1d(rt["y"], rt["vx"], kind="linear", fill_value="extrapolate")
            self.ref_vy = interp1d(rt["y"], rt["vy"], kind="linear", fill_value="extrapolate")

    # -- aero -----------------------------------------------------------
    def cd(self, mach, alpha_rad):
        aoa = math.degrees(alpha_rad)          # CD_func passes degrees ...
        lim = math.radians(10)                 # ... into a clamp written for radians
        if aoa > lim:
            return self.cd_rbf(mach, lim)
        elif aoa < -lim:
            return self.cd_rbf(mach, -lim)
        return self.cd_rbf(mach, aoa)

    def cl(self, mach, alpha_rad):
        aoa_deg = math.degrees(math.degrees(alpha_rad))   # converted twice
        if aoa_deg > 10:
            return self.cl_rbf(mach, 10)
        elif aoa_deg < -10:
            return self.cl_rbf(mach, -10)
        elif abs(aoa_deg) < 1e-6:
            return 0.0
        elif aoa_deg < 0:
            return -self.cl_rbf(mach, abs(aoa_deg))
        return self.cl_rbf(mach, aoa_deg)

    def grid_fin_ca(self, mach):
        if mach < self.ca_min_mach:
            return self.ca_min
        return self.ca_f(mach)

    def grid_fin_cn(self, mach, alpha_rad):
        deg = math.degrees(alpha_rad)
        if mach < self.cn_min_mach:
            return self.cn_min * deg
        elif mach <= self.cn_max_mach:
            return self.cn_f(mach) * deg
        return (self.cn_max + self.cn_slope * (mach - self.cn_max_mach)) * deg

    def cog_inertia(self, fill):
        c = self.inertia
        h_ox_t = c["h_ox"] * fill
        h_f_t = c["h_f"] * fill
        m_ox_t = c["m_ox"] * fill
        m_f_t = c["m_f"] * fill
        x_prop = (m_ox_t * (c["h_lower"] + h_ox_t / 2)
                  + m_f_t * (c["h_lower"] + c["h_ox"] + h_f_t / 2)) / (m_ox_t + m_f_t)
        I_ox = 1 / 12 * m_ox_t * h_ox_t ** 2 + m_ox_t * (c["h_lower"] + h_ox_t / 2 - x_prop) ** 2
        I_f = 1 / 12 * m_f_t * h_f_t ** 2 + \
            m_f_t * (c["h_lower"] + c["h_ox"] + h_f_t / 2 - x_prop) ** 2
        I_prop = I_ox + I_f
        x_wet = (c["m_dry"] * c["x_dry"] + (m_ox_t + m_f_t) * x_prop) / (c["m_dry"] + m_ox_t + m_f_t)
        I_dry_hat = c["I_dry"] + c["m_dry"] * (c["x_dry"] - x_wet) ** 2
        I_prop_hat = I_prop + (m_ox_t + m_f_t) * (x_prop - x_wet) ** 2
        return x_wet, I_dry_hat + I_prop_hat

    def cog_inertia_full(self, fill):
        """full_rocket_inertia (rocket_dimensions.py:199-241), the ascent phases' closure."""
        c = self.inertia_full
        h_ox = c["h_1_ox"] * fill
        h_f = c["h_1_f"] * fill
        m_ox = c["m_1_ox"] * fill
        m_f = c["m_1_f"] * fill
        m_prop = m_ox + m_f
        x_prop = (m_ox * (c["h_lower_1"] + h_ox / 2)
                  + c["m_1_f"] * (c["h_lower_1"] + h_ox + h_f / 2)) / (m_ox + m_f)
        I_ox = 1 / 12 * m_ox * h_ox ** 2 + m_ox * (c["h_lower_1"] + h_ox / 2 - x_prop) ** 2
        I_f = 1 / 12 * m_f * h_f ** 2 + m_f * (c["h_lower_1"] + h_ox + h_f / 2 - x_prop) ** 2
        I_prop = I_ox + I_f
        x_rocket = (c["m_s_1"] * c["x_dry_1"] + (c["m_2"] + c["m_pay"]) * (c["x_wet_2_initial"] + c["h_1"])
                    + m_prop * x_prop) / (c["m_s_1"] + c["m_2"] + c["m_pay"] + m_prop)
        I_rocket = c["I_dry_1"] + c["m_s_1"] * (c["x_dry_1"] - x_rocket) ** 2 \
            + c["I_wet_2_initial"] + c["m_2"] * (c["x_wet_2_initial"] - x_rocket) ** 2 \
            + I_prop + m_prop * (x_prop - x_rocket) ** 2
        return x_rocket, I_rocket

    def acs(self, alpha_eff, theta, q, mach, x_cog, cmd_left_deg, cmd_right_deg,
            prev_left, prev_right, dt):
        p = self.p
        d_cmd_l = math.radians(cmd_left_deg * 60)
        d_cmd_r = math.radians(cmd_right_deg * 60)
        d_l = prev_left + dt * ((-prev_left + d_cmd_l) / 0.5)
        d_r = prev_right + dt * ((-prev_right + d_cmd_r) / 0.5)
        a_l = alpha_eff - d_l
        a_r = alpha_eff - d_r
        qS = q * p["grid_fin_area"]
        Ca = self.grid_fin_ca(mach)
        Cn_L = self.grid_fin_cn(mach, a_l)
        Cn_R = self.grid_fin_cn(mach, a_r)
        f_perp = qS * (Cn_R * math.cos(d_r) - Cn_L * math.cos(d_l)
                       - Ca * (math.sin(d_l) - math.sin(d_r)))
        f_par = qS * (Ca * (2 + math.cos(d_l) + math.cos(d_r))
                      - Cn_L * math.sin(d_l) + Cn_R * math.sin(d_r))
        m_z = -(p["d_base_grid_fin"] - x_cog) * f_perp + p["rocket_radius"] * qS * (
            Ca * (math.sin(d_r) - math.sin(d_l)) - Cn_L * math.cos(d_l) + Cn_R * math.cos(d_r))
        return f_perp, f_par, m_z, d_cmd_l, d_cmd_r


def _unpack_action_P(actions):
    if not isinstance(actions, tuple) and not isinstance(actions, list):
        if actions.ndim == 2:
            return actions[0][0]
        return actions[0]
    elif isinstance(actions, list):
        return float(actions[0])
    return actions


def _unpack_action_G(actions):
    if not isinstance(actions, tuple):
        if actions.ndim == 2:
            return actions[0]
        return actions
    return actions


class OracleEnv:
    """Scalar env == rocket_environment_pre_wrap: phases P and G with type 'pso' | 'rl',
    phases S, U, B, C with type 'rl' (their PSO closures are broken upstream)."""

    def __init__(self, flight_phase=PHASE_P, type="pso", enable_wind=False,
                 stochastic_wind=False, horiontal_wind_percentile=50, tables=None,
                 wind_noise=None, fast_rbf=False, trajectory_length=1, discount_factor=0.99):
        assert flight_phase in (PHASE_P, PHASE_G, PHASE_F) + RL_ONLY_PHASES
        assert type in ("pso", "rl", "supervisory")
        if flight_phase == PHASE_F and type != "supervisory":
            raise TypeError("flip_over_boostbackburn: only type='supervisory' works upstream (its rl and pso "
                            "truncated_func take one argument, rtd_rl.py:132 / rtd_pso.py:107)")
        if flight_phase in RL_ONLY_PHASES and type == "pso":
            raise TypeError(f"{flight_phase}: the reference's pso closures have the wrong arity "
                            "(rtd_pso.py:38-157); only type='rl' works upstream")
        self.trajectory_length = trajectory_length
        self.discount_factor = discount_factor
        self.T = tables or Tables(fast_rbf=fast_rbf)
        self.flight_phase = flight_phase
        self.type = type
        self.dt = 0.1
        p = self.T.p
        n_gim = int(p["n_engines_gimballed"])
        self.n_sub = 4
        if flight_phase == PHASE_P:
            self.n_eng = n_gim
            self.nominal_throttle = (0 * 0.4) / n_gim
            self.dt_phys = 0.025
            self.dt_act = 0.025
        elif flight_phase == PHASE_G:
            self.n_eng = n_gim + 2
            self.nominal_throttle = (3 * 0.4) / n_gim
            self.dt_phys = 0.1        # landing_burn integrates with the env dt ...
            self.dt_act = 0.025       # ... but filters its actuators with dt_temp
        else:
            # one Euler step of the env dt, no sub-stepping (rockets_physics.py:727-801, 959-997)
            self.n_sub = 1
            self.n_eng = n_gim
            self.nominal_throttle = 0.5 if flight_phase in (PHASE_S, PHASE_U) else (0 * 0.4) / n_gim
            self.dt_phys = 0.1
            self.dt_act = 0.1
        self.ascent = flight_phase in (PHASE_S, PHASE_U)
        self.enable_wind = enable_wind
        if enable_wind:
            self.wind = OracleWind(p["wind_table"], self.dt, stochastic_wind,
                                   horiontal_wind_percentile, noise=wind_noise)
        else:
            self.wind = None
        if flight_phase in (PHASE_S, PHASE_U, PHASE_B, PHASE_F):
            self.state_initial = list(self.T.initial_states[flight_phase])
        else:
            self.state_initial = list(self.T.initial_state)
        # the landing closures take y_0 / mass_0 from load_landing_burn_initial_state()
        self.y_0 = self.T.initial_state[1]
        self.mass_0 = self.T.initial_state[8]
        if self.ascent:
            hyper = SUBSONIC_HYPER if flight_phase == PHASE_S else SUPERSONIC_HYPER
            cols = list(zip(*hyper))
            f = lambda k: interp1d(cols[0], cols[k], kind="linear", fill_value="extrapolate")
            (self.f_max_x, self.f_max_vy, self.f_max_vx, self.f_max_alpha, self.f_w_alpha,
             self.f_w_x, self.f_w_vy, self.f_w_vx) = [f(k) for k in range(1, 9)]
            if flight_phase == PHASE_S:
                self.terminal_mach = 1.0
            else:       # rtd_rl.py:582-588
                xt, yt, vxt, vyt, mt = self.T.other["ref_traj_ascent_terminal"]
                _, _, a_t = isa(yt)
                self.terminal_mach = math.sqrt(vxt ** 2 + vyt ** 2) / a_t
        self.truncation_id = 0
        self.last = {}
        self.reset()

    # ------------------------------------------------------------------
    def reset(self):
        self.state = self.state_initial
        self.previous_state = self.state
        self.truncation_id = 0
        self.gimbal_prev = 0.0
        self.delta_l_prev = 0.0
        self.delta_r_prev = 0.0
        if self.enable_wind:
            self.wind.reset()
        self.g_window = []
        return self.state

    def set_state(self, state, g_window=(), gimbal_prev=0.0, delta_l_prev=0.0, delta_r_prev=0.0):
        self.state = [np.float64(v) for v in state]
        self.previous_state = self.state
        self.g_window = list(g_window)
        self.gimbal_prev, self.delta_l_prev, self.delta_r_prev = gimbal_prev, delta_l_prev, delta_r_prev

    # ------------------------------------------------------------------
    def _control_P(self, actions, p_atm, theta, alpha_eff, q, x_cog, mach):
        p = self.T.p
        u0 = _unpack_action_P(actions)
        throttle = (u0 + 1) / 2 * (1 - self.nominal_throttle) + self.nominal_throttle
        t_full = p["thrust_per_engine"] + (p["nozzle_exit_pressure"] - p_atm) * p["nozzle_exit_area"]
        thrust = t_full * self.n_eng * throttle
        n_tot = thrust / t_full
        mass_flow = (p["thrust_per_engine"] / p["v_exhaust"]) * n_tot
        f_perp, f_par, m_z, _, _ = self.T.acs(alpha_eff, theta, q, mach, x_cog, 0.0, 0.0, 0.0, 0.0,
                                              self.dt_act)
        return thrust + f_par, f_perp, m_z, mass_flow, throttle, None

    def _control_G(self, actions, p_atm, d_thrust_cg, theta, alpha_eff, q, x_cog, mach):
        p = self.T.p
        u0, u1, u2, u3 = _unpack_action_G(actions)
        max_gimbal_rad = math.radians(5)
        max_defl_rad = math.radians(20)
        gimbal_rad = u0 * max_gimbal_rad
        max_gimbal_deg = math.degrees(max_gimbal_rad)
        x = self.gimbal_prev
        gimbal_deg = x + self.dt_act * ((-x + math.degrees(gimbal_rad)) / 1.0)
        gimbal_deg = np.clip(gimbal_deg, -max_gimbal_deg, max_gimbal_deg)
        gimbal_rad = math.radians(gimbal_deg)
        throttle = (u1 + 1) / 2 * (1 - self.nominal_throttle) + self.nominal_throttle
        t_full = p["thrust_per_engine"] + (p["nozzle_exit_pressure"] - p_atm) * p["nozzle_exit_area"]
        thrust_g = t_full * self.n_eng * throttle
        t_par = thrust_g * math.cos(gimbal_rad)
        t_perp = -thrust_g * math.sin(gimbal_rad)
        m_z = -thrust_g * math.sin(gimbal_rad) * d_thrust_cg
        total = np.sqrt(t_par ** 2 + t_perp ** 2)
        n_tot = total / t_full
        mass_flow = (p["thrust_per_engine"] / p["v_exhaust"]) * n_tot
        gimbal_deg = math.degrees(gimbal_rad)
        cmd_l = u2 * max_defl_rad      # named "_deg" upstream, is radians-scaled
        cmd_r = u3 * max_defl_rad
        f_perp, f_par, a_mz, d_cmd_l, d_cmd_r = self.T.acs(
            alpha_eff, theta, q, mach, x_cog, cmd_l, cmd_r, self.delta_l_prev,
            self.delta_r_prev, self.dt_act)
        return (t_par + f_par, t_perp + f_perp, m_z + a_mz, mass_flow, throttle,
                (gimbal_deg, d_cmd_l, d_cmd_r))

    def _control_ascent(self, actions, p_atm, d_thrust_cg):
        """force_moment_decomposer_ascent, rockets_physics.py:17-56 (gimbal +-7 deg, nominal 0.5)."""
        p = self.T.p
        u0, u1 = actions
        gimbal_rad = u0 * math.radians(7.0)
        nn = (u1 + 1) / 2
        throttle = nn * (1 - self.nominal_throttle) + self.nominal_throttle
        t_full = p["thrust_per_engine"] + (p["nozzle_exit_pressure"] - p_atm) * p["nozzle_exit_area"]
        n_g = int(p["n_engines_gimballed"])
        n_ng = int(self.T.other["n_engines_stage1"]) - n_g
        thrust_g = t_full * n_g * throttle
        thrust_ng = t_full * n_ng * throttle
        t_par = thrust_ng + thrust_g * math.cos(gimbal_rad)
        t_perp = -thrust_g * math.sin(gimbal_rad)
        m_z = -thrust_g * math.sin(gimbal_rad) * d_thrust_cg
        total = np.sqrt(t_par ** 2 + t_perp ** 2)
        n_tot = total / t_full
        mass_flow = (p["thrust_per_engine"] / p["v_exhaust"]) * n_tot
        return t_par, t_perp, m_z, mass_flow, throttle, None

    def _control_flip(self, action, p_atm, d_thrust_cg):
        """force_moment_decomposer_flipoverboostbackburn, rockets_physics.py:63-92 (max gimbal 10 deg,
        first-order low-pass tau 1.0 on the env dt, throttle 1, the gimballed engines only)."""
        p = self.T.p
        cmd_deg = action * 10
        x = self.gimbal_prev
        gimbal_deg = x + self.dt_act * ((-x + cmd_deg) / 1.0)
        gimbal_rad = math.radians(gimbal_deg)
        throttle = 1
        t_full = p["thrust_per_engine"] + (p["nozzle_exit_pressure"] - p_atm) * p["nozzle_exit_area"]
        thrust = t_full * int(p["n_engines_gimballed"]) * throttle
        t_par = thrust * math.cos(gimbal_rad)
        t_perp = -thrust * math.sin(gimbal_rad)
        m_z = -thrust * math.sin(gimbal_rad) * d_thrust_cg
        total = np.sqrt(t_par ** 2 + t_perp ** 2)
        n_tot = total / t_full
        mass_flow = (p["thrust_per_engine"] / p["v_exhaust"]) * n_tot
        return t_par, t_perp, m_z, mass_flow, None, (gimbal_deg, 0.0, 0.0)

    def _control_rcs(self, action, x_cog):
        """RCS, rockets_physics.py:149-166."""
        o = self.T.other
        thruster_force = o["max_rcs_force_per_thruster"] * action
        m_z = (-thruster_force * (x_cog - o["d_base_rcs_bottom"])
               + thruster_force * (o["d_base_rcs_top"] - x_cog))
        if not (type(m_z) == np.float64 or type(m_z) == float):
            m_z = m_z[0]
        return 0, 0, m_z, 0, None, None

    def _control_C(self, actions_v_ref, p_atm, theta, alpha_eff, q, x_cog, mach, speed):
        """force_moment_decomposer_landing_burn_throttle_PID, rockets_physics.py:402-451."""
        if actions_v_ref.ndim == 2:
            v_ref = actions_v_ref[0][0]
        else:
            v_ref = actions_v_ref[0]
        error = v_ref - speed
        nn = np.clip(error * -0.08, 0, 1)
        return self._control_P([2 * (nn - 0.5)], p_atm, theta, alpha_eff, q, x_cog, mach)

    def substep(self, state, actions):
        T = self.T
        x, y, vx, vy, theta, theta_dot, gamma, alpha, mass, m_prop, time = state
        rho, p_atm, a_snd = isa(y)
        speed = math.sqrt(vx ** 2 + vy ** 2)
        if a_snd != 0.0:
            mach = min(speed / a_snd, 10.0)
        else:
            mach = 0.0
        q = 0.5 * rho * speed ** 2
        fuel_frac = (T.m_prop0 - m_prop) / T.m_prop0
        if fuel_frac == 0.0:
            fuel_frac = 1e-6
        if self.ascent:
            x_cog, inertia = T.cog_inertia_full(1 - fuel_frac)
            d_thrust_cg = x_cog + T.other["engine_height_full"]
        else:
            x_cog, inertia = T.cog_inertia(1 - fuel_frac)
            d_thrust_cg = x_cog + T.p["engine_height"]
        if vy < 0:
            alpha_eff = gamma - theta - math.pi
        else:
            alpha_eff = alpha
        d_cp_cg = x_cog - (T.cop_full if self.ascent else T.cop)
        if self.wind is not None:
            ug, vg = self.wind(y)
        else:
            ug, vg = 0.0, 0.0
        S = T.p["frontal_area"]
        f_wind_x = 0.5 * rho * ug ** 2 * S * T.c_gust_x
        f_wind_y = 0.5 * rho * vg ** 2 * S * T.c_gust_y
        m_wind_z = -d_cp_cg * f_wind_y
        if a_snd != 0.0:
            C_L = T.cl(mach, alpha_eff)
            C_D = T.cd(mach, alpha_eff)
        else:
            C_L = 0.0
            C_D = 0.0
        drag = 0.5 * rho * speed ** 2 * C_D * S
        lift = 0.5 * rho * speed ** 2 * C_L * S
        if vy >= 0.0:
            a_par = lift * math.sin(alpha_eff) - drag * math.cos(alpha_eff)
            a_perp = -lift * math.cos(alpha_eff) - drag * math.sin(alpha_eff)
        else:
            a_par = drag * math.cos(alpha_eff) - lift * math.sin(alpha_eff)
            a_perp = -drag * math.sin(alpha_eff) - lift * math.cos(alpha_eff)
        aero_x = a_par * math.cos(theta) + a_perp * math.sin(theta)
        aero_y = a_par * math.sin(theta) - a_perp * math.cos(theta)
        aero_mz = a_perp * d_cp_cg
        if self.flight_phase == PHASE_P:
            c_par, c_perp, c_mz, mass_flow, throttle, act = self._control_P(
                actions, p_atm, theta, alpha_eff, q, x_cog, mach)
        elif self.flight_phase == PHASE_G:
            c_par, c_perp, c_mz, mass_flow, throttle, act = self._control_G(
                actions, p_atm, d_thrust_cg, theta, alpha_eff, q, x_cog, mach)
        elif self.ascent:
            c_par, c_perp, c_mz, mass_flow, throttle, act = self._control_ascent(
                actions, p_atm, d_thrust_cg)
        elif self.flight_phase == PHASE_B:
            c_par, c_perp, c_mz, mass_flow, throttle, act = self._control_rcs(actions, x_cog)
        elif self.flight_phase == PHASE_F:
            c_par, c_perp, c_mz, mass_flow, throttle, act = self._control_flip(actions, p_atm, d_thrust_cg)
            # "No aerodynamic forces in upper atmosphere, this is a redundancy." (:556-560)
            aero_x = 0.0
            aero_y = 0.0
            aero_mz = 0.0
        else:
            c_par, c_perp, c_mz, mass_flow, throttle, act = self._control_C(
                actions, p_atm, theta, alpha_eff, q, x_cog, mach, speed)
        # NaN guards: an if/elif chain, only the first NaN is cleared
        if math.isnan(c_par):
            c_par = 0.0
        elif math.isnan(c_perp):
            c_perp = 0.0
        elif math.isnan(c_mz):
            c_mz = 0.0
        c_x = c_par * math.cos(theta) + c_perp * math.sin(theta)
        c_y = c_par * math.sin(theta) - c_perp * math.cos(theta)
        g = gravity(y)
        fx = aero_x + c_x + f_wind_x
        fy = aero_y + c_y + f_wind_y
        dt = self.dt_phys
        vx_dot = fx / mass
        vy_dot = fy / mass - g
        vx += vx_dot * dt
        vy += vy_dot * dt
        x += vx * dt
        y += vy * dt
        mz = c_mz + aero_mz + m_wind_z
        theta_dot += (mz / inertia) * dt
        theta += theta_dot * dt
        gamma = math.atan2(vy, vx)
        if theta > 2 * math.pi:
            theta -= 2 * math.pi
        if gamma < 0:
            gamma = 2 * math.pi + gamma
        alpha = theta - gamma
        m_prop -= mass_flow * dt
        mass -= mass_flow * dt
        time += dt
        info = dict(mach=mach, q=q, CL=C_L, CD=C_D, rho=rho, p_atm=p_atm, a=a_snd,
                    x_cog=x_cog, inertia=inertia, mass_flow=mass_flow, throttle=throttle,
                    alpha_eff=alpha_eff, drag=drag, lift=lift, ug=ug, vg=vg,
                    c_par=c_par, c_perp=c_perp, c_mz=c_mz, act=act)
        return [x, y, vx, vy, theta, theta_dot, gamma, alpha, mass, m_prop, time], info

    def step(self, actions):
        state = self.state
        for _ in range(self.n_sub):
            state, info = self.substep(state, actions)
        self.state = state
        if self.flight_phase in (PHASE_G, PHASE_F):
            # only the 4th sub-step's actuator outputs are fed back, and the deltas
            # fed back are the *commands*
            self.gimbal_prev, self.delta_l_prev, self.delta_r_prev = info["act"]
        x, y, vx, vy = state[:4]
        vxp, vyp = self.previous_state[2], self.previous_state[3]
        v = math.sqrt(vx ** 2 + vy ** 2)
        v_p = math.sqrt(vxp ** 2 + vyp ** 2)
        g_load = abs(v - v_p) / self.dt * 1 / 9.81
        if len(self.g_window) < 10:
            self.g_window.append(g_load)
        else:
            self.g_window.pop(0)
            self.g_window.append(g_load)
        info["g_load_1_sec_window"] = sum(self.g_window) / 10
        truncated, self.truncation_id = self._truncated(state, info)
        done = self._done(state)
        reward = self._reward(state, done, truncated, actions, info)
        self.previous_state = state
        self.last = info
        return state, reward, done, truncated, info

    # -- reward / truncation / done ---------------------------------------
    @staticmethod
    def _mach_rtd(y, vx, vy):
        _, _, a = isa(y)
        speed = math.sqrt(vx ** 2 + vy ** 2)
        return speed / a if (speed != 0 and a != 0) else 0

    # -- type 'supervisory': src/envs/supervisory/rtd_supervisory_mock.py:6-104 (reward 0)
    def _sup_done(self, s):
        x, y, vx, vy, theta, theta_dot, gamma = s[:7]
        rho, _, a = isa(y)
        speed = math.sqrt(vx ** 2 + vy ** 2)
        if self.flight_phase == PHASE_S:
            return bool(speed / a > 1.1)
        if self.flight_phase == PHASE_U:
            return bool(y > self.T.other["ref_traj_ascent_terminal"][1])
        if self.flight_phase == PHASE_B:
            return bool(0.5 * rho * speed ** 2 > 65000 and abs(gamma - theta - math.pi) < math.radians(3))
        if self.flight_phase == PHASE_F:
            return bool(vx < -60)
        return bool(y < 1)

    def _sup_truncated(self, s, info):
        x, y, vx, vy, theta, theta_dot, gamma, alpha, mass, m_prop, time = s
        rho, _, _ = isa(y)
        speed = math.sqrt(vx ** 2 + vy ** 2)
        q = 0.5 * rho * speed ** 2
        if self.ascent or self.flight_phase == PHASE_F:
            return (True, 1) if m_prop <= 0 else (False, 0)
        if self.flight_phase == PHASE_B:
            return (True, 1) if (q > 35000 and abs(gamma - theta - math.pi) > math.radians(3)) else (False, 0)
        if y < -10:
            return True, 1
        elif m_prop <= 0:
            return True, 2
        elif theta > math.pi + math.radians(2):
            return True, 3
        elif q > 65000:
            return True, 4
        elif info["g_load_1_sec_window"] > 6.0:
            return True, 5      # upstream's print in this branch raises NameError ('acceleration')
        elif vy > 0.0:
            return True, 6
        return False, 0

    def _done(self, s):
        x, y, vx, vy = s[:4]
        speed = math.sqrt(vx ** 2 + vy ** 2)
        if self.type == "supervisory":
            return self._sup_done(s)
        if self.ascent:                     # rtd_rl.py:35-51
            if any(math.isnan(v) for v in s):
                return False
            return bool(s[9] >= 0 and self._mach_rtd(y, vx, vy) > self.terminal_mach)
        if self.flight_phase == PHASE_B:    # rtd_rl.py:148-158
            rho, _, _ = isa(y)
            q = 0.5 * rho * speed ** 2
            return bool(q > 10000 and abs(s[6] - s[4] - math.pi) < math.radians(3))
        if self.flight_phase == PHASE_C:    # rtd_rl.py:356-367
            return bool(y > 0 and y < 5 and speed < 1)
        if self.type == "pso" and self.flight_phase == PHASE_G:
            dist = math.sqrt(x ** 2 + y ** 2)
            return bool(dist > 0 and dist < 1 and speed < 2.5)
        thr = 5.5 if self.type == "pso" else 5.0
        return bool(y > 0 and y < 1 and speed < thr)

    def _truncated(self, s, info):
        x, y, vx, vy, theta, theta_dot, gamma, alpha, mass, m_prop, time = s
        rho, _, _ = isa(y)
        speed = math.sqrt(vx ** 2 + vy ** 2)
        q = 0.5 * rho * speed ** 2
        g1 = info["g_load_1_sec_window"]
        if vy < 0:
            a_eff = abs(gamma - theta - math.pi)
        else:
            a_eff = abs(theta - gamma)
        if self.type == "supervisory":
            return self._sup_truncated(s, info)
        if self.ascent:                     # rtd_rl.py:53-88
            if any(math.isnan(v) for v in s):
                return True, 0
            xr, vxr, vyr = self.T.ref_x(y), self.T.ref_vx(y), self.T.ref_vy(y)
            mach = self._mach_rtd(y, vx, vy)
            if m_prop <= 0:
                return True, 1
            elif mach > self.terminal_mach + 0.09:
                return True, 2
            elif abs(x - xr) > self.f_max_x(mach):
                return True, 3
            elif y < 0:
                return True, 4
            elif abs(alpha) > math.radians(self.f_max_alpha(mach)):
                return True, 5
            elif abs(vx - vxr) > self.f_max_vx(mach):
                return True, 6
            elif abs(vy - vyr) > self.f_max_vy(mach):
                return True, 7
            return False, 0
        if self.flight_phase == PHASE_B:    # rtd_rl.py:160-170
            if q > 10000 - 2000 and abs(gamma - theta - math.pi) > math.radians(5):
                return True, 1
            return False, 0
        if self.flight_phase == PHASE_C:    # rtd_rl.py:369-401
            if y < -10:
                return True, 1
            elif m_prop <= 0:
                return True, 2
            elif theta > math.pi + math.radians(2):
                return True, 3
            elif q > 65000:
                return True, 4
            elif g1 > 6.0:
                return True, 5
            elif vy > 0.0:
                return True, 6
            return False, 0
        if self.type == "pso" and self.flight_phase == PHASE_P:
            if y < 0.0:
                return True, 1
            elif m_prop <= 0:
                return True, 2
            elif theta > math.pi + math.radians(2):
                return True, 3
            elif q > 65000:
                return True, 4
            elif vy > 0.0:
                return True, 6
            elif g1 > 6.0:
                return True, 7
            return False, 0
        if self.type == "pso":
            over = self._overshoot(x, y)
            if over > 0.5:
                return True, 1
            elif m_prop <= 0:
                return True, 2
            elif a_eff > math.radians(10):
                return True, 3
            elif q > 65000:
                return True, 4
            elif vy > 0.0:
                return True, 6
            elif g1 > 6.0:
                return True, 7
            elif y > 1000 and vx > 0.0:
                return True, 8
            return False, 0
        # rl (both landing phases share the closure, rtd_rl.py:208-240)
        if y < -10:
            return True, 1
        elif m_prop <= 0:
            return True, 2
        elif theta > math.pi + math.radians(2):
            return True, 3
        elif q > 65000:
            return True, 4
        elif g1 > 6.0:
            return True, 5
        elif vy > 0.0:
            return True, 6
        elif vx > 0.01:
            return True, 7
        return False, 0

    @staticmethod
    def _overshoot(x, y):
        if x < 0 and y < 0:
            return math.sqrt(x ** 2 + y ** 2)
        elif x < 0:
            return -x
        elif y < 0:
            return -y
        return 0

    def _reward(self, s, done, truncated, actions, info):
        x, y, vx, vy, theta, theta_dot, gamma, alpha, mass, m_prop, time = s
        speed = math.sqrt(vx ** 2 + vy ** 2)
        if self.type == "supervisory":
            return 0
        if self.type == "pso":
            reward = 0
            if self.flight_phase == PHASE_P:
                if truncated and y > 0:
                    reward = -abs(y)
                elif truncated and y < 0:
                    reward = 200 - abs(speed)
                elif done:
                    reward = m_prop
            else:
                over = self._overshoot(x, y)
                dist = math.sqrt(x ** 2 + y ** 2)
                if truncated and over < 0.5:
                    reward = -abs(dist)
                elif truncated:
                    reward = 200 - abs(speed)
                elif done:
                    reward = m_prop
            return reward
        if self.ascent:                     # rtd_rl.py:90-120
            if any(math.isnan(v) for v in s):
                return 0
            mach = self._mach_rtd(y, vx, vy)
            reward = 0
            xr, vxr, vyr = self.T.ref_x(y), self.T.ref_vx(y), self.T.ref_vy(y)
            if y < 0:
                return 0
            reward += math.exp(-4 * (vx - vxr) ** 2 / self.f_max_vx(mach) ** 2) * self.f_w_vx(mach)
            reward += math.exp(-4 * (vy - vyr) ** 2 / self.f_max_vy(mach) ** 2) * self.f_w_vy(mach)
            reward += math.exp(-4 * (x - xr) ** 2 / self.f_max_x(mach) ** 2) * self.f_w_x(mach)
            reward += math.exp(-4 * math.degrees(alpha) ** 2 / self.f_max_alpha(mach) ** 2) * self.f_w_alpha(mach)
            if done:
                reward += 2.5
            reward /= 10 ** 4
            return reward
        if self.flight_phase == PHASE_B:    # rtd_rl.py:172-181
            reward = (math.pi - abs(gamma - theta - math.pi)) / math.pi
            if done:
                reward += 3.5
            reward /= 100
            return reward
        rho, _, _ = isa(y)
        if self.flight_phase == PHASE_C:    # rtd_rl.py:478-531 (the second definition wins)
            speed = math.hypot(vx, vy)
            q = 0.5 * rho * speed ** 2
            v_ref = actions[0][0] if actions.ndim == 2 else actions[0]
            reward = 0.0
            if q > 60_000.0:
                q_ex = (q - 60_000.0) / (65_000.0 - 60_000.0)
                reward -= 1.0 * min(q_ex ** 2, 1.0)
            g1 = info["g_load_1_sec_window"]
            if g1 > 5.5:
                g_ex = (g1 - 5.5) / (6.0 - 5.5)
                reward -= 1.0 * min(g_ex ** 2, 1.0)
            prog = (self.y_0 - y) / self.y_0
            vel_tracking = max(0.0, 1.0 - abs(speed - v_ref) / 10.0)
            w_prog = 0.5 if (q <= 60_000.0 and g1 <= 5.5) else 0.5 * 0.1
            reward += w_prog * prog * vel_tracking
            if y < 100.0:
                reward += 0.5 * max(0.0, 1.0 - abs(vy - 0.0) / 50.0)
            reward += 0.01 * (1 - self.discount_factor)
            if done and not truncated:
                reward += 5.0
                mass_used = self.y_0 * 0.0 + (self.mass_0 - mass)
                reward -= min(0.1 * mass_used, 1.0)
            elif truncated:
                reward -= min(4.0 * (y / self.y_0) * (abs(vy) / 100.0), 5.0)
            return np.clip(reward, -10.0, 10.0)
        if self.flight_phase == PHASE_P:
            speed = math.hypot(vx, vy)
            q = 0.5 * rho * speed ** 2
            reward = 0.0
            if q > 60_000.0:
                q_ex = (q - 60_000.0) / (65_000.0 - 60_000.0)
                reward -= 1.0 * min(q_ex ** 2, 1.0)
            g1 = info["g_load_1_sec_window"]
            if g1 > 5.5:
                g_ex = (g1 - 5.5) / (6.0 - 5.5)
                reward -= 1.0 * min(g_ex ** 2, 1.0)
            prog = (self.y_0 - y) / self.y_0
            w_prog = 0.5 if (q <= 60_000.0 and g1 <= 5.5) else 0.5 * 0.1
            reward += w_prog * prog
            if y < 100.0:
                reward += 5.5 * (1.0 - abs(vy) / 50.0)
            if done and not truncated:
                reward += 400.0 * m_prop / self.mass_0
            elif truncated and y > 0:
                reward -= 50.0 * (abs(y) / self.y_0)
            elif truncated and y < 0:
                reward -= 50.0 * (abs(vy) / 10)
            if not done or not (truncated and y < 0):
                reward = np.clip(reward, -10.0, 10.0)
            return reward
        # gimballed landing burn, rtd_rl.py:243-269
        a_eff = abs(gamma - theta - math.pi)
        reward = 0
        u0 = actions[0][0] if actions.ndim == 2 else actions[0]
        tau = (u0 + 1) / 2
        reward += (1.5 - math.log(1 + a_eff) / (math.log(1 + math.radians(20))) - tau * 0.5) \
            * (1 - y / self.y_0) * 2 / 3
        if y < 100:
            reward += 1 - math.tanh((speed - 15) / 15)
        if truncated and y < 5:
            reward += 1 - math.tanh((speed - 5) / 5)
        if done:
            reward += 5
        reward *= (1 - self.discount_factor) / (1 - self.discount_factor ** self.trajectory_length)
        return reward


# ---------------------------------------------------------------------------
class PsoModel:
    """pso_wrapped_env: per-particle torch MLP + episode loop -> fitness."""

    def __init__(self, flight_phase=PHASE_P, enable_wind=False, stochastic_wind=False,
                 horiontal_wind_percentile=50, tables=None, wind_noise=None, fast_rbf=False,
                 max_steps=None):
        import torch
        import torch.nn as nn
        self.torch = torch
        self.env = OracleEnv(flight_phase, "pso", enable_wind, stochastic_wind,
                             horiontal_wind_percentile, tables, wind_noise, fast_rbf)
        self.flight_phase = flight_phase
        if flight_phase == PHASE_P:
            i, o, n, h = 2, 1, 3, 8
        else:
            i, o, n, h = 5, 4, 4, 8
        self.net = nn.Sequential(
            nn.Linear(i, h), nn.ReLU(),
            *[nn.Sequential(nn.Linear(h, h), nn.ReLU()) for _ in range(n)],
            nn.Linear(h, o), nn.Tanh())
        self.n_params = sum(p.numel() for p in self.net.parameters())
        self.bounds = [(-1.5, 1.5)] * self.n_params
        self.max_steps = max_steps
        self.steps = 0

    def set_weights(self, individual):
        torch = self.torch
        k = 0
        for _, prm in self.net.named_parameters():
            n = prm.numel()
            prm.data = torch.tensor(individual[k:k + n], dtype=torch.float32).view(prm.shape)
            k += n

    def obs(self, state):
        x, y, vx, vy, theta = state[:5]
        nv = self.env.T.norm_vals
        if self.flight_phase == PHASE_P:
            return np.array([y / nv[0], vy / nv[1]])
        k_theta = float(np.arctanh(0.75) / math.radians(25))
        return np.array([x / nv[-2], y / nv[0], vx / nv[-1], vy / nv[1],
                         math.tanh(k_theta * (theta - math.pi / 2))])

    def act(self, obs):
        torch = self.torch
        with torch.no_grad():
            return self.net(torch.tensor(obs, dtype=torch.float32))

    def objective_function(self, individual, record=None):
        self.set_weights(individual)
        obs = self.obs(self.env.reset())
        total = 0
        steps = 0
        while True:
            a = self.act(obs).detach().numpy()          # float32 1-D
            state, reward, done, truncated, info = self.env.step(a)
            obs = self.obs(state)
            total -= reward
            steps += 1
            if record is not None:
                record.append((list(map(float, state)), a.copy(), float(reward)))
            if done or truncated:
                break
            if self.max_steps is not None and steps >= self.max_steps:
                break
        self.steps = steps
        return total


class RlEnv:
    """rl_wrapped_env_pytorch for P and G: fp32-rounded obs, G action log-compression."""

    def __init__(self, flight_phase=PHASE_P, enable_wind=False, stochastic_wind=True,
                 horiontal_wind_percentile=50, tables=None, wind_noise=None, fast_rbf=False,
                 trajectory_length=1, discount_factor=0.99):
        self.env = OracleEnv(flight_phase, "rl", enable_wind, stochastic_wind,
                             horiontal_wind_percentile, tables, wind_noise, fast_rbf,
                             trajectory_length, discount_factor)
        self.flight_phase = flight_phase
        self.state_dim, self.action_dim = {PHASE_P: (2, 1), PHASE_G: (5, 4), PHASE_S: (8, 2),
                                           PHASE_U: (8, 2), PHASE_B: (4, 1), PHASE_C: (1, 1)}[flight_phase]
        if flight_phase == PHASE_C:
            vx0, vy0 = self.env.T.initial_state[2], self.env.T.initial_state[3]
            self.speed0 = math.sqrt(vx0 ** 2 + vy0 ** 2)

    def augment_action(self, a):
        if self.flight_phase == PHASE_C:        # env_wrapped_rl_pytorch.py:158-164
            u0 = a[0] if a.ndim == 2 else a
            return np.array([(u0 + 1) / 2 * self.speed0])
        if self.flight_phase != PHASE_G:
            return a
        u0, u1, u2, u3 = a[0] if a.ndim == 2 else a
        f = lambda u, c: math.copysign(math.log(1 + c * abs(u)) / math.log(1 + c), u)
        out = [f(u0, 10), u1, f(u2, 5), f(u3, 5)]
        return np.array([out]) if a.ndim == 2 else np.array(out)

    def obs(self, state):
        s = np.asarray(state, dtype=np.float32).reshape(-1)
        x, y, vx, vy, theta, theta_dot, gamma = s[:7]
        nv = self.env.T.norm_vals
        if self.flight_phase in (PHASE_S, PHASE_U, PHASE_B):    # env_wrapped_rl_pytorch.py:169-177
            alpha, mass = s[7], s[8]
            if self.flight_phase == PHASE_B:
                o = np.array([theta, theta_dot, gamma, alpha])
            else:
                o = np.array([x, y, vx, vy, theta, theta_dot, alpha, mass])
            o /= np.array(self.env.T.other["norm_vals"][self.flight_phase])
            return o
        if self.flight_phase == PHASE_C:
            return np.array([(1 - y / nv[0]) * 2 - 1])
        if self.flight_phase == PHASE_P:
            return np.array([(1 - y / nv[0]) * 2 - 1, (1 - vy / nv[1]) * 2 - 1])
        k = float(np.arctanh(0.75) / math.radians(5))
        kd = float(np.arctanh(0.75) / 0.01)
        return np.array([y / nv[0], vy / nv[1], math.tanh(k * (theta - math.pi / 2)),
                         math.tanh(kd * theta_dot), math.tanh(k * (gamma - 3 / 2 * math.pi))])

    def reset(self):
        return self.obs(self.env.reset())

    def step(self, action):
        a = action if isinstance(action, np.ndarray) else np.array(action)
        a = self.augment_action(a)
        if a.ndim == 2:
            a = a[0]
        state, reward, done, truncated, info = self.env.step(a)
        return self.obs(state), float(reward), bool(done), bool(truncated), info


class SupervisoryEnv:
    """supervisory_wrapper (src/envs/supervisory/env_wrapped_supervisory.py:6-116): base env with the
    supervisory closures, float64 observation over the caller's scales, wind always off."""

    def __init__(self, input_normalisation_values, flight_phase=PHASE_S, tables=None, fast_rbf=False):
        self.env = OracleEnv(flight_phase, "supervisory", False, False, 95, tables, None, fast_rbf)
        self.flight_phase = flight_phase
        self.nv = input_normalisation_values[:2] if flight_phase == PHASE_P else input_normalisation_values
        self._rl = RlEnv.__new__(RlEnv)              # borrows augment_action (same shaping, :58-104)
        self._rl.flight_phase = flight_phase
        if flight_phase == PHASE_C:
            vx0, vy0 = self.env.T.initial_state[2], self.env.T.initial_state[3]
            self._rl.speed0 = math.sqrt(vx0 ** 2 + vy0 ** 2)

    def obs(self, state):                            # :35-56
        x, y, vx, vy, theta, theta_dot, gamma, alpha, mass = state[:9]
        if self.flight_phase in (PHASE_S, PHASE_U, PHASE_G):
            return np.array([x, y, vx, vy, theta, theta_dot, alpha, mass]) / self.nv
        if self.flight_phase == PHASE_B:
            return np.array([theta, theta_dot, gamma, alpha]) / self.nv
        if self.flight_phase == PHASE_F:
            return np.array([theta, theta_dot]) / self.nv
        if self.flight_phase == PHASE_P:
            return np.array([(1 - y / self.nv[0]) * 2 - 1, (1 - vy / self.nv[1]) * 2 - 1])
        return np.array([(1 - y / self.nv[0]) * 2 - 1])

    def reset(self):
        return self.obs(self.env.reset())

    def step(self, action):                          # :106-111 (no ndim squeeze here)
        state, reward, done, truncated, info = self.env.step(self._rl.augment_action(np.array(action)))
        return self.obs(state), reward, done, truncated, info


def classical_rollout(tables=None, fast_rbf=False, max_steps=50000):
    """LandingBurn(test_case='control').run_closed_loop(): P controller on v_ref(y),
    physics only (no rtd), float64 2-D action."""
    env = OracleEnv(PHASE_P, "pso", tables=tables, fast_rbf=fast_rbf)
    a_opt, b_opt = env.T.p["v_opt_a"], env.T.p["v_opt_b"]
    state = list(env.state_initial)
    x, y, vx, vy, theta, theta_dot, gamma, alpha, mass, m_prop, time = state
    speed = np.sqrt(vx ** 2 + vy ** 2)
    q = 0.0
    alpha_eff = gamma - theta - math.pi
    n = 0
    rows = []
    while m_prop > 0 and y > 1 and q < 65e3 and n < max_steps and vy < 0 \
            and alpha_eff < math.degrees(5):
        v_ref = a_opt * y ** 2 + b_opt * y
        nn_thr = np.clip(-0.10 * (v_ref - speed) + 0.0 * 0.0, 0.0, 1.0)
        u0 = 2 * (nn_thr - 0.5)
        for _ in range(4):
            state, info = env.substep(state, np.array([[u0]]))
        x, y, vx, vy, theta, theta_dot, gamma, alpha, mass, m_prop, time = state
        speed = np.sqrt(vx ** 2 + vy ** 2)
        alpha_eff = gamma - theta - math.pi
        q = info["q"]
        rows.append(list(map(float, state)) + [float(u0)])
        n += 1
    return n, rows
