"""ctypes binding of libpd_b200.so (include/pd_b200.h).  No CPU fallback: if the shared
library is missing or no CUDA device is present, construction fails loudly."""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from .params import RocketParams
from . import rbf_sets

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PD_LIB_PATH") or os.path.join(HERE, "libpd_b200.so")
CACHE_DIR = os.path.join(HERE, "_cache")

PHASES = {"landing_burn_pure_throttle": 0, "landing_burn": 1, "subsonic": 2, "supersonic": 3,
          "ballistic_arc_descent": 4, "landing_burn_pure_throttle_Pcontrol": 5, "flip_over_boostbackburn": 6}
RL_ONLY_PHASES = ("subsonic", "supersonic", "ballistic_arc_descent", "landing_burn_pure_throttle_Pcontrol")
SUPERVISORY_ONLY_PHASES = ("flip_over_boostbackburn",)
RTD = {"pso": 0, "rl": 1, "supervisory": 2}
PRECISION = {"fp64": 0, "fp32": 1}
OBS_DIM = {0: 2, 1: 5, 2: 8, 3: 8, 4: 4, 5: 1, 6: 2}
ACT_DIM = {0: 1, 1: 4, 2: 2, 3: 2, 4: 1, 5: 1, 6: 1}
N_PARAMS = {0: 249, 1: 372}
INERTIA_FULL_ORDER = ("m_s_1", "x_dry_1", "I_dry_1", "m_2", "m_pay", "x_wet_2_initial", "I_wet_2_initial",
                      "h_1", "h_1_ox", "h_1_f", "m_1_ox", "m_1_f", "h_lower_1")


class PdRbfGrid(C.Structure):
    _fields_ = [("m0", C.c_double), ("dm", C.c_double), ("a0", C.c_double), ("da", C.c_double),
                ("nm", C.c_int32), ("na", C.c_int32), ("n_impure", C.c_int32), ("_pad", C.c_int32),
                ("cells", C.c_void_p), ("imp_hint", C.c_void_p), ("imp_id", C.c_void_p)]


class PdRbfTable(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("n_points", C.c_int32), ("n_sets", C.c_int32),
                ("hash_size", C.c_int32), ("levels", C.c_double * 5), ("level_off", C.c_int32 * 6),
                ("n_grids", C.c_int32), ("mach_sorted", C.c_void_p), ("points", C.c_void_p),
                ("rows", C.c_void_p), ("hash_keys", C.c_void_p), ("hash_vals", C.c_void_p),
                ("grids", PdRbfGrid * 2)]


class PdOtherPhases(C.Structure):
    _fields_ = [("n_engines_stage1", C.c_int32), ("n_ref", C.c_int32),
                ("max_rcs_force_per_thruster", C.c_double), ("d_base_rcs_bottom", C.c_double),
                ("d_base_rcs_top", C.c_double), ("inertia_full", C.c_double * 13),
                ("engine_height_full", C.c_double), ("cop_full", C.c_double),
                ("initial_state", (C.c_double * 11) * 4), ("norm_vals", (C.c_double * 8) * 4),
                ("ref_y", C.c_void_p), ("ref_x", C.c_void_p), ("ref_vx", C.c_void_p), ("ref_vy", C.c_void_p),
                ("ref_terminal", C.c_double * 5)]


class PdParams(C.Structure):
    _fields_ = [("thrust_per_engine", C.c_double), ("nozzle_exit_pressure", C.c_double),
                ("nozzle_exit_area", C.c_double), ("v_exhaust", C.c_double),
                ("n_engines_gimballed", C.c_int32), ("_pad0", C.c_int32),
                ("grid_fin_area", C.c_double), ("d_base_grid_fin", C.c_double),
                ("rocket_radius", C.c_double), ("frontal_area", C.c_double),
                ("propellant_mass_stage1", C.c_double), ("c_gust_x", C.c_double),
                ("c_gust_y", C.c_double), ("inertia", C.c_double * 8),
                ("engine_height", C.c_double), ("cop", C.c_double),
                ("initial_state", C.c_double * 11), ("norm_vals", C.c_double * 7),
                ("v_opt_a", C.c_double), ("v_opt_b", C.c_double),
                ("n_gf_ca", C.c_int32), ("n_gf_cn", C.c_int32),
                ("gf_ca_mach", C.c_void_p), ("gf_ca_val", C.c_void_p),
                ("gf_cn_mach", C.c_void_p), ("gf_cn_val", C.c_void_p),
                ("n_wind", C.c_int32), ("_pad1", C.c_int32),
                ("wind_alt_km", C.c_void_p), ("wind_speed", C.c_void_p),
                ("vk_Adu", C.c_double * 4), ("vk_Bdu", C.c_double * 2),
                ("vk_Adv", C.c_double * 4), ("vk_Bdv", C.c_double * 2),
                ("cd", PdRbfTable), ("cl", PdRbfTable), ("other", C.POINTER(PdOtherPhases))]


class PdConfig(C.Structure):
    _fields_ = [("phase", C.c_int32), ("rtd", C.c_int32), ("precision", C.c_int32),
                ("enable_wind", C.c_int32), ("stochastic_wind", C.c_int32),
                ("auto_reset", C.c_int32), ("n_envs", C.c_int32), ("device", C.c_int32),
                ("seed", C.c_uint64), ("rl_reward_scale", C.c_double), ("discount_factor", C.c_double),
                ("raw_actions", C.c_int32), ("exact_aero", C.c_int32)]


class PdSharedActor(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("deterministic", C.c_int32), ("max_action", C.c_float),
                ("fp32_path", C.c_int32), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p),
                ("b2", C.c_void_p), ("wm", C.c_void_p), ("bm", C.c_void_p), ("ws", C.c_void_p),
                ("bs", C.c_void_p), ("seed", C.c_uint64)]


EXPORTS = ["pd_last_error", "pd_version", "pd_create", "pd_destroy", "pd_reset", "pd_step",
           "pd_get_state", "pd_set_state", "pd_set_wind_tape", "pd_rollout_pso", "pd_rollout_policy",
           "pd_collect_shared_actor", "pd_actor_forward", "pd_pso_update", "pd_pso_seed_mean", "pd_pso_select", "pd_pso_gather", "pd_pso_apply",
           "pd_set_rollout_stream", "pd_measure_fma_peak", "pd_check_status", "pd_launch_count",
           "pd_set_info_mode", "pd_set_rollout_handoff", "pd_set_rollout_stages", "pd_set_rollout_lanes", "pd_aero_patch_stats", "pd_release_aero_patches"]

_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m psso_sac_for_powered_descent_b200.build` "
            "(needs nvcc; there is no CPU fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, u8p = C.c_void_p, C.c_int, C.c_void_p
    lib.pd_last_error.restype = C.c_char_p
    lib.pd_version.restype = C.c_int
    lib.pd_launch_count.restype = C.c_uint64
    lib.pd_create.argtypes = [C.POINTER(PdConfig), C.POINTER(PdParams), C.POINTER(vp)]
    lib.pd_destroy.argtypes = [vp]
    lib.pd_reset.argtypes = [vp, u8p, vp]
    lib.pd_step.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.pd_get_state.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.pd_set_state.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.pd_set_wind_tape.argtypes = [vp, vp, i32, vp]
    lib.pd_rollout_pso.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.pd_rollout_policy.argtypes = [vp, i32, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.pd_collect_shared_actor.argtypes = [vp, C.POINTER(PdSharedActor), i32, vp, vp, vp, vp, vp, vp, vp]
    lib.pd_actor_forward.argtypes = [vp, C.POINTER(PdSharedActor), vp, i32, vp, vp, vp]
    lib.pd_check_status.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.pd_set_rollout_stream.argtypes = [vp, C.c_int64, C.c_uint32]
    lib.pd_set_info_mode.argtypes = [vp, i32]
    lib.pd_set_rollout_handoff.argtypes = [vp, i32]
    lib.pd_set_rollout_stages.argtypes = [vp, i32, i32]
    lib.pd_set_rollout_lanes.argtypes = [vp, i32, i32]
    lib.pd_aero_patch_stats.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    lib.pd_release_aero_patches.argtypes = []
    lib.pd_release_aero_patches.restype = C.c_int64
    lib.pd_measure_fma_peak.argtypes = [i32, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.pd_pso_seed_mean.argtypes = [vp, i32, i32, vp, vp]
    lib.pd_pso_select.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    lib.pd_pso_gather.argtypes = [vp, C.c_int64, i32, i32, vp, vp, i32, vp, vp]
    lib.pd_pso_apply.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.pd_pso_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, C.c_int64, C.c_double,
                                  C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, i32, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("pd_last_error", "pd_launch_count"):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError(load_library().pd_last_error().decode())


# ---------------------------------------------------------------------------------------
# host-side constant preparation
# ---------------------------------------------------------------------------------------
_ISA = ((-5.0e3, 320.65, -6.5e-3, 1.77687e5), (0.0e3, 288.15, -6.5e-3, 1.01325e5),
        (11.0e3, 216.65, 0.0, 2.26320e4), (20.0e3, 216.65, 1.0e-3, 5.47487e3),
        (32.0e3, 228.65, 2.8e-3, 8.68014e2), (47.0e3, 270.65, 0.0, 1.10906e2),
        (51.0e3, 270.65, -2.8e-3, 6.69384e1), (71.0e3, 214.65, -2.0e-3, 3.95639e0))


def _speed_of_sound(alt):
    """ISA speed of sound; only used to place the initial RBF neighbour-set hints."""
    alt = min(max(alt, 0.0), 81019.0)
    H = 6356766.0 * alt / (6356766.0 + alt)
    lay = [l for l in _ISA if H >= l[0]][-1]
    T = lay[1] + lay[2] * (H - lay[0])
    return math.sqrt(1.4 * 287.05287 * T)


def wind_profile(wind_table, percentile):
    """compile_horizontal_fixed_wind / interpolate_percentile
    (src/envs/wind/HorizontalWindSpeed.py:44-114): the reference passes a number that never
    matches the string keys, so two tabulated percentiles are always blended."""
    from scipy.interpolate import interp1d
    if percentile in wind_table:
        speed = np.array(wind_table[percentile]["wind_speed"])
        alt = np.array(wind_table[percentile]["altitude_km"])
    else:
        req = float(percentile)
        names = list(wind_table.keys())
        vals = [float(n.split("_")[0]) for n in names]
        idx = int(np.searchsorted(vals, req))
        if idx == 0:
            lo = hi = names[0]
            w = 1.0
        elif idx == len(vals):
            lo = hi = names[-1]
            w = 0.0
        else:
            lo, hi = names[idx - 1], names[idx]
            w = (req - vals[idx - 1]) / (vals[idx] - vals[idx - 1])
        la, ha = np.array(wind_table[lo]["altitude_km"]), np.array(wind_table[hi]["altitude_km"])
        alt = np.unique(np.concatenate([la, ha]))
        lf = interp1d(la, np.array(wind_table[lo]["wind_speed"]), kind="linear", bounds_error=False,
                      fill_value="extrapolate")
        hf = interp1d(ha, np.array(wind_table[hi]["wind_speed"]), kind="linear", bounds_error=False,
                      fill_value="extrapolate")
        speed = lf(alt) * (1 - w) + hf(alt) * w
    order = np.argsort(alt)
    return np.ascontiguousarray(alt[order], float), np.ascontiguousarray(speed[order], float)


def gust_filter(L, V=100.0, dt=0.1):
    """Unit-sigma ZOH discretisation of the 2-state gust filter (src/envs/wind/vonkarman.py:17-31);
    B_d is linear in sigma, the device multiplies by the per-env sigma."""
    from scipy.signal import cont2discrete
    w0 = V / L
    zeta = 1.0 / math.sqrt(2.0)
    scale = math.sqrt(math.pi / (2.0 * w0 ** 3))
    A = np.array([[0.0, 1.0], [-w0 ** 2, -2.0 * zeta * w0]])
    B = np.array([[0.0], [1.0 * scale]])
    Ad, Bd, _, _, _ = cont2discrete((A, B, np.array([[0.0, 1.0]]), np.zeros((1, 1))), dt)
    return Ad.reshape(-1).copy(), Bd.reshape(-1).copy()


_table_cache = {}


def aero_tables(p: RocketParams):
    key = id(p)
    if key not in _table_cache:
        cd = rbf_sets.build_table(p.cd_mach, p.cd_aoa, p.cd_val, rbf_sets.CD_BOXES,
                                  cache_dir=CACHE_DIR, grids=rbf_sets.CD_GRIDS)
        cl = rbf_sets.build_table(p.cl_mach, p.cl_aoa, p.cl_val, rbf_sets.CL_BOXES,
                                  cache_dir=CACHE_DIR, grids=rbf_sets.CL_GRIDS)
        _table_cache[key] = (cd, cl)
    return _table_cache[key]


def _rbf_struct(tbl, keep):
    t = PdRbfTable()
    L = len(tbl.levels)
    t.n_levels, t.n_points, t.n_sets, t.hash_size = L, len(tbl.mach_sorted), tbl.n_sets, len(tbl.hash_keys)
    for l in range(L):
        t.levels[l] = float(tbl.levels[l])
    for l in range(6):
        t.level_off[l] = int(tbl.level_off[min(l, L)])
    arrs = [np.ascontiguousarray(tbl.mach_sorted, np.float64),
            np.ascontiguousarray(tbl.points, np.float64).reshape(-1),
            np.ascontiguousarray(tbl.rows, np.uint8).reshape(-1),
            np.ascontiguousarray(tbl.hash_keys, np.uint64),
            np.ascontiguousarray(tbl.hash_vals, np.int32)]
    keep.extend(arrs)
    t.mach_sorted, t.points, t.rows, t.hash_keys, t.hash_vals = [a.ctypes.data for a in arrs]
    t.n_grids = len(tbl.grids)
    for i, g in enumerate(tbl.grids):
        cg = t.grids[i]
        cg.m0, cg.dm, cg.a0, cg.da, cg.nm, cg.na = g.m0, g.dm, g.a0, g.da, g.nm, g.na
        cells = np.ascontiguousarray(g.cells, np.int32)
        ih = np.ascontiguousarray(g.imp_hint, np.uint64)
        ii = np.ascontiguousarray(g.imp_id, np.int32)
        keep.extend([cells, ih, ii])
        cg.n_impure = len(ii)
        cg.cells, cg.imp_hint, cg.imp_id = cells.ctypes.data, ih.ctypes.data, ii.ctypes.data
    return t


def make_params(p: RocketParams, percentile=50):
    """RocketParams -> (PdParams, keepalive list of the numpy buffers it points into)."""
    keep = []
    c = PdParams()
    c.thrust_per_engine, c.nozzle_exit_pressure = p.thrust_per_engine, p.nozzle_exit_pressure
    c.nozzle_exit_area, c.v_exhaust = p.nozzle_exit_area, p.v_exhaust
    c.n_engines_gimballed = p.n_engines_gimballed
    c.grid_fin_area, c.d_base_grid_fin = p.grid_fin_area, p.d_base_grid_fin
    c.rocket_radius, c.frontal_area = p.rocket_radius, p.frontal_area
    c.propellant_mass_stage1 = p.propellant_mass_stage1
    c.c_gust_x, c.c_gust_y = p.c_gust_x, 0.0
    for i, k in enumerate(("I_dry", "h_f", "h_lower", "h_ox", "m_dry", "m_f", "m_ox", "x_dry")):
        c.inertia[i] = p.inertia[k]
    c.engine_height, c.cop = p.engine_height, p.cop
    for i in range(11):
        c.initial_state[i] = p.initial_state[i]
    for i in range(7):
        c.norm_vals[i] = p.norm_vals[i]
    c.v_opt_a, c.v_opt_b = p.v_opt_a, p.v_opt_b

    def sorted_xy(x, y):
        x, y = np.asarray(x, float), np.asarray(y, float)
        o = np.argsort(x, kind="mergesort")       # scipy interp1d sorts its table the same way
        a, b = np.ascontiguousarray(x[o]), np.ascontiguousarray(y[o])
        keep.extend([a, b])
        return a, b
    cam, cav = sorted_xy(p.gf_ca_mach, p.gf_ca_val)
    cnm, cnv = sorted_xy(p.gf_cn_mach, p.gf_cn_val)
    c.n_gf_ca, c.n_gf_cn = len(cam), len(cnm)
    c.gf_ca_mach, c.gf_ca_val = cam.ctypes.data, cav.ctypes.data
    c.gf_cn_mach, c.gf_cn_val = cnm.ctypes.data, cnv.ctypes.data
    alt, spd = wind_profile(p.wind_table, percentile)
    keep.extend([alt, spd])
    c.n_wind = len(alt)
    c.wind_alt_km, c.wind_speed = alt.ctypes.data, spd.ctypes.data
    Adu, Bdu = gust_filter(100.0)
    Adv, Bdv = gust_filter(30.0)
    for i in range(4):
        c.vk_Adu[i], c.vk_Adv[i] = Adu[i], Adv[i]
    for i in range(2):
        c.vk_Bdu[i], c.vk_Bdv[i] = Bdu[i], Bdv[i]
    cd, cl = aero_tables(p)
    c.cd = _rbf_struct(cd, keep)
    c.cl = _rbf_struct(cl, keep)
    o = getattr(p, "other_phases", None)
    if o:
        t = PdOtherPhases()
        t.n_engines_stage1 = int(o["n_engines_stage1"])
        t.max_rcs_force_per_thruster = o["max_rcs_force_per_thruster"]
        t.d_base_rcs_bottom, t.d_base_rcs_top = o["d_base_rcs_bottom"], o["d_base_rcs_top"]
        for i, k in enumerate(INERTIA_FULL_ORDER):
            t.inertia_full[i] = o["inertia_full"][k]
        t.engine_height_full = o["engine_height_full"]
        t.cop_full = o["cop_d0_full"] * o["cop_length_full"]
        for r, ph in enumerate(("subsonic", "supersonic", "ballistic_arc_descent", "flip_over_boostbackburn")):
            for i in range(11):
                t.initial_state[r][i] = o["initial_states"][ph][i]
            for i, v in enumerate(o["norm_vals"][ph]):
                t.norm_vals[r][i] = v
        rt = o["ref_traj_ascent"]
        arrs = [np.ascontiguousarray(rt[k], np.float64) for k in ("y", "x", "vx", "vy")]
        keep.extend(arrs)
        t.n_ref = len(arrs[0])
        t.ref_y, t.ref_x, t.ref_vx, t.ref_vy = [a.ctypes.data for a in arrs]
        for i in range(5):
            t.ref_terminal[i] = o["ref_traj_ascent_terminal"][i]
        keep.append(t)
        c.other = C.pointer(t)
    return c, keep
