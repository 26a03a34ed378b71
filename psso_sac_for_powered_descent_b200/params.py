"""Vehicle / environment constants for the powered-descent hot path.

The reference scatters its constants over `data/rocket_parameters/sizing_results.csv`
(read in `src/envs/rockets_physics.py:715-719`), two dill pickles of closures
(`rockets_physics.py:721-722`), the V2 aero tables
(`src/envs/utils/aerodynamic_coefficients.py:8-55`), the grid-fin tables
(`src/envs/utils/grid_fin_aerodynamics.py:7-46`), the wind table
(`src/envs/wind/HorizontalWindSpeed.py:4-42`) and the recorded ballistic-arc
trajectory (`src/envs/load_initial_states.py:56-62`,
`src/envs/utils/input_normalisation.py:73-89`).  `RocketParams` gathers them
into one plain object, loaded either

* from a reference data directory (`RocketParams.from_reference_data(root)`),
  parsing the same files with the same parsers the reference uses (csv+float()
  where it uses csv, pandas where it uses pandas - the two differ in the last
  ulp), or
* from the committed snapshot `data/rocket_parameters_snapshot.json`
  (`RocketParams.default()`), which `tools/extract_params.py` generated from the
  reference checkout; floats are stored with `repr` so they round-trip exactly.

Nothing here computes physics; see csrc/ for the kernels.
"""
from __future__ import annotations

import csv
import json
import math
import os
from dataclasses import dataclass, field, asdict
from typing import Dict, List

import numpy as np

_SNAPSHOT = os.path.join(os.path.dirname(__file__), "data", "rocket_parameters_snapshot.json")

STATE_FIELDS = ("x", "y", "vx", "vy", "theta", "theta_dot", "gamma", "alpha", "mass",
                "mass_propellant", "time")

# Hard-coded knobs of compile_physics (rockets_physics.py:803-836, 909-934).
PHASES = ("landing_burn_pure_throttle", "landing_burn", "subsonic", "supersonic",
          "ballistic_arc_descent", "landing_burn_pure_throttle_Pcontrol")


@dataclass
class RocketParams:
    # sizing_results.csv
    thrust_per_engine: float = 0.0          # 'Thrust engine stage 1' [N]
    nozzle_exit_pressure: float = 0.0       # 'Nozzle exit pressure stage 1' [Pa]
    nozzle_exit_area: float = 0.0           # 'Nozzle exit area' [m^2]
    v_exhaust: float = 0.0                  # 'Exhaust velocity stage 1' [m/s]
    n_engines_gimballed: int = 0            # 'Number of engines gimballed stage 1'
    grid_fin_area: float = 0.0              # 'S_grid_fins'
    d_base_grid_fin: float = 0.0            # 'd_base_grid_fin'
    rocket_radius: float = 0.0              # 'Rocket Radius'
    frontal_area: float = 0.0               # 'Rocket frontal area'
    propellant_mass_stage1_ton: float = 0.0  # 'Actual propellant mass stage 1'
    stage1_mass_ton: float = 0.0            # 'Stage 1 Mass'
    # closure cells of rocket_functions.pkl (x_cog_inertia_subrocket_2_lambda etc.)
    inertia: Dict[str, float] = field(default_factory=dict)
    engine_height: float = 0.0
    cop_length: float = 0.0                 # lengths[2]
    cop_d0: float = 0.75
    # velocity-profile closure (landing_initial_velocity_profile_guess.pkl)
    v_opt_a: float = 0.0
    v_opt_b: float = 0.0
    # recorded ballistic-arc trajectory
    initial_state: List[float] = field(default_factory=list)
    norm_vals: List[float] = field(default_factory=list)   # landing_burn_input_normalisation()
    # aero tables in the reference's original (level-major, csv) order
    cd_mach: List[float] = field(default_factory=list)
    cd_aoa: List[float] = field(default_factory=list)
    cd_val: List[float] = field(default_factory=list)
    cl_mach: List[float] = field(default_factory=list)
    cl_aoa: List[float] = field(default_factory=list)
    cl_val: List[float] = field(default_factory=list)
    # grid fin tables, raw csv order
    gf_ca_mach: List[float] = field(default_factory=list)
    gf_ca_val: List[float] = field(default_factory=list)
    gf_cn_mach: List[float] = field(default_factory=list)
    gf_cn_val: List[float] = field(default_factory=list)
    # wind percentile table: name -> {"wind_speed": [...], "altitude_km": [...]}
    wind_table: Dict[str, Dict[str, List[float]]] = field(default_factory=dict)
    # flight phases outside the two landing burns (subsonic, supersonic, ballistic_arc_descent,
    # landing_burn_pure_throttle_Pcontrol): stage-1 engine count, RCS geometry, the full-rocket
    # inertia closure (x_cog_inertia_subrocket_0_lambda), per-phase initial states and
    # normalisation vectors, the ascent reference trajectory (tools/extract_params.py)
    other_phases: Dict = field(default_factory=dict)

    # ------------------------------------------------------------------ derived
    @property
    def propellant_mass_stage1(self) -> float:
        # rockets_physics.py:941  float(...)*1000
        return self.propellant_mass_stage1_ton * 1000

    @property
    def cop(self) -> float:
        # cop_func = d_0 * L  (src/RocketSizing/functions/cop_estimation.py:20-21)
        return self.cop_d0 * self.cop_length

    @property
    def c_gust_x(self) -> float:
        # src/envs/wind/size_gust_coefficients.py:11-22
        burnout_mass = (self.stage1_mass_ton - self.propellant_mass_stage1_ton) * 1000.0
        a_max_gust = 0.5 * 9.81
        rho_0 = 1.225
        v_wind_max = 10 + 6.0
        return 2 * burnout_mass * a_max_gust / (rho_0 * v_wind_max ** 2 * self.frontal_area)

    # ------------------------------------------------------------------ io
    def to_json(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump(asdict(self), f, indent=1)

    @classmethod
    def from_json(cls, path: str) -> "RocketParams":
        with open(path) as f:
            return cls(**json.load(f))

    @classmethod
    def default(cls) -> "RocketParams":
        return cls.from_json(_SNAPSHOT)

    @classmethod
    def from_reference_data(cls, root: str, closure_cells: dict | None = None) -> "RocketParams":
        """Parse the reference's own data files under `root` (its repo root).

        `closure_cells` supplies the constants held in the two dill pickles
        (they are Python<=3.10 bytecode and cannot be executed on 3.12; see
        tools/extract_params.py which reads their closure cells).  When omitted
        the committed snapshot's values are used for those fields only.
        """
        import pandas as pd  # the reference reads these files with pandas

        p = cls()
        sizing = {}
        with open(os.path.join(root, "data/rocket_parameters/sizing_results.csv")) as f:
            for row in csv.reader(f):
                sizing[row[0]] = row[2]
        p.thrust_per_engine = float(sizing["Thrust engine stage 1"])
        p.nozzle_exit_pressure = float(sizing["Nozzle exit pressure stage 1"])
        p.nozzle_exit_area = float(sizing["Nozzle exit area"])
        p.v_exhaust = float(sizing["Exhaust velocity stage 1"])
        p.n_engines_gimballed = int(sizing["Number of engines gimballed stage 1"])
        p.grid_fin_area = float(sizing["S_grid_fins"])
        p.d_base_grid_fin = float(sizing["d_base_grid_fin"])
        p.rocket_radius = float(sizing["Rocket Radius"])
        p.frontal_area = float(sizing["Rocket frontal area"])
        p.propellant_mass_stage1_ton = float(sizing["Actual propellant mass stage 1"])
        p.stage1_mass_ton = float(sizing["Stage 1 Mass"])

        if closure_cells is None:
            closure_cells = read_pickled_closures(root)
        if closure_cells is None:
            import warnings
            warnings.warn("rocket_functions.pkl could not be read here (needs `dill` and the reference's own "
                          "`src` package under the given root); taking the inertia / engine-height / CoP / "
                          "velocity-profile constants from the committed snapshot", RuntimeWarning)
            snap = cls.default()
            so = snap.other_phases
            closure_cells = dict(inertia=snap.inertia, engine_height=snap.engine_height,
                                 cop_length=snap.cop_length, cop_d0=snap.cop_d0,
                                 v_opt_a=snap.v_opt_a, v_opt_b=snap.v_opt_b,
                                 inertia_full=so["inertia_full"], engine_height_full=so["engine_height_full"],
                                 cop_length_full=so["cop_length_full"], cop_d0_full=so["cop_d0_full"])
        p.inertia = {k: float(v) for k, v in closure_cells["inertia"].items()}
        p.engine_height = float(closure_cells["engine_height"])
        p.cop_length = float(closure_cells["cop_length"])
        p.cop_d0 = float(closure_cells["cop_d0"])
        p.v_opt_a = float(closure_cells["v_opt_a"])
        p.v_opt_b = float(closure_cells["v_opt_b"])

        traj = pd.read_csv(os.path.join(
            root, "data/reference_trajectory/ballistic_arc_descent_controls/"
                  "state_action_ballistic_arc_descent_control.csv"))
        last = traj.iloc[-1]
        p.initial_state = [float(last[c]) for c in (
            "x[m]", "y[m]", "vx[m/s]", "vy[m/s]", "theta[rad]", "theta_dot[rad/s]",
            "gamma[rad]", "alpha[rad]", "mass[kg]", "mass_propellant[kg]", "time[s]")]
        cols = traj[["y[m]", "vy[m/s]", "theta[rad]", "theta_dot[rad/s]", "gamma[rad]",
                     "x[m]", "vx[m/s]"]].values
        p.norm_vals = [float(np.max(np.abs(cols[:, 0])) + 100),
                       float(np.max(np.abs(cols[:, 1])) + 50),
                       math.pi / 2, 0.01, math.pi * 3 / 2,
                       float(np.max(np.abs(cols[:, 5])) + 1500),
                       float(np.max(np.abs(cols[:, 6])) + 30)]

        for name, fn in (("cd", "V2_drag_coefficient.csv"), ("cl", "V2_lift_coefficient.csv")):
            mach, aoa, val = _parse_v2_table(os.path.join(
                root, "data/rocket_parameters/V2_aerodynamics", fn))
            setattr(p, f"{name}_mach", mach)
            setattr(p, f"{name}_aoa", aoa)
            setattr(p, f"{name}_val", val)

        ca = pd.read_csv(os.path.join(root, "data/rocket_parameters/GridFin/C_D_grid_fin.csv"),
                         header=None)
        p.gf_ca_mach = [float(v) for v in ca[0].values]
        p.gf_ca_val = [float(v) for v in ca[1].values]
        cn = pd.read_csv(os.path.join(root, "data/rocket_parameters/GridFin/C_N_alpha_grid_fin.csv"),
                         skiprows=2, header=None)
        p.gf_cn_mach = [float(v) for v in cn[0].values]
        p.gf_cn_val = [float(v) for v in cn[1].values]

        p.wind_table = _parse_wind_table(os.path.join(root, "data/Wind/horizontal_wind.csv"))
        p.other_phases = _other_phases(root, sizing, closure_cells, traj)
        return p


def read_pickled_closures(root):
    """The constants the reference keeps inside its two dill pickles
    (`data/rocket_parameters/rocket_functions.pkl`, loaded at rockets_physics.py:721-722, and
    `landing_initial_velocity_profile_guess.pkl`, landing_burn_pure_throttle.py:175-176), read from the
    pickles themselves: the closures were dilled under Python <= 3.10 and cannot be *executed* on a
    newer interpreter, but their closure cells (the numbers stage_inertia / full_rocket_inertia /
    d_cg_thrusters / cop_func close over) unpickle fine.  Needs `dill` and the reference's `src` package
    under `root` (the pickles reference its classes); plotting imports are stubbed for the duration.
    Returns None when that is not available."""
    import sys
    from unittest.mock import MagicMock
    try:
        import dill._dill as _d
    except Exception:
        return None
    stubs = ["matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.patches", "matplotlib.lines",
             "matplotlib.cm", "matplotlib.colors", "matplotlib.ticker", "matplotlib.animation",
             "matplotlib.collections", "mpl_toolkits", "mpl_toolkits.mplot3d", "pyswarm", "ambiance",
             "gymnasium", "lmdb"]

    def missing(m):
        import importlib.util
        try:
            return m not in sys.modules and importlib.util.find_spec(m.split(".")[0]) is None
        except (ImportError, ValueError):
            return True
    added = [m for m in stubs if missing(m)]            # only what this interpreter does not have
    cwd, had_path = os.getcwd(), root in sys.path
    before = set(sys.modules)
    try:
        for m in added:
            sys.modules[m] = MagicMock()
        if not had_path:
            sys.path.insert(0, root)
        os.chdir(root)                                  # the reference's data paths are relative
        import importlib
        importlib.import_module("src.RocketSizing.main_sizing")     # classes the pickle refers to

        def cells(fn):
            return {n: c.cell_contents for n, c in zip(fn.__code__.co_freevars, fn.__closure__ or ())}
        with open(os.path.join(root, "data/rocket_parameters/rocket_functions.pkl"), "rb") as f:
            raw = _d.load(f)
        with open(os.path.join(root, "data/reference_trajectory/landing_burn_controls/"
                                     "landing_initial_velocity_profile_guess.pkl"), "rb") as f:
            vc = cells(_d.load(f))
        lengths = cells(raw["cop_subrocket_2_lambda"])["self"].lengths
        return dict(
            inertia={k: float(v) for k, v in cells(raw["x_cog_inertia_subrocket_2_lambda"]).items()},
            engine_height=float(cells(raw["d_cg_thrusters_subrocket_2_lambda"])["self"].engine_height),
            cop_length=float(lengths[2]), cop_d0=0.75,         # main_sizing.py:215-217
            v_opt_a=float(vc["a_opt"]), v_opt_b=float(vc["b_opt"]),
            inertia_full={k: float(v) for k, v in cells(raw["x_cog_inertia_subrocket_0_lambda"]).items()},
            engine_height_full=float(cells(raw["d_cg_thrusters_subrocket_0_lambda"])["self"].engine_height),
            cop_length_full=float(lengths[0]), cop_d0_full=0.25)
    except Exception:
        return None
    finally:
        os.chdir(cwd)
        if not had_path and root in sys.path:
            sys.path.remove(root)
        for m in added:
            sys.modules.pop(m, None)
        for m in set(sys.modules) - before:             # the reference's `src.*` modules imported for unpickling
            if m == "src" or m.startswith("src."):
                sys.modules.pop(m, None)


_STATE_COLS = ("x[m]", "y[m]", "vx[m/s]", "vy[m/s]", "theta[rad]", "theta_dot[rad/s]", "gamma[rad]",
               "alpha[rad]", "mass[kg]", "mass_propellant[kg]", "time[s]")


def _other_phases(root, sizing, cc, ballistic_traj):
    """Constants of subsonic / supersonic / ballistic_arc_descent, parsed the way the reference
    parses them: load_initial_states.py:5-54, input_normalisation.py:5-71,
    reference_trajectory_interpolation.py:5-35, rockets_physics.py:727-801."""
    import pandas as pd
    fl = lambda v: [float(x) for x in v]
    asc = os.path.join(root, "data/reference_trajectory/ascent_controls")
    sub = pd.read_csv(os.path.join(asc, "subsonic_state_action_ascent_control.csv"))
    sup = pd.read_csv(os.path.join(asc, "supersonic_state_action_ascent_control.csv"))
    flip = pd.read_csv(os.path.join(root, "data/reference_trajectory/flip_over_and_boostbackburn_controls/"
                                          "state_action_flip_over_and_boostbackburn_control.csv"))
    ref = pd.read_csv(os.path.join(asc, "reference_trajectory_ascent_control.csv"))
    init_sub = np.array([0, 1.5, 0, 0, np.pi / 2, 0, 0, 0,
                         float(sizing["Initial mass (subrocket 0)"]) * 1000,
                         float(sizing["Actual propellant mass stage 1"]) * 1000, 0])

    def ascent_norm(d, add):
        st = d[["x[m]", "y[m]", "vx[m/s]", "vy[m/s]", "theta[rad]", "theta_dot[rad/s]", "alpha[rad]",
                "mass[kg]"]].values
        m = np.max(np.abs(st), axis=0)
        return [m[0] + add[0], m[1] + add[1], m[2] + add[2], m[3] + add[3], m[4] + math.radians(add[4]),
                m[5] * 2.5, m[6] + math.radians(3), m[7]]
    bs = ballistic_traj[["theta[rad]", "theta_dot[rad/s]", "alpha[rad]", "gamma[rad]"]].values
    bm = np.max(np.abs(bs), axis=0)
    states = ref[["x[m]", "y[m]", "vx[m/s]", "vy[m/s]", "mass[kg]"]].values
    # load_flip_over_initial_state (load_initial_states.py:33-45): last supersonic row, mass rebuilt
    # from the propellant left and the stage-1 structural mass
    last = sup.iloc[-1]
    init_flip = [float(last[c]) for c in _STATE_COLS]
    init_flip[8] = float(last["mass_propellant[kg]"]) + float(sizing["Actual structural mass stage 1"]) * 1000
    fs = flip[["theta[rad]", "theta_dot[rad/s]"]].values
    fm = np.max(np.abs(fs), axis=0)
    return dict(
        n_engines_stage1=int(sizing["Number of engines stage 1"]),
        max_rcs_force_per_thruster=float(sizing["max_RCS_force_per_thruster"]),
        d_base_rcs_bottom=float(sizing["d_base_rcs_bottom"]),
        d_base_rcs_top=float(sizing["d_base_rcs_top"]),
        inertia_full={k: float(v) for k, v in cc["inertia_full"].items()},
        engine_height_full=float(cc["engine_height_full"]),
        cop_length_full=float(cc["cop_length_full"]), cop_d0_full=float(cc["cop_d0_full"]),
        initial_states=dict(subsonic=fl(init_sub), supersonic=fl(sub.iloc[-1][list(_STATE_COLS)]),
                            ballistic_arc_descent=fl(flip.iloc[-1][list(_STATE_COLS)]),
                            flip_over_boostbackburn=fl(init_flip)),
        norm_vals=dict(subsonic=fl(ascent_norm(sub, (100, 500, 5, 50, 2))),
                       supersonic=fl(ascent_norm(sup, (2500, 5000, 100, 150, 5))),
                       ballistic_arc_descent=fl([bm[0] + math.radians(5), bm[1] * 2.5,
                                                 bm[3] + math.radians(5), bm[2] + math.radians(5)]),
                       flip_over_boostbackburn=fl([fm[0] + math.radians(5), fm[1] * 2.5])),
        ref_traj_ascent=dict(y=fl(ref["y[m]"].values), x=fl(ref["x[m]"].values),
                             vx=fl(ref["vx[m/s]"].values), vy=fl(ref["vy[m/s]"].values)),
        ref_traj_ascent_terminal=fl(states[-1]))


def _parse_v2_table(path):
    """Same acceptance rules as aerodynamic_coefficients.py:8-49 (header 'N_deg,,',
    a unit row that is skipped, ragged columns)."""
    with open(path) as f:
        lines = f.readlines()
    header = lines[0].strip().split(",")
    levels = [float(h.split("_")[0]) for h in header if "deg" in h]
    per_level = {a: ([], []) for a in levels}
    for line in lines[2:]:
        if not line.strip():
            continue
        vals = line.strip().split(",")
        if len(vals) < len(header):
            continue
        for i, a in enumerate(levels):
            mi, ci = 2 * i, 2 * i + 1
            if ci < len(vals) and vals[mi].strip() and vals[ci].strip():
                try:
                    m, c = float(vals[mi]), float(vals[ci])
                except ValueError:
                    continue
                per_level[a][0].append(m)
                per_level[a][1].append(c)
    mach, aoa, val = [], [], []
    for a in levels:
        mach += per_level[a][0]
        val += per_level[a][1]
        aoa += [a] * len(per_level[a][0])
    return mach, aoa, val


def _parse_wind_table(path):
    """HorizontalWindSpeed.py:4-42."""
    with open(path) as f:
        lines = f.readlines()
    names = [h for h in lines[0].strip().split(",") if h and not h.isspace()]
    out = {n: {"wind_speed": [], "altitude_km": []} for n in names}
    for line in lines[2:]:
        if not line.strip():
            continue
        vals = line.strip().split(",")
        if len(vals) < 2 * len(names):
            continue
        for i, n in enumerate(names):
            try:
                w, a = float(vals[2 * i]), float(vals[2 * i + 1])
            except (ValueError, IndexError):
                continue
            out[n]["wind_speed"].append(w)
            out[n]["altitude_km"].append(a)
    return out
