"""Regenerate psso_sac_for_powered_descent_b200/data/rocket_parameters_snapshot.json
from a reference checkout (container-only tool; uses tools/ref_harness.py to read
the closure cells of the two dill pickles)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tools.ref_harness import load_reference, REFERENCE_ROOT

def closure_cells():
    import dill
    load_reference()
    import src.RocketSizing.main_sizing  # noqa: F401  (classes needed to unpickle)
    real_load = getattr(dill.load, "__wrapped__", None)
    import dill._dill as _d
    with open("data/rocket_parameters/rocket_functions.pkl", "rb") as f:
        raw = _d.load(f)
    def cells(fn):
        return {n: c.cell_contents for n, c in zip(fn.__code__.co_freevars, fn.__closure__ or ())}
    inertia = {k: float(v) for k, v in cells(raw["x_cog_inertia_subrocket_2_lambda"]).items()}
    eh = float(cells(raw["d_cg_thrusters_subrocket_2_lambda"])["self"].engine_height)
    lengths = cells(raw["cop_subrocket_2_lambda"])["self"].lengths
    with open("data/reference_trajectory/landing_burn_controls/landing_initial_velocity_profile_guess.pkl", "rb") as f:
        vopt = _d.load(f)
    vc = cells(vopt)
    inertia_full = {k: float(v) for k, v in cells(raw["x_cog_inertia_subrocket_0_lambda"]).items()}
    eh0 = float(cells(raw["d_cg_thrusters_subrocket_0_lambda"])["self"].engine_height)
    return dict(inertia=inertia, engine_height=eh, cop_length=float(lengths[2]), cop_d0=0.75,
                v_opt_a=float(vc["a_opt"]), v_opt_b=float(vc["b_opt"]),
                inertia_full=inertia_full, engine_height_full=eh0, cop_length_full=float(lengths[0]),
                cop_d0_full=0.25)


def other_phases(cc):
    """Constants of the flight phases outside the two landing burns: taken from
    the reference's own loaders, executed unmodified."""
    import csv
    load_reference()
    from src.envs.load_initial_states import (load_subsonic_initial_state, load_supersonic_initial_state,
                                              load_high_altitude_ballistic_arc_initial_state,
                                              load_flip_over_initial_state)
    from src.envs.utils.input_normalisation import find_input_normalisation_vals
    from src.envs.utils.reference_trajectory_interpolation import reference_trajectory_lambda_func_y
    import pandas as pd
    sizing = {}
    with open("data/rocket_parameters/sizing_results.csv") as f:
        for row in csv.reader(f):
            sizing[row[0]] = row[2]
    data = pd.read_csv("data/reference_trajectory/ascent_controls/reference_trajectory_ascent_control.csv")
    _, term = reference_trajectory_lambda_func_y("subsonic")
    fl = lambda v: [float(x) for x in v]
    return dict(
        n_engines_stage1=int(sizing["Number of engines stage 1"]),
        max_rcs_force_per_thruster=float(sizing["max_RCS_force_per_thruster"]),
        d_base_rcs_bottom=float(sizing["d_base_rcs_bottom"]),
        d_base_rcs_top=float(sizing["d_base_rcs_top"]),
        inertia_full=cc["inertia_full"], engine_height_full=cc["engine_height_full"],
        cop_length_full=cc["cop_length_full"], cop_d0_full=cc["cop_d0_full"],
        initial_states=dict(subsonic=fl(load_subsonic_initial_state()),
                            supersonic=fl(load_supersonic_initial_state()),
                            ballistic_arc_descent=fl(load_high_altitude_ballistic_arc_initial_state()),
                            flip_over_boostbackburn=fl(load_flip_over_initial_state())),
        norm_vals={ph: fl(find_input_normalisation_vals(ph))
                   for ph in ("subsonic", "supersonic", "ballistic_arc_descent", "flip_over_boostbackburn")},
        ref_traj_ascent=dict(y=fl(data["y[m]"].values), x=fl(data["x[m]"].values),
                             vx=fl(data["vx[m/s]"].values), vy=fl(data["vy[m/s]"].values)),
        ref_traj_ascent_terminal=fl(term))

if __name__ == "__main__":
    from psso_sac_for_powered_descent_b200.params import RocketParams
    cc = closure_cells()
    p = RocketParams.from_reference_data(REFERENCE_ROOT, closure_cells=cc)
    p.other_phases = other_phases(cc)
    out = os.path.join(REPO, "psso_sac_for_powered_descent_b200/data/rocket_parameters_snapshot.json")
    p.to_json(out)
    print("wrote", out, os.path.getsize(out), "bytes")
    print("initial_state", p.initial_state)
    print("norm_vals", p.norm_vals)
    print("C_gust_x", p.c_gust_x, "cop", p.cop, cc)
