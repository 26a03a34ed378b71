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
    return dict(inertia=inertia, engine_height=eh, cop_length=float(lengths[2]), cop_d0=0.75,
                v_opt_a=float(vc["a_opt"]), v_opt_b=float(vc["b_opt"]))

if __name__ == "__main__":
    from psso_sac_for_powered_descent_b200.params import RocketParams
    cc = closure_cells()
    p = RocketParams.from_reference_data(REFERENCE_ROOT, closure_cells=cc)
    out = os.path.join(REPO, "psso_sac_for_powered_descent_b200/data/rocket_parameters_snapshot.json")
    p.to_json(out)
    print("wrote", out, os.path.getsize(out), "bytes")
    print("initial_state", p.initial_state)
    print("norm_vals", p.norm_vals)
    print("C_gust_x", p.c_gust_x, "cop", p.cop, cc)
