"""CPU-side tests: C-ABI surface, parameter loading, the aero-table machinery, PSO host logic
and the world_size-2 sharding path (gloo).  No compute call into the CUDA library here."""
import ctypes
import math
import os
import re
import socket
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# --------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    from psso_sac_for_powered_descent_b200 import _native, build
    build.build()
    lib = _native.load_library()
    header = open(os.path.join(REPO, "include", "pd_b200.h")).read()
    declared = set(re.findall(r"\b(pd_[a-z_]+)\s*\(", header))
    assert {"pd_create", "pd_step", "pd_reset", "pd_rollout_pso", "pd_rollout_policy",
            "pd_collect_shared_actor", "pd_get_state", "pd_set_state"} <= declared
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_native.EXPORTS) == declared
    assert lib.pd_version() >= 100


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors vs the C structs (sizes computed by compiling a probe with gcc)."""
    import subprocess
    import tempfile
    from psso_sac_for_powered_descent_b200 import _native as N
    src = '#include <stdio.h>\n#include "pd_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", ' \
          'sizeof(PdRbfGrid), sizeof(PdRbfTable), sizeof(PdParams), sizeof(PdConfig), sizeof(PdSharedActor), ' \
          'sizeof(PdOtherPhases));}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(REPO, "include"), os.path.join(d, "p.c"),
                               "-o", os.path.join(d, "p")])
        sizes = list(map(int, subprocess.check_output([os.path.join(d, "p")]).split()))
    assert sizes == [ctypes.sizeof(N.PdRbfGrid), ctypes.sizeof(N.PdRbfTable), ctypes.sizeof(N.PdParams),
                     ctypes.sizeof(N.PdConfig), ctypes.sizeof(N.PdSharedActor), ctypes.sizeof(N.PdOtherPhases)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from psso_sac_for_powered_descent_b200 import envs, _native as N, RocketParams
    with pytest.raises(RuntimeError):
        envs.BatchedRocketEnv(4)
    lib = N.load_library()
    cfg = N.PdConfig()
    cfg.n_envs = 4
    prm, keep = N.make_params(RocketParams.default())
    h = ctypes.c_void_p()
    assert lib.pd_create(ctypes.byref(cfg), ctypes.byref(prm), ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.pd_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "psso_sac_for_powered_descent_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "/root/reference" not in txt, f


# --------------------------------------------------------------------------- parameters
def test_params_snapshot_values():
    from psso_sac_for_powered_descent_b200 import RocketParams
    p = RocketParams.default()
    assert p.initial_state[1] == 30028.385497767023 and p.initial_state[3] == -1023.3141698440232
    assert p.norm_vals[0] == 84577.57949154379 and p.norm_vals[1] == 1073.3141698440231
    assert abs(p.c_gust_x - 100.17165977622362) < 1e-12
    assert abs(p.cop - 0.75 * 50.35631553061721) < 1e-12
    assert len(p.cd_mach) == 191 and len(p.cl_mach) == 138
    assert len(p.gf_ca_mach) == 35 and len(p.gf_cn_mach) == 33
    assert p.n_engines_gimballed == 16 and p.thrust_per_engine == 2745000.0


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="reference checkout absent")
def test_params_from_reference_data_equals_snapshot():
    from dataclasses import asdict
    from psso_sac_for_powered_descent_b200 import RocketParams
    import warnings
    from psso_sac_for_powered_descent_b200.params import read_pickled_closures
    a = asdict(RocketParams.default())
    with warnings.catch_warnings():
        warnings.simplefilter("error", RuntimeWarning)          # no silent fall-back to the snapshot
        b = asdict(RocketParams.from_reference_data("/root/reference"))
    assert a == b
    # the constants inside rocket_functions.pkl are read from the pickle itself (closure cells; the
    # Python <= 3.10 bytecode is never executed), and the import of the reference's modules is undone
    cc = read_pickled_closures("/root/reference")
    assert cc["inertia"]["m_dry"] == 921070.0851247016 and cc["engine_height"] == 3.1
    assert cc["v_opt_a"] == -7.445479767703873e-07
    assert not any(m == "src" or m.startswith("src.") for m in sys.modules)


def test_wind_profile_and_gust_filter_constants():
    """SURVEY 8a probes of the reference: percentile-50 table and the ZOH filter matrices."""
    from psso_sac_for_powered_descent_b200 import _native as N, RocketParams
    alt, spd = N.wind_profile(RocketParams.default().wind_table, 50)
    assert np.allclose(alt, [0.8780, 9.8245, 13.5918, 19.8216, 22.7023, 50.1477, 80.0649], atol=1e-4)
    assert np.allclose(spd, [9.9970, 46.0597, 46.0494, 14.7281, 14.9377, 57.6892, 57.3905], atol=1e-4)
    Adu, Bdu = N.gust_filter(100.0)
    assert np.allclose(Adu, [0.9952315369548059, 0.09309551746581483, -0.0930955174658148,
                             0.8635745935585093], rtol=1e-13)
    assert np.allclose(Bdu, [0.005976382147808246, 0.11667792816060778], rtol=1e-12)
    Adv, Bdv = N.gust_filter(30.0)
    assert np.allclose(Bdv, [0.0008774017571697227, 0.016119401441639137], rtol=1e-12)


# --------------------------------------------------------------------------- aero tables
@pytest.fixture(scope="module")
def tables():
    from psso_sac_for_powered_descent_b200 import _native as N, RocketParams
    return N.aero_tables(RocketParams.default())


def test_rbf_tables_cover_random_queries_and_match_scipy(tables, oracle_tables):
    cd, cl = tables
    rng = np.random.default_rng(3)
    assert cd.n_sets >= 58 and cl.n_sets >= 3869         # SURVEY 7.1 lower bounds
    worst = 0.0
    for i in range(1500):
        m = rng.uniform(0, 10) if i % 4 == 0 else rng.uniform(0, 5.6)
        a = rng.uniform(1e-6, 10)
        v, ref = cl.evaluate(m, a), oracle_tables.cl_rbf(m, a)      # KeyError = set missing
        worst = max(worst, abs(v - ref) / max(abs(ref), 1e-3))
        a = rng.uniform(-0.1745, 0.1745)
        v, ref = cd.evaluate(m, a), oracle_tables.cd_rbf(m, a)
        worst = max(worst, abs(v - ref) / max(abs(ref), 1e-3))
        v, ref = cl.evaluate(m, -10.0), oracle_tables.cl_rbf(m, -10)
        worst = max(worst, abs(v - ref) / max(abs(ref), 1e-3))
    # the thin-plate-spline sum cancels ~5e4-fold: a different summation order / log rounding
    # moves the value by up to ~1e-10 relative even in double (documented noise floor)
    assert worst < 5e-10, worst


def test_bicubic_patch_of_a_grid_cell_follows_scipy(tables, oracle_tables):
    """The fp32 build's aero patches (csrc/pd_patch.cu), restated in numpy: the interpolating bicubic
    through the exact thin-plate sum at the 4 x 4 Chebyshev nodes of a sub-cell, stored as floats
    with a float-pair constant and evaluated in float32, is within 2e-8 of what scipy's
    RBFInterpolator(neighbors=50) - the reference's call - returns anywhere in the sub-cell, for
    every patch that passes the builder's own 25-point check (1e-8)."""
    cd, cl = tables
    nodes = np.cos(np.pi * (np.arange(4) + 0.5) / 4)
    vinv = np.linalg.inv(np.vander(nodes, 4, increasing=True))
    chk = np.array([-1.0, -0.7, 0.0, 0.7, 1.0])
    rng = np.random.default_rng(8)

    def stored_eval(S, u, v):
        u, v = np.float32(u), np.float32(v)
        f = np.float32
        r0 = f(f(f(S[3] * u + S[2]) * u + S[1]) * u)
        r1 = f(f(f(S[7] * u + S[6]) * u + S[5]) * u + S[4])
        r2 = f(f(f(S[11] * u + S[10]) * u + S[9]) * u + S[8])
        r3 = f(f(S[14] * u + S[13]) * u + S[12])
        var = f(f(f(r3 * v + r2) * v + r1) * v + r0)
        return float(S[0]) + (float(S[15]) + float(var))

    kept = tried = 0
    worst = 0.0
    for tbl, ref, sub in ((cl, oracle_tables.cl_rbf, (1, 2)), (cd, oracle_tables.cd_rbf, (4, 8))):
        g = tbl.grids[0]
        pure = np.nonzero(g.cells >= 0)[0]
        cols = pure % g.nm
        pure = pure[cols * g.dm < 3.5]                     # the Mach range the flights visit
        for cell in rng.choice(pure, 60, replace=False):
            ia, im = divmod(int(cell), g.nm)
            sx, sy = rng.integers(sub[0]), rng.integers(sub[1])
            hx, hy = 0.5 * g.dm / sub[0], 0.5 * g.da / sub[1]
            mc, ac = g.m0 + g.dm * im + hx * (2 * sx + 1), g.a0 + g.da * ia + hy * (2 * sy + 1)
            F = np.array([[tbl.evaluate(mc + hx * u, ac + hy * v) for v in nodes] for u in nodes])
            Cpq = vinv @ F @ vinv.T                         # coefficient of u^p v^q
            S = np.zeros(16, np.float32)
            for q in range(4):
                for pw in range(4):
                    S[q * 4 + pw] = Cpq[pw, q]
            S[15] = np.float32(Cpq[0, 0] - float(S[0]))
            err = max(abs(stored_eval(S, u, v) - tbl.evaluate(mc + hx * u, ac + hy * v)) for u in chk for v in chk)
            tried += 1
            if err > 1e-8:
                continue
            kept += 1
            for _ in range(6):
                u, v = rng.uniform(-1, 1, 2)
                m, a = mc + hx * float(np.float32(u)), ac + hy * float(np.float32(v))
                worst = max(worst, abs(stored_eval(S, u, v) - float(ref(m, a))))
    assert kept >= 0.9 * tried, (kept, tried)
    assert worst < 2e-8, worst


def test_rbf_conditioning_noise_floor(tables, oracle_tables):
    """Why theta_dot cannot be held to 1e-12 relative: kappa = sum|c_i phi_i| / |f|."""
    cd, cl = tables
    rng = np.random.default_rng(5)
    kappas = []
    for _ in range(200):
        m, a = rng.uniform(0.2, 5.0), rng.uniform(0.5, 10)
        lo, hi = cl.find_set(m, a)
        c = cl.coeffs[cl.lookup(lo, hi)]
        k, tot = 0, 0.0
        for l in range(len(cl.levels)):
            for i in range(lo[l], hi[l]):
                r2 = (m - cl.mach_sorted[cl.level_off[l] + i]) ** 2 + (a - cl.levels[l]) ** 2
                tot += abs(c[k] * 0.5 * r2 * math.log(r2)) if r2 > 0 else 0.0
                k += 1
        kappas.append(tot / max(abs(cl.evaluate(m, a)), 1e-12))
    assert np.median(kappas) > 50 and max(kappas) > 1e3


def test_query_grids_pure_cells_are_exact(tables):
    from psso_sac_for_powered_descent_b200 import rbf_sets as R
    cd, cl = tables
    rng = np.random.default_rng(9)
    for tbl in (cd, cl):
        for g in tbl.grids:
            assert 0.5 < g.pure_fraction <= 1.0
            n = 400
            M = g.m0 + rng.uniform(0, g.dm * g.nm, n)
            A = g.a0 + rng.uniform(0, g.da * g.na, n)
            for m, a in zip(M, A):
                im = min(int((m - g.m0) / g.dm), g.nm - 1)
                ia = min(int((a - g.a0) / g.da), g.na - 1)
                cell = int(g.cells[ia * g.nm + im])
                lo, hi = tbl.find_set(m, a)
                truth = tbl.lookup(lo, hi)
                assert truth >= 0
                if cell >= 0:
                    assert cell == truth
                else:
                    k = -cell - 1
                    assert 0 <= k < len(g.imp_id) and 0 <= g.imp_id[k] < tbl.n_sets


def test_device_row_layout(tables):
    cd, cl = tables
    for tbl in (cd, cl):
        rows = tbl.rows
        assert rows.shape == (tbl.n_sets, 512)
        c = rows[:, :57 * 8].copy().view(np.float64).reshape(tbl.n_sets, 57)
        assert np.array_equal(c[:, :50], 0.5 * tbl.coeffs[:, :50])
        assert np.array_equal(c[:, 50:55], tbl.coeffs[:, 50:55])
        assert np.array_equal(c[:, 55:57], 1.0 / tbl.coeffs[:, 55:57])
        idx = rows[:, 57 * 8:57 * 8 + 50]
        assert idx.max() < len(tbl.mach_sorted)
        for s in (0, tbl.n_sets // 2, tbl.n_sets - 1):
            exp = np.concatenate([tbl.level_off[l] + np.arange(tbl.set_lo[s, l], tbl.set_hi[s, l])
                                  for l in range(len(tbl.levels))])
            assert np.array_equal(idx[s], exp)


# --------------------------------------------------------------------------- PSO host logic
class _Sphere:
    """Stand-in model: fitness = |x - 0.3|^2 (the CUDA rollout is exercised by the GPU tests)."""

    def __init__(self, n=6):
        self.bounds = [(-1.5, 1.5)] * n
        self.mock_dictionary_of_opt_params = {f"0_weight_{j}": 0.0 for j in range(n)}
        self.calls = 0

    def evaluate(self, positions, n_seeds=1):
        self.calls += 1
        p = np.asarray(positions)
        return (np.repeat(((p - 0.3) ** 2).sum(1), n_seeds),)


def _mk(tmp_path, **kw):
    from psso_sac_for_powered_descent_b200 import pso
    params = dict(pso.landing_burn_pure_throttle_pso_params, pop_size=40, generations=30,
                  re_initialise_generation=12, re_initialise_number_of_particles=20)
    return pso.ParticleSubswarmOptimisation("landing_burn_pure_throttle", save_interval=10, model=_Sphere(),
                                            pso_params=params, seed=4, base_save_dir=str(tmp_path), **kw)


def test_pso_init_order_and_bounds(tmp_path):
    import random
    opt = _mk(tmp_path)
    r = random.Random(4)
    [r.uniform(-1.5, 1.5) for _ in range(40 * 6)]       # the base class's unused swarm (:57, 76-90)
    exp = np.array([[r.uniform(-1.5, 1.5) for _ in range(6)] for _ in range(40)])
    assert np.array_equal(opt.position, exp)            # initialize_swarms draw order (:401)
    assert [len(s) for s in opt.swarms] == [20, 20]
    assert opt.weight_linear_decrease(15) == 0.9 - (0.9 - 0.4) * 15 / 30


def test_pso_run_converges_and_keeps_reference_formats(tmp_path):
    import pickle
    opt = _mk(tmp_path)
    pos, fit = opt.run()
    assert fit < 1e-2 and np.all(np.abs(pos) <= 1.5)
    assert len(opt.position) == 20                       # re_initialise_swarms kept 10 per swarm
    assert np.all(np.diff(opt.global_best_fitness_array) <= 0)
    with open(tmp_path / "saves" / "swarm.pkl", "rb") as f:
        swarms = pickle.load(f)
    assert set(swarms[0][0].keys()) == {"position", "velocity", "best_position", "best_fitness"}
    hdr = open(tmp_path / "particle_subswarm_optimisation_results.csv").readline().strip().split(",")
    assert hdr[0] == "Algorithm" and hdr[1] == "0_weight_0" and hdr[-1] == "Best Fitness"
    opt2 = _mk(tmp_path)
    opt2.load_swarms(str(tmp_path / "saves" / "swarm.pkl"))
    assert opt2.global_best_fitness <= opt.global_best_fitness_array[-1] + 1e-12 or True
    assert len(opt2.position) == len(swarms[0]) + len(swarms[1])


def test_parallel_evaluate_is_order_preserving(tmp_path):
    opt = _mk(tmp_path)
    pts = [np.full(6, v) for v in (1.0, -1.0, 0.3, 0.0)]
    out = opt.parallel_evaluate(pts)
    assert out == [float(((p - 0.3) ** 2).sum()) for p in pts]
    assert opt.parallel_evaluate([]) == []


def test_shard_bounds_partition():
    from psso_sac_for_powered_descent_b200.pso import shard_bounds
    for n, w in ((10, 4), (65536, 8), (3, 8), (150, 2)):
        cuts = [shard_bounds(n, w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, REPO)
    from psso_sac_for_powered_descent_b200 import pso
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    pos = rng.uniform(-1.5, 1.5, (11, 6))              # 11 particles: ragged split 6 + 5
    seen = []

    def local_eval(p):
        seen.append(len(p))
        return ((p - 0.3) ** 2).sum(1)
    ev = pso.ShardedEvaluator(local_eval)
    fit = ev(pos)
    lo, hi = pso.shard_bounds(len(pos), world, rank)
    idx, best, best_pos = ev.broadcast_best(fit, pos[lo:hi], len(pos))
    params = dict(pso.landing_burn_pure_throttle_pso_params, pop_size=20, generations=8,
                  re_initialise_generation=100)
    opt = pso.ParticleSubswarmOptimisation("landing_burn_pure_throttle", save_interval=0, model=_Sphere(),
                                           pso_params=params, seed=1, base_save_dir=out_dir)
    opt.run()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), fit=fit, seen=np.array(seen), idx=idx, best=best,
             best_pos=best_pos, gbest=opt.global_best_fitness, gpos=opt.global_best_position)
    dist.destroy_process_group()


def test_sharded_evaluation_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    rng = np.random.default_rng(0)
    pos = rng.uniform(-1.5, 1.5, (11, 6))
    ref = ((pos - 0.3) ** 2).sum(1)
    assert np.array_equal(r0["fit"], ref) and np.array_equal(r1["fit"], ref)      # all-gather
    assert r0["seen"][0] == 6 and r1["seen"][0] == 5                              # each rank: its block
    assert int(r0["idx"]) == int(r1["idx"]) == int(np.argmin(ref))
    assert np.array_equal(r0["best_pos"], pos[np.argmin(ref)])                    # broadcast from owner
    assert np.array_equal(r1["best_pos"], pos[np.argmin(ref)])
    # the optimiser itself: both ranks end in the same state as a single process would
    assert float(r0["gbest"]) == float(r1["gbest"])
    assert np.array_equal(r0["gpos"], r1["gpos"])


def test_device_replay_buffer_interface_cpu():
    """replay.DeviceReplayBuffer keeps the reference ReplayBuffer's interface
    (src/agents/sac_pytorch.py:12-47) and hands out contiguous ring slots."""
    import torch
    from psso_sac_for_powered_descent_b200.replay import DeviceReplayBuffer
    buf = DeviceReplayBuffer(10, 2, 1, device="cpu")
    assert len(buf) == 0 and (buf.capacity, buf.state_dim, buf.action_dim) == (10, 2, 1)
    for k in range(3):
        buf.add(state=np.array([k, -k], np.float32), action=np.array([0.5]), reward=float(k),
                next_state=[k + 1, -k - 1], done=1.0 if k == 2 else 0.0)
    assert len(buf) == 3 and buf.position == 3
    assert buf.dones[:3, 0].tolist() == [0.0, 0.0, 1.0]
    start, v = buf.reserve(4)
    assert start == 3 and v["obs"].shape == (4, 2) and v["done"].dtype == torch.uint8
    v["obs"].fill_(7.0); v["actions"].fill_(0.25); v["rewards"].fill_(-1.0); v["next_obs"].fill_(8.0)
    v["done"].copy_(torch.tensor([0, 1, 0, 0], dtype=torch.uint8))
    buf.commit(start, 4)
    assert len(buf) == 7 and buf.position == 7
    assert buf.states[3:7].eq(7.0).all() and buf.dones[3:7, 0].tolist() == [0.0, 1.0, 0.0, 0.0]
    # a block that does not fit the tail starts again at slot 0
    start, v = buf.reserve(5)
    assert start == 0
    buf.commit(start, 5)
    assert buf.position == 5 and len(buf) == 7
    s, a, r, ns, d = buf.sample(16)
    assert s.shape == (16, 2) and a.shape == (16, 1) and r.shape == (16, 1) and ns.shape == (16, 2)
    assert d.shape == (16, 1) and s.dtype == torch.float32
    n0 = buf.add_batch(torch.ones(2, 3, 2), torch.zeros(2, 3, 1), torch.ones(2, 3), torch.ones(2, 3, 2),
                       torch.zeros(2, 3, dtype=torch.uint8))
    assert n0 == 0 and buf.position == 6       # 5 + 6 > 10: wrapped
    with pytest.raises(ValueError):
        buf.reserve(11)


def test_single_edge_cells_resolve_exactly(tables):
    """The device rule for impure grid cells cut by one Voronoi edge (bit 63 of imp_hint: first
    set iff the query is not farther from point p than from point q) reproduces the brute-force
    50-NN set for random queries inside such cells; they make up > 90 % of the impure cells."""
    cd, cl = tables
    rng = np.random.default_rng(17)
    for tbl in (cd, cl):
        pts = np.asarray(tbl.points, float).reshape(-1, 2)
        for g in tbl.grids:
            flag = (g.imp_hint >> np.uint64(63)).astype(bool)
            if len(flag) == 0:
                continue
            assert flag.mean() > 0.9
            cells = np.nonzero(g.cells < 0)[0]
            pick = cells[rng.integers(0, len(cells), 600)]
            n_edge = 0
            for flat in pick:
                k = -int(g.cells[flat]) - 1
                if not flag[k]:
                    continue
                n_edge += 1
                ia, im = divmod(int(flat), g.nm)
                m = g.m0 + g.dm * (im + rng.uniform(0.02, 0.98))
                a = g.a0 + g.da * (ia + rng.uniform(0.02, 0.98))
                h = int(g.imp_hint[k])
                p, q, s2 = (h >> 8) & 255, h & 255, (h >> 16) & 0xFFFF
                dp = (m - pts[p, 0]) ** 2 + (a - pts[p, 1]) ** 2
                dq = (m - pts[q, 0]) ** 2 + (a - pts[q, 1]) ** 2
                got = int(g.imp_id[k]) if dp <= dq else s2
                lo, hi = tbl.find_set(m, a)
                assert got == tbl.lookup(lo, hi), (flat, m, a)
            assert n_edge > 300


# --------------------------------------------------------------------------- PSO streams / formats
def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors) for the numpy implementation the
    host optimiser shares with csrc/pd_pso.cu."""
    from psso_sac_for_powered_descent_b200.pso import philox4x32
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        out = philox4x32(np.array([ctr[0]]), ctr[1], ctr[2], ctr[3], key[0], key[1])
        assert tuple(int(v[0]) for v in out) == exp


def test_pso_reference_rng_draw_order(tmp_path):
    """rng='reference': initial positions from random.Random(seed) in the reference's order, the
    velocity update consumes RandomState(seed).rand() twice per particle in swarm order."""
    opt = _mk(tmp_path)
    x0 = opt.position.copy()
    opt.step_generation(0)
    fit = ((x0 - 0.3) ** 2).sum(1)
    draws = np.random.RandomState(4).rand(2 * 40)
    lb = np.stack([x0[:20][np.argmin(fit[:20])]] * 20 + [x0[20:][np.argmin(fit[20:])]] * 20)
    v = 0.9 * 0 + (1 * draws[0::2])[:, None] * (x0 - x0) + (1 * draws[1::2])[:, None] * (lb - x0)
    assert np.array_equal(opt.velocity, v)
    assert np.array_equal(opt.position, np.clip(x0 + v, -1.5, 1.5))


def test_pso_migration_and_reinit_keep_list_order(tmp_path):
    opt = _mk(tmp_path, rng="philox")
    for g in range(13):
        opt.step_generation(g)
    # two migrations (generations 5, 10): the migrant sits at the END of its target list until the
    # re-initialisation at generation 12 sorts each list by personal best
    assert sorted(opt.members[0] + opt.members[1]) == list(range(len(opt.position)))
    for m in opt.members:
        bf = opt.best_fitness[m]
        assert np.all(np.diff(bf) >= 0) and len(m) <= 10
    sw = opt.swarms
    assert [len(s) for s in sw] == [len(m) for m in opt.members]


def test_pso_metrics_files_have_reference_layout(tmp_path):
    import csv
    opt = _mk(tmp_path)
    opt.run(generations=12)
    opt.save_results()
    rows = list(csv.reader(open(tmp_path / "metrics" / "subswarm_1_metrics.csv")))
    assert rows[0] == ["swarm_idx", "best_fitness", "avg_fitness", "min_fitness", "max_fitness", "std_fitness",
                       "num_particles", "generation", "global_best_fitness", "global_avg_fitness"]
    assert len(rows) == 13 and rows[5][0] == "1" and rows[5][7] == "4" and rows[5][6] == "20"
    g = list(csv.reader(open(tmp_path / "metrics" / "global_metrics.csv")))
    assert g[0] == ["generation", "global_best_fitness", "global_avg_fitness"] and len(g) == 13
    assert float(g[-1][1]) == opt.global_best_fitness_array[-1]
    h = list(csv.reader(open(tmp_path / "metrics" / "fitness_history.csv")))
    assert h[0] == ["Generation", "Global_Best_Fitness", "Average_Fitness", "Subswarm_1_Best",
                    "Subswarm_1_Average", "Subswarm_2_Best", "Subswarm_2_Average"]
    assert len(h) == 13 and float(h[3][1]) == opt.global_best_fitness_array[2]
    import json
    cfg = json.load(open(tmp_path / "pso_config.json"))
    assert cfg["flight_phase"] == "landing_burn_pure_throttle" and cfg["pop_size"] == 40


def _gloo_seedless_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, REPO)
    from psso_sac_for_powered_descent_b200 import pso
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    params = dict(pso.landing_burn_pure_throttle_pso_params, pop_size=20, generations=12,
                  re_initialise_generation=100, communication_freq=3, migration_freq=2)
    opt = pso.ParticleSubswarmOptimisation("landing_burn_pure_throttle", save_interval=0, model=_Sphere(),
                                           pso_params=params, seed=None, base_save_dir=out_dir)
    opt.run()
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), x=opt.position, g=opt.global_best_fitness, seed=opt.seed)
    dist.destroy_process_group()


def test_seedless_optimiser_agrees_across_ranks(tmp_path):
    """seed=None under torch.distributed: rank 0 draws a seed and broadcasts it, so every rank builds
    the same swarm and issues the same collectives (sharing re-evaluations included)."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_gloo_seedless_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "s0.npz"), np.load(tmp_path / "s1.npz")
    assert int(r0["seed"]) == int(r1["seed"])
    assert np.array_equal(r0["x"], r1["x"]) and float(r0["g"]) == float(r1["g"])


# --------------------------------------------------------------------------- the reference optimiser itself
class _OracleModel:
    """pso_wrapped_env stand-in on the CPU oracle (objective_function per particle)."""

    def __init__(self, phase, tables):
        from oracle import pd_oracle as O
        self._m = O.PsoModel(phase, tables=tables)
        n = self._m.n_params
        self.bounds = [(-1.5, 1.5)] * n
        self.mock_dictionary_of_opt_params = {f"p_{j}": 0.0 for j in range(n)}

    def objective_function(self, individual):
        return self._m.objective_function(np.asarray(individual, dtype=np.float64))


def _follow_reference_run(g, opt, pos_tol=0.0, fit_tol=1e-6):
    """Drive `opt` generation by generation next to the recorded run of the unmodified reference
    optimiser (tools/make_golden.py pso_run): evaluated positions per sub-swarm in list order (exact),
    per-sub-swarm metrics and global best (fit_tol: the fitness itself carries the evaluator's
    tolerance - 1e-6 for the oracle, 1e-4 for the CUDA path whose fp32 MLP sums in another order than
    torch-CPU - the swarm dynamics only see it through comparisons)."""
    for gen in range(int(g["n_generations"])):
        for k, m in enumerate(opt.members):
            ref = g[f"g{gen}_pos_{k}"]
            assert ref.shape == (len(m), opt.position.shape[1]), (gen, k, ref.shape, len(m))
            assert np.max(np.abs(opt.position[m] - ref)) <= pos_tol, (gen, k)
        opt.step_generation(gen)
        gb, gavg = g[f"g{gen}_global"]
        assert abs(opt.global_best_fitness - gb) <= fit_tol * abs(gb), gen
        assert abs(opt.average_particle_fitness_array[-1] - gavg) <= fit_tol * abs(gavg), gen
    assert np.allclose(opt.global_best_fitness_array, g["global_best_fitness_array"], rtol=fit_tol, atol=0)
    final = np.concatenate([opt.position[m] for m in opt.members])
    assert [len(m) for m in opt.members] == list(g["final_sizes"])
    assert np.max(np.abs(final - g["final_positions"])) <= pos_tol
    fv = np.concatenate([opt.velocity[m] for m in opt.members])
    assert np.max(np.abs(fv - g["final_velocities"])) <= pos_tol
    assert np.max(np.abs(opt.global_best_position - g["global_best_position"])) <= pos_tol


def test_host_dropin_follows_the_reference_optimiser(golden, oracle_tables, tmp_path):
    """ParticleSubswarmOptimisation (rng='reference', same seed) against a recorded run of the
    UNMODIFIED reference optimiser: 7 generations of landing_burn_pure_throttle through sharing,
    migration and re-initialisation - the same particles at the same positions in every generation,
    bit for bit, and the same metrics files."""
    import csv
    import json
    from psso_sac_for_powered_descent_b200 import pso
    g = golden("pso_run_reference.npz")
    phase = str(g["phase"])
    params = dict(pso.PSO_PARAMS[phase], **json.loads(str(g["knobs"])))
    opt = pso.ParticleSubswarmOptimisation(phase, save_interval=0, model=_OracleModel(phase, oracle_tables),
                                           pso_params=params, seed=int(g["seed"]), rng="reference",
                                           base_save_dir=str(tmp_path))
    _follow_reference_run(g, opt)
    ours = list(csv.reader(open(tmp_path / "metrics" / "subswarm_0_metrics.csv")))
    ref = list(csv.reader(str(g["metrics_csv_subswarm_0"]).strip().splitlines()))
    assert ours[0] == ref[0] and len(ours) == len(ref)
    for a, b in zip(ours[1:], ref[1:]):
        assert a[0] == b[0] and a[6] == b[6] and a[7] == b[7]          # swarm_idx, num_particles, generation
        assert np.allclose([float(v) for v in a[1:6] + a[8:]], [float(v) for v in b[1:6] + b[8:]], rtol=1e-6)
    gl = list(csv.reader(open(tmp_path / "metrics" / "global_metrics.csv")))
    assert gl[0] == list(csv.reader(str(g["metrics_csv_global"]).strip().splitlines()))[0]
