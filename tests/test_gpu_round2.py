"""GPU parity tests of the round-2 surface, through the C ABI: the CTA-resident layout, the device reference-trajectory
pipeline (against the scripts' own nested loops restated in tests/reference_loops.py and against the reference's
committed outputs), the LTV / x0-from-prediction closed loops run by ONE mpcv_closed_loop_ex call (against the golden
dados.csv / dados2.csv and the oracle), per-problem bounds, acceptable-level termination."""
import math
import os

import numpy as np
import pytest

from mpc_verde_b200 import problems
from mpc_verde_b200 import spec as S
from oracle import mpc_oracle as O
from tests import common, reference_loops

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1
OPTS = {"ipopt": {"max_iter": 2000, "print_level": 0, "acceptable_tol": 1e-8, "acceptable_obj_change_tol": 1e-6},
        "print_time": 0}


@pytest.fixture(scope="module")
def mv():
    import mpc_verde_b200 as m
    return m


def _solver(mv, prob, **kw):
    return mv.nlpsol("solver", "ipopt", prob, dict(OPTS, **kw))


# ---- CTA-resident layout ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["c2", "zeros", "tracker20", "pendulum"])
def test_resident_layout_vs_pipeline_and_oracle(mv, case):
    """MPCV_LAYOUT_RESIDENT runs the same phase functions on shared-memory rows with continuous batching: it must
    reproduce the pipeline's iteration path (statuses and iteration counts identical) and the oracle's optimum."""
    import torch
    if case in ("c2", "zeros"):
        prob = problems.unicycle_multiple_shooting()
        sp = prob["spec"]
        x0s, p = common.unicycle_batch(1024 if case == "c2" else 256, seed=20261 if case == "c2" else 77)
        lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0 if case == "c2" else None)
        w0 = problems.cold_start(sp, x0s) if case == "c2" else None      # zeros: slow path, failures
    elif case == "tracker20":
        prob = problems.unicycle_tracking(N=20, T=0.05, M=1)
        sp = prob["spec"]
        rng = np.random.default_rng(5)
        B = 512
        tt = np.arange(sp.N)[None, :] * sp.T + rng.uniform(0, 50, (B, 1))
        stage = np.stack([np.cos(0.1 * tt), np.sin(0.1 * tt), math.pi / 2 + 0.1 * tt, 0.1 + 0 * tt, 0.1 + 0 * tt], 2)
        x0 = stage[:, 0, :3] + rng.normal(size=(B, 3)) * 0.1
        p = np.concatenate([x0, stage.reshape(B, -1)], 1)
        lbx, ubx = problems.control_box(sp, (-1, -math.pi / 4), (1, math.pi / 4), (-20, -2, -np.inf), (20, 2, np.inf))
        w0 = problems.cold_start(sp, x0)
    else:
        sp, lbx, ubx, pglob, _, _ = common.pendulum_setup(N=10, ntu=3)
        prob = {"spec": sp}
        x0, p = common.pendulum_batch(sp, pglob, 256)
        w0 = problems.cold_start(sp, x0)
    res = {}
    for lay in (S.LAYOUT_RESIDENT, S.LAYOUT_PHASED):
        solver = _solver(mv, prob, layout=lay)
        sol = solver(x0=None if w0 is None else torch.as_tensor(w0).cuda(), lbx=lbx, ubx=ubx, p=torch.as_tensor(p).cuda())
        res[lay] = ({k: v.cpu().numpy() for k, v in sol.items()}, solver.stats())
    (a, sa), (b, sb) = res[S.LAYOUT_RESIDENT], res[S.LAYOUT_PHASED]
    assert np.array_equal(sa["status_code"], sb["status_code"])
    assert np.array_equal(sa["iter_count"], sb["iter_count"])
    ok = sa["status_code"] == 0
    assert ok.mean() > 0.95
    assert np.abs(a["x"][ok] - b["x"][ok]).max() <= 1e-11
    n = min(128, p.shape[0])
    ref = O.solve(sp, None if w0 is None else w0[:n], lbx, ubx, p[:n], nthreads=NCPU)
    assert np.array_equal(ref["status"], sa["status_code"][:n])
    same = (ref["status"] == 0) & (ref["iters"] == sa["iter_count"][:n])
    assert same.mean() >= 0.98
    assert np.abs(a["x"][:n][same] - ref["x"][same]).max() <= 1e-9
    assert np.abs(a["f"][:n][same] - ref["f"][same]).max() <= 1e-6 * np.abs(ref["f"]).max()


# ---- device reference-trajectory pipeline -----------------------------------------------------------------------
def test_device_reference_pipeline_vs_script_loops_and_goldens(mv):
    from mpc_verde_b200 import reference as R
    g = common.golden("lane_change.csv")
    a, b, c = g[:, 0], g[:, 1], g[:, 2]
    # lateral-error tables: the scripts' nested loops, and the columns dados2.csv stores (par[:, 0, t])
    for Nt in (5, 20):
        par = reference_loops.lateral_error_par(a, b, Nt, 0.05)                  # [4, Nt, Nsim]
        win = R.lateral_windows(a, b, Nt, 0.05).cpu().numpy()[0]                 # [Nsim, Nt, 4]
        assert np.abs(win - par.transpose(2, 1, 0)).max() <= 1e-11
    g2 = common.golden("lateral_lti_dados2.csv")
    win = R.lateral_windows(a, b, 5, 0.05).cpu().numpy()[0]
    assert np.abs(win[:, 0, :] - g2[:, 6:10]).max() <= 1e-12
    # scenario scaling = the loops run on the scaled path
    sc = np.array([[1.0, 1.0], [1.2, 0.7]])
    win2 = R.lateral_windows(a, b, 5, 0.05, scale=sc).cpu().numpy()
    par2 = reference_loops.lateral_error_par(a * 1.2, b * 0.7, 5, 0.05)
    assert np.abs(win2[0] - win).max() == 0.0 and np.abs(win2[1] - par2.transpose(2, 1, 0)).max() <= 1e-10
    # lane_change.py's extended path = its committed output out.csv
    o = common.golden("lane_change_out.csv")
    xt, yt, c2 = (q.cpu().numpy() for q in R.lane_change_extended(a, b, c))
    assert xt.size == o.shape[0] == 2210
    assert np.abs(xt - o[:, 0]).max() <= 1e-13 and np.abs(yt - o[:, 1]).max() <= 1e-13 and np.array_equal(c2, o[:, 2])
    # circle of Trajectory_tracking.py:84-97 (sliding window: entry t + k)
    par = reference_loops.circle_reference_par(10, 20, 0.2)
    cir = R.circle_reference(30, 0.2).cpu().numpy()
    for t in (0, 7, 19):
        assert np.abs(cir[t:t + 10] - par[:, :, t].T).max() <= 1e-14
    # Frenet bicycle tables of test2.py:79-100 (swapped p[2] / p[3] kept)
    n_steps = 480
    fw = R.frenet_windows(a, b, c, 20, 0.05, n_steps).cpu().numpy()[0]
    for t in (0, 1, 10, 250, 479):
        p = reference_loops.frenet_reference_par(a, b, c, 20, 0.05, t)
        assert np.abs(fw[t] - p).max() <= 1e-9 * (1 + np.abs(p).max()), t
    assert np.allclose(fw[10, :, 2], c[10:30])
    # unicycle references from a path: numpy's central differences
    T, dt = 200, 0.05
    ur = R.unicycle_path_reference(a[:T], b[:T], dt, scale=np.array([[1.1, 0.9]])).cpu().numpy()[0]
    xs, ys = a[:T] * 1.1, b[:T] * 0.9
    th = np.arctan2(np.gradient(ys), np.gradient(xs))
    v = np.hypot(np.gradient(xs), np.gradient(ys)) / dt
    w = np.gradient(th) / dt
    assert np.abs(ur[:, 0] - xs).max() == 0 and np.abs(ur[:, 2] - th).max() <= 1e-13
    assert np.abs(ur[:, 3] - np.clip(v, -1, 1)).max() <= 1e-12
    assert np.abs(ur[:, 4] - np.clip(w, -math.pi / 4, math.pi / 4)).max() <= 1e-9
    # per-step exact ZOH of the LTV models = mpc.util.c2d per step
    pgt = R.ltv_lateral(c, 0.05, 50, speed_scale=np.array([1.0, 1.3])).cpu().numpy()
    for bb, spd in ((0, 1.0), (1, 1.3)):
        for t in (0, 17, 49):
            A, Bd = problems.c2d(*problems.lateral_error_matrices(c[t] * spd), 0.05)
            assert np.abs(pgt[bb, t] - np.concatenate([A.ravel(), Bd.ravel()])).max() <= 1e-13
    vprof = np.linspace(0.4, 0.8, 12)
    pgd = R.ltv_dynamic_bicycle(vprof, 0.05, 12).cpu().numpy()[0]
    for t in (0, 5, 11):
        A, Bd = problems.c2d(*problems.dynamic_bicycle_matrices(vprof[t]), 0.05)
        ref = np.concatenate([A.ravel(), Bd.ravel()])
        assert np.abs(pgd[t] - ref).max() <= 1e-11 * (1 + np.abs(ref).max())


# ---- LTV closed loops in ONE call ---------------------------------------------------------------------------------
@pytest.mark.parametrize("ltv", [True, False])
def test_lateral_closed_loop_one_call_vs_dados(mv, ltv):
    """Trjectory_tracking_le_LTV.py:126-143 re-discretises Ac(c[t]) and rebuilds the solver every step; here the whole
    500-step loop — reference tables, per-step ZOH, solves, plant — is device work behind ONE closed-loop call,
    checked against the reference's own dumps dados.csv (LTV) / dados2.csv (LTI)."""
    from mpc_verde_b200 import reference as R
    g = common.golden("lane_change.csv")
    a, b, c = g[:, 0], g[:, 1], g[:, 2]
    Nt, Delta, Nsim = 5, 0.05, 500
    prob = problems.linear_tracking(3, Nt, Q=(10.0, 1.0, 0.0), R=0.01, T=Delta, ntu=1)     # Ntu = 1, no Du cost
    solver = _solver(mv, prob)
    sp = solver.spec
    assert sp.model == S.MODEL_LINEAR3_DU and sp.ntu == 1
    lbx, ubx = problems.control_box(sp, -0.3491, 0.3491)
    win = R.lateral_windows(a, b, Nt, Delta)
    kw = {}
    if ltv:
        kw["pglob_traj"] = R.ltv_lateral(c, Delta, Nsim)
    else:
        A, Bd = problems.c2d(*problems.lateral_error_matrices(c.mean()), Delta)
        kw["pglob"] = np.concatenate([A.ravel(), Bd.ravel()])
    r = solver.closed_loop(np.zeros((1, 4)), ptraj=win, windows=True, lbx=lbx, ubx=ubx, n_steps=Nsim,
                           warm_mode=S.WARM_COLD, horizons=True, step_times=True, **kw)
    assert r["status"][0] == 0 and r["steps"][0] == Nsim
    gold = common.golden("lateral_ltv_dados.csv" if ltv else "lateral_lti_dados2.csv")
    assert np.abs(r["controls"][0, :, 0] - gold[:, 3]).max() <= 1e-5
    assert np.abs(r["states"][0, 1:, :3] - gold[:, 0:3]).max() <= 1e-4
    # the same loop driven from Python, one oracle solve per step (tests/common.py)
    solve = lambda s, w0, lb, ub, p: O.solve(s, w0, lb, ub, p)["x"]
    u, x, _ = common.lateral_error_closed_loop(solve, ltv=ltv, nsim=60)
    assert np.abs(r["controls"][0, :60, 0] - u).max() <= 1e-8
    assert np.abs(r["states"][0, :61, :3] - x).max() <= 1e-8
    # histories: the predicted horizon of step t starts at the state of step t; step timers ran
    assert np.abs(r["horizons"][0, :, 0, :3] - r["states"][0, :-1, :3]).max() <= 1e-12
    assert r["step_ms"].shape == (Nsim,) and np.all(r["step_ms"] > 0)


def test_dynamic_bicycle_ltv_warm_started_loop_vs_oracle(mv):
    """BASELINE config 5 in small: dynamic bicycle, N = 50, LTV in v_ref[t] (c2d per scenario and step on the device),
    warm-started closed loop in one call, against a Python loop of oracle solves with the same shifted guess."""
    from mpc_verde_b200 import reference as R
    N, dt, nst, B = 50, 0.05, 4, 6
    prob = problems.linear_tracking(4, N, Q=(1, 1, 1, 1), R=1.0, T=dt)
    for layout in (S.LAYOUT_AUTO, S.LAYOUT_WARP):
        solver = _solver(mv, prob, layout=layout)
        sp = solver.spec
        rng = np.random.default_rng(11)
        lbx, ubx = problems.control_box(sp, -20.0, 20.0)
        x0 = rng.normal(size=(B, 4)) * np.array([0.2, 0.05, 0.1, 0.05])
        vprof = rng.uniform(0.4, 0.8, (B, nst))
        pgt = R.ltv_dynamic_bicycle(vprof, dt, nst)
        ptraj = np.zeros((B, nst + N, 5))
        ptraj[:, :, 0] = np.linspace(0, 1, nst + N)[None, :] * rng.uniform(0.5, 1.5, (B, 1))
        r = solver.closed_loop(x0, ptraj=ptraj, pglob_traj=pgt, lbx=lbx, ubx=ubx, n_steps=nst, warm_mode=S.WARM_SHIFT)
        assert np.all(r["status"] == 0)
        pg = pgt.cpu().numpy()
        nz = sp.nx + sp.nu
        xs, guess = x0.copy(), problems.cold_start(sp, x0)
        for t in range(nst):
            p = np.concatenate([xs, pg[:, t], ptraj[:, t:t + N].reshape(B, -1)], 1)
            ref = O.solve(sp, guess, lbx, ubx, p, nthreads=NCPU)
            assert np.all(ref["status"] == 0)
            u0 = ref["x"][:, sp.nx]
            assert np.abs(r["controls"][:, t, 0] - u0).max() <= 1e-8
            A, Bd = pg[:, t, :16].reshape(B, 4, 4), pg[:, t, 16:]
            xs = np.einsum("bij,bj->bi", A, xs) + Bd * u0[:, None]
            assert np.abs(r["states"][:, t + 1] - xs).max() <= 1e-9
            w = ref["x"]
            guess = w.copy()
            for k in range(N):
                guess[:, k * nz:k * nz + sp.nx] = w[:, (k + 1) * nz:(k + 1) * nz + sp.nx]
                ks = min(k + 1, N - 1)
                guess[:, k * nz + sp.nx:(k + 1) * nz] = w[:, ks * nz + sp.nx:(ks + 1) * nz]


def test_x0_from_prediction_mode_and_horizons(mv):
    """Trajectory_tracking.py:101-118: `saveguess()` + `fixvar("x",0,var["x",1])` — the next solve starts from the
    predicted x_1 while the plant is simulated on.  Plant and prediction use the same RK4 here, so the mode must
    reproduce the plant-state loop, and the recorded horizons must chain: x_1 of step t = x_0 of step t+1."""
    from mpc_verde_b200 import reference as R
    Nt, Delta, nst = 10, 0.2, 25
    solver = _solver(mv, problems.unicycle_tracking(N=Nt, T=Delta, M=1))
    sp = solver.spec
    lbx, ubx = problems.control_box(sp, (-1, -math.pi / 4), (1, math.pi / 4), (-20, -2, -np.inf), (20, 2, np.inf))
    ptraj = R.circle_reference(nst + Nt, Delta)
    x0 = np.array([[0.8, 0.1, 1.4], [1.1, -0.1, 1.7]])
    a = solver.closed_loop(x0, ptraj=ptraj, lbx=lbx, ubx=ubx, n_steps=nst, horizons=True)
    b = solver.closed_loop(x0, ptraj=ptraj, lbx=lbx, ubx=ubx, n_steps=nst, horizons=True, x0_from_prediction=True)
    assert np.all(a["status"] == 0) and np.all(b["status"] == 0)
    assert np.abs(a["controls"] - b["controls"]).max() <= 1e-9
    assert np.abs(b["horizons"][:, :-1, 1, :] - b["horizons"][:, 1:, 0, :]).max() == 0.0
    assert np.abs(a["horizons"][:, :, 1, :] - a["states"][:, 1:, :]).max() <= 1e-12
    ro = O.closed_loop(sp, x0, None, ptraj.cpu().numpy()[None].repeat(2, 0), lbx, ubx, nst, S.WARM_SHIFT, 0.0)
    assert np.abs(a["controls"] - ro["controls"]).max() <= 1e-7


def test_per_problem_bounds_equal_separate_calls(mv):
    import torch
    solver = _solver(mv, problems.unicycle_multiple_shooting())
    sp = solver.spec
    x0s, p = common.unicycle_batch(96)
    w0 = problems.cold_start(sp, x0s)
    lb1, ub1 = problems.unicycle_bounds(sp, x_box=20.0)
    lb2, ub2 = problems.control_box(sp, (-0.6, -0.5), (0.8, 0.4), (-25, -25, -np.inf), (25, 25, np.inf))
    lb = np.where((np.arange(96) % 2 == 0)[:, None], lb1[None, :], lb2[None, :])
    ub = np.where((np.arange(96) % 2 == 0)[:, None], ub1[None, :], ub2[None, :])
    both = solver(x0=torch.as_tensor(w0).cuda(), lbx=lb, ubx=ub, p=torch.as_tensor(p).cuda())
    st = solver.stats()
    assert st["success"]
    x = both["x"].cpu().numpy()
    for sel, l, u in ((slice(0, None, 2), lb1, ub1), (slice(1, None, 2), lb2, ub2)):
        one = solver(x0=w0[sel], lbx=l, ubx=u, p=p[sel])
        assert np.array_equal(solver.stats()["iter_count"], st["iter_count"][sel])
        assert np.abs(one["x"] - x[sel]).max() <= 1e-12
    assert np.abs(x[1::2, 3::5]).max() <= 0.8 + 1e-12                      # the tighter box is honoured
    host = solver(x0=w0, lbx=lb, ubx=ub, p=p)                              # host buffers, per-problem rows
    assert np.abs(host["x"] - x).max() == 0.0


def test_move_blocking_without_du_cost_and_ntu_validation(mv):
    """linear_tracking(..., ntu=K) without R1 must block moves (Trajectory_tracking_lateral_error.py:17-18),
    not silently ignore ntu; the C ABI refuses ntu on a model without u_prev."""
    import ctypes as C
    from mpc_verde_b200 import _lib
    prob = problems.linear_tracking(3, 8, Q=(10.0, 1.0, 0.0), R=0.01, T=0.05, ntu=3)
    sp = prob["spec"]
    assert sp.model == S.MODEL_LINEAR3_DU and sp.R1 == 0.0
    solver = _solver(mv, prob)
    A, Bd = problems.c2d(*problems.lateral_error_matrices(0.6), 0.05)
    rng = np.random.default_rng(2)
    B = 64
    x0 = np.concatenate([rng.normal(size=(B, 3)) * 0.1, np.zeros((B, 1))], 1)
    stage = np.zeros((B, 8, 4))
    stage[:, :, 0] = rng.normal(size=(B, 1)) * 0.2
    p = np.concatenate([x0, np.tile(np.concatenate([A.ravel(), Bd.ravel()]), (B, 1)), stage.reshape(B, -1)], 1)
    lbx, ubx = problems.control_box(solver.spec, -0.3491, 0.3491)
    w0 = problems.cold_start(solver.spec, x0)
    sol = solver(x0=w0, lbx=lbx, ubx=ubx, p=p)
    assert solver.stats()["success"]
    u = sol["x"][:, 4::5][:, :8]
    assert np.abs(u[:, 3:] - u[:, 2:3]).max() <= 1e-12                     # moves 3.. repeat move 2
    ref = O.solve(solver.spec, w0, lbx, ubx, p, nthreads=NCPU)
    assert np.abs(sol["x"] - ref["x"]).max() <= 1e-8
    bad = S.linear_tracking(3, 8, (10.0, 1.0, 0.0), 0.01, T=0.05)
    bad.ntu = 3                                                            # plain model: no u_prev to block against
    lib = _lib.lib()
    assert not lib.mpcv_create(C.byref(bad))
    assert b"ntu" in lib.mpcv_last_error()


def test_acceptable_level_termination_matches_oracle(mv):
    """IPOPT's acceptable-level exit (acceptable_tol / acceptable_iter / acceptable_obj_change_tol of the scripts'
    opts, single_shooting_v1.py:121-129).  Loose settings make it fire; statuses and iteration counts must be the
    oracle's.  With the scripts' own values (acceptable_tol = tol) it never fires before the regular test."""
    import torch
    prob = problems.unicycle_multiple_shooting()
    x0s, p = common.unicycle_batch(512)
    sp0 = prob["spec"]
    lbx, ubx = problems.unicycle_bounds(sp0, x_box=20.0)
    w0 = problems.cold_start(sp0, x0s)
    opts = {"ipopt": {"max_iter": 2000, "acceptable_tol": 1e-3, "acceptable_iter": 2, "acceptable_obj_change_tol": 1e20}}
    solver = mv.nlpsol("solver", "ipopt", prob, opts)
    sol = solver(x0=torch.as_tensor(w0).cuda(), lbx=lbx, ubx=ubx, p=torch.as_tensor(p).cuda())
    st = solver.stats()
    ref = O.solve(solver.spec, w0, lbx, ubx, p, nthreads=NCPU)
    assert (st["status_code"] == 1).sum() > 0 and set(np.unique(st["status_code"])) <= {0, 1}
    assert np.array_equal(st["status_code"], ref["status"]) and np.array_equal(st["iter_count"], ref["iters"])
    assert "Solved_To_Acceptable_Level" in st["return_status"]
    assert np.abs(sol["x"].cpu().numpy() - ref["x"]).max() <= 1e-9
    strict = _solver(mv, prob)
    strict(x0=torch.as_tensor(w0).cuda(), lbx=lbx, ubx=ubx, p=torch.as_tensor(p).cuda())
    assert np.all(strict.stats()["status_code"] == 0)


def test_calls_on_two_streams_through_one_handle_are_ordered(mv):
    """the handle owns the workspaces; two calls on different streams must not overlap."""
    import torch
    solver = _solver(mv, problems.unicycle_multiple_shooting())
    sp = solver.spec
    lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
    outs = []
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    batches = []
    for i, st in enumerate(streams):
        x0s, p = common.unicycle_batch(3000, seed=100 + i)
        batches.append((torch.as_tensor(problems.cold_start(sp, x0s)).cuda(), torch.as_tensor(p).cuda()))
    torch.cuda.synchronize()
    for (w0, p), st in zip(batches, streams):
        with torch.cuda.stream(st):
            outs.append(solver(x0=w0, lbx=lbx, ubx=ubx, p=p, outputs=("x", "f"))["x"])
    torch.cuda.synchronize()
    for (w0, p), x in zip(batches, outs):
        again = solver(x0=w0, lbx=lbx, ubx=ubx, p=p, outputs=("x", "f"))["x"]
        torch.cuda.synchronize()
        assert torch.equal(again, x)


# ---- feasibility restoration, one-sided-bound damping ---------------------------------------------------------------
@pytest.mark.parametrize("layout", ["PHASED", "RESIDENT", "WARP", "THREAD"])
def test_zeros_guess_batch_recovers_through_restoration(mv, layout):
    """the scripts' w0 = 0 from a far-away x0 drives IPOPT into its restoration phase.  The batch
    that ended with 7 of 512 `Restoration_Failed` in round 1 must now solve everywhere, in every layout, on the oracle's
    iteration path (the restoration steps count as iterations)."""
    import torch
    prob = problems.unicycle_multiple_shooting()
    sp = prob["spec"]
    x0s, p = common.unicycle_batch(512, seed=77)
    lbx, ubx = problems.unicycle_bounds(sp)
    solver = _solver(mv, prob, layout=getattr(S, "LAYOUT_" + layout))
    sol = solver(x0=None, lbx=lbx, ubx=ubx, p=torch.as_tensor(p).cuda())
    st = solver.stats()
    ref = O.solve(sp, None, lbx, ubx, p, nthreads=NCPU)
    assert np.all(ref["status"] == 0)
    assert np.all(st["status_code"] == 0), np.unique(st["status_code"], return_counts=True)
    same = st["iter_count"] == ref["iters"]
    assert same.mean() >= 0.97, same.mean()
    assert np.abs(sol["x"].cpu().numpy()[same] - ref["x"][same]).max() <= 1e-8
    assert np.abs(sol["f"].cpu().numpy()[same] - ref["f"][same]).max() <= 1e-8 * np.abs(ref["f"]).max()
    # the filters of this batch grow past the 8 entries the round-1 filter held (oracle: up to 17); the 16-entry
    # device filter with IPOPT's pruning of dominated entries follows them and reports when it had to drop one
    assert solver.diagnostics()["filter_overflows"] <= 2


def test_diagnostics_are_clean_on_the_baseline_batch(mv):
    import torch
    prob = problems.unicycle_multiple_shooting()
    sp = prob["spec"]
    x0s, p = common.unicycle_batch(4096)
    lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
    solver = _solver(mv, prob)
    solver(x0=torch.as_tensor(problems.cold_start(sp, x0s)).cuda(), lbx=lbx, ubx=ubx, p=torch.as_tensor(p).cuda())
    assert solver.stats()["success"] and solver.diagnostics() == {"filter_overflows": 0}


def test_one_sided_bounds_are_damped_like_the_oracle(mv):
    """kappa_d: a variable with only one finite bound gets IPOPT's linear damping term in the barrier objective and
    its gradient.  None of the scripts has such a bound; the C2 batch with v >= -1, w <= pi/4 only exercises it."""
    import torch
    prob = problems.unicycle_multiple_shooting()
    sp = prob["spec"]
    x0s, p = common.unicycle_batch(512)
    lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
    lbx, ubx = np.array(lbx, dtype=float), np.array(ubx, dtype=float)
    nz = sp.nx + sp.nu
    for k in range(sp.N):
        ubx[k * nz + sp.nx] = np.inf          # v: lower bound only
        lbx[k * nz + sp.nx + 1] = -np.inf     # w: upper bound only
    w0 = problems.cold_start(sp, x0s)
    ref = O.solve(sp, w0, lbx, ubx, p, nthreads=NCPU)
    for lay in (S.LAYOUT_PHASED, S.LAYOUT_RESIDENT):
        solver = _solver(mv, prob, layout=lay)
        sol = solver(x0=torch.as_tensor(w0).cuda(), lbx=lbx, ubx=ubx, p=torch.as_tensor(p).cuda())
        st = solver.stats()
        assert np.array_equal(st["status_code"], ref["status"])
        same = (ref["status"] == 0) & (st["iter_count"] == ref["iters"])
        assert same.mean() >= 0.98
        assert np.abs(sol["x"].cpu().numpy()[same] - ref["x"][same]).max() <= 1e-9


def test_pipes_knob_and_batches_in_flight_do_not_change_results(mv):
    """`opts={"pipes": 1}` (mpcv_set_knob "phase_pipes") for callers that keep several batches in flight on several
    solvers: one pipe or four, one batch at a time or three solvers on three streams at once — same bits; knobs are
    refused once the handle has laid out its pipes."""
    import torch
    from mpc_verde_b200 import _lib
    prob = problems.unicycle_multiple_shooting()
    sp = prob["spec"]
    lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
    batches = []
    for i in range(3):
        x0s, p = common.unicycle_batch(40000, seed=300 + i)
        batches.append((torch.as_tensor(problems.cold_start(sp, x0s)).cuda(), torch.as_tensor(p).cuda()))
    ref = _solver(mv, prob)                                    # library default: four pipes above 32,768 problems
    want = [ref(x0=w0, lbx=lbx, ubx=ubx, p=p, outputs=("x", "f"))["x"].clone() for w0, p in batches]
    solvers = [_solver(mv, prob, pipes=1) for _ in batches]
    streams = [torch.cuda.Stream() for _ in batches]
    torch.cuda.synchronize()
    got = []
    for sv, st, (w0, p) in zip(solvers, streams, batches):
        with torch.cuda.stream(st):
            got.append(sv(x0=w0, lbx=lbx, ubx=ubx, p=p, outputs=("x", "f"))["x"])
    torch.cuda.synchronize()
    for a, b, sv in zip(want, got, solvers):
        assert sv.stats()["success"]
        # (the hand-off to the tail kernel depends on the pipe's share: a problem can move between lane widths there)
        assert (a == b).all(dim=1).float().mean() >= 0.999 and (a - b).abs().max() <= 1e-9
    with pytest.raises(_lib.MpcvError):
        solvers[0].set_knob("phase_pipes", 2)                  # after the first solve: -EBUSY
    with pytest.raises(_lib.MpcvError):
        _solver(mv, prob).set_knob("no_such_knob", 1)
