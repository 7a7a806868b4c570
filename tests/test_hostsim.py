"""The device headers (hand-derived sweeps + Riccati interior point) compiled as plain C++ with
one lane per problem, against the oracle (generic AD + condensed Cholesky).  This is a
development harness for a container without a GPU — the product never runs on the CPU."""
import math

import numpy as np
import pytest

from mpc_verde_b200 import problems
from mpc_verde_b200 import spec as S
from oracle import mpc_oracle as O
from tests import common
from tests.hostsim import hostsim as H

MODELS = [
    ("quad", lambda: S.unicycle_multiple_shooting()),
    ("euler", lambda: S.unicycle_single_shooting_euler()),
    ("node", lambda: S.unicycle_tracking(M=1)),
    ("node_m3", lambda: S.unicycle_tracking(M=3)),
    ("lin3", lambda: S.linear_tracking(3, 5, (10, 1, 0.5, 0), 0.01)),
    ("lin4", lambda: S.linear_tracking(4, 5, (1, 2, 3, 4), 1.0)),
    ("lin4du", lambda: S.linear_tracking(4, 5, (1.44, 0, 1, 0), 0.0, R1=1e-4)),
    ("lin3du", lambda: S.linear_tracking(3, 5, (10, 1, 0, 0), 0.01, R1=0.5)),
    ("frenet", lambda: S.frenet_bicycle(N=20, T=0.05, M=1)),
    ("frenet_m2", lambda: S.frenet_bicycle(N=5, T=0.1, M=2)),
]


@pytest.mark.parametrize("name,mk", MODELS)
def test_hand_derivatives_vs_ad(name, mk):
    sp = mk()
    rng = np.random.default_rng(1)
    B = 64
    z = rng.normal(size=(B, sp.nx + sp.nu)) * 2
    ps = rng.normal(size=(B, max(sp.npg + sp.nps, 1)))
    if name.startswith("frenet"):
        z *= 0.15                     # stay away from the poles of tan(delta) and of 1/(1-(y-yt) kappat)
        ps *= 0.1
    lam = rng.normal(size=(B, sp.nx)) * 3
    a, b = O.stage_derivs(sp, z, ps, lam), H.stage_derivs(sp, z, ps, lam)
    for k in a:
        assert np.abs(a[k] - b[k]).max() <= 1e-13 * (1 + np.abs(a[k]).max()), k


def test_riccati_ipm_equals_condensed_ipm_on_random_batch():
    sp = S.unicycle_multiple_shooting()
    x0s, p = common.unicycle_batch(200)
    lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
    w0 = problems.cold_start(sp, x0s)
    a, b = O.solve(sp, w0, lbx, ubx, p), H.solve(sp, w0, lbx, ubx, p)
    assert np.all(a["status"] == 0) and np.all(b["status"] == 0)
    assert np.mean(a["iters"] == b["iters"]) >= 0.99
    assert np.abs(a["x"] - b["x"]).max() <= 1e-9
    assert np.abs(a["f"] - b["f"]).max() <= 1e-9 * np.abs(a["f"]).max()


def test_zero_guess_far_from_target_exercises_backtracking_and_soc():
    sp = S.unicycle_multiple_shooting()
    x0s, p = common.unicycle_batch(100, seed=77)
    lbx, ubx = problems.unicycle_bounds(sp)
    a = O.solve(sp, None, lbx, ubx, p, want_stats=True)          # all-zeros guess (the script's w0 = 0)
    b = H.solve(sp, None, lbx, ubx, p)
    assert a["stats"][:, 1].sum() > 0                              # backtracks happened
    ok = (a["status"] == 0) & (b["status"] == 0)
    assert ok.mean() > 0.9
    assert np.array_equal(a["status"], b["status"])
    same = ok & (a["iters"] == b["iters"])
    assert same.mean() > 0.95
    assert np.abs(a["x"][same] - b["x"][same]).max() <= 1e-8


@pytest.mark.parametrize("mode", [S.WARM_REFERENCE, S.WARM_SHIFT, S.WARM_COLD])
def test_closed_loops_match_oracle_and_golden(mode):
    g = common.golden("unicycle_ms_1exemplo.csv")
    sp = S.unicycle_multiple_shooting()
    lbx, ubx = problems.unicycle_bounds(sp)
    a = O.closed_loop(sp, [0, 0, 0], [10, 10, 0], None, lbx, ubx, 100, mode, 0.1)
    b = H.closed_loop(sp, [0, 0, 0], [10, 10, 0], None, lbx, ubx, 100, mode, 0.1)
    assert a["steps"][0] == b["steps"][0] == 84
    assert a["iters"][0] == b["iters"][0]
    assert np.abs(a["controls"] - b["controls"]).max() <= 1e-10
    assert np.abs(b["controls"][0, :84] - g[:84, 3:5]).max() <= 1e-5
    for spx in (S.unicycle_single_shooting_rk4(), S.unicycle_single_shooting_euler()):
        lb2, ub2 = problems.unicycle_bounds(spx)
        a = O.closed_loop(spx, [0, 0, 0], [10, 10, 0], None, lb2, ub2, 100, mode, 0.1)
        b = H.closed_loop(spx, [0, 0, 0], [10, 10, 0], None, lb2, ub2, 100, mode, 0.1)
        assert a["steps"][0] == b["steps"][0] == 84
        assert np.abs(a["controls"] - b["controls"]).max() <= 1e-9


def test_pendulum_move_blocking_closed_loop():
    g = common.golden("pendulum_invertpend.csv")
    sp, lbx, ubx, pglob, _, _ = common.pendulum_setup(N=50, ntu=5)
    nst = 200
    ptraj = np.tile([10, 0, 0, 0, 0.0], (1, nst + 50, 1))
    b = H.closed_loop(sp, [0, 0, 0, 0, 0], pglob, ptraj, lbx, ubx, nst, S.WARM_REFERENCE, 0.0)
    assert b["status"][0] == 0
    assert np.abs(b["controls"][0, :nst, 0] - g[:nst, 4]).max() <= 1e-5
    assert np.abs(b["states"][0, :nst + 1, :4] - g[:nst + 1, :4]).max() <= 1e-4


def _same(a, b):
    return all(np.array_equal(a[k], b[k]) for k in ("x", "f", "g", "lam_g", "lam_x", "status", "iters"))


def test_phase_pipeline_schedule_is_bit_identical_to_one_kernel_solve():
    """mpcv_phase.cuh runs the SAME phase functions as Ipm::solve(), one batch-wide launch per phase with
    the scalar state parked in the workspace, per-stage costs reduced in the per-problem phase and an
    active list compacted every sweep.  Replayed on the CPU the results must not differ in one bit."""
    sp = S.unicycle_multiple_shooting()
    x0s, p = common.unicycle_batch(200)
    lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
    w0 = problems.cold_start(sp, x0s)
    a, b = H.solve(sp, w0, lbx, ubx, p), H.solve(sp, w0, lbx, ubx, p, phased=True)
    assert np.all(a["status"] == 0) and _same(a, b)
    assert b["sweeps"] == a["iters"].max() + 1          # one extra sweep detects the last convergence
    # the script's w0 = 0 far from the target: backtracking, second-order corrections, failures
    x0s, p = common.unicycle_batch(100, seed=77)
    lbx, ubx = problems.unicycle_bounds(sp)
    a, b = H.solve(sp, None, lbx, ubx, p), H.solve(sp, None, lbx, ubx, p, phased=True)
    assert _same(a, b)
    # linear model with input-increment cost and move blocking (fixed variables, chain-rule folding)
    sp3, lbx3, ubx3, pglob3, _, _ = common.pendulum_setup(N=50, ntu=5)
    x0, pp = common.pendulum_batch(sp3, pglob3, 8)
    w = problems.cold_start(sp3, x0)
    assert _same(H.solve(sp3, w, lbx3, ubx3, pp), H.solve(sp3, w, lbx3, ubx3, pp, phased=True))
    # per-stage reference parameters
    spt = S.unicycle_tracking(N=20, T=0.05, M=1)
    rng = np.random.default_rng(4)
    B = 32
    t = np.arange(spt.N) * spt.T
    x0 = np.stack([1 + rng.normal(size=B) * 0.1, rng.normal(size=B) * 0.1, math.pi / 2 + rng.normal(size=B) * 0.1], 1)
    stage = np.stack([np.cos(0.1 * t), np.sin(0.1 * t), math.pi / 2 + 0.1 * t, np.ones_like(t) * 0.1, np.ones_like(t) * 0.1], 1)
    pt = np.concatenate([x0, np.tile(stage.ravel(), (B, 1))], 1)
    lbt, ubt = problems.control_box(spt, (-1, -math.pi / 4), (1, math.pi / 4), (-20, -2, -np.inf), (20, 2, np.inf))
    wt = problems.cold_start(spt, x0)
    assert _same(H.solve(spt, wt, lbt, ubt, pt), H.solve(spt, wt, lbt, ubt, pt, phased=True))


def frenet_batch(sp, B, seed=20265):
    """Synthetic C4-secondary batch: lane-change-like references (Trajectory Tracking/test2.py:79-100 unpacks
    p = (yt, phit, kappat, vdes)), random lateral / heading / speed offsets."""
    rng = np.random.default_rng(seed)
    t = np.arange(sp.N) * sp.T
    amp = rng.uniform(0.5, 1.5, (B, 1))
    yt = amp * 0.5 * (1 - np.cos(0.8 * t))[None, :]
    phit = amp * 0.4 * np.sin(0.8 * t)[None, :] * 0.2
    kap = amp * 0.02 * np.cos(0.8 * t)[None, :]
    vdes = np.tile(rng.uniform(0.4, 0.8, (B, 1)), (1, sp.N))
    stage = np.stack([yt, phit, kap, vdes], 2).reshape(B, -1)
    x0 = np.stack([rng.normal(size=B) * 0.1, rng.normal(size=B) * 0.05, vdes[:, 0] + rng.normal(size=B) * 0.05,
                   np.zeros(B)], 1)
    return x0, np.concatenate([x0, stage], 1)


def test_frenet_bicycle_ipm_vs_oracle():
    sp = S.frenet_bicycle(N=20, T=0.05, M=1)
    lbx, ubx = problems.frenet_bounds(sp)
    x0, p = frenet_batch(sp, 48)
    w0 = problems.cold_start(sp, x0)
    a, b, c = O.solve(sp, w0, lbx, ubx, p), H.solve(sp, w0, lbx, ubx, p), H.solve(sp, w0, lbx, ubx, p, phased=True)
    assert np.all(a["status"] == 0) and np.all(b["status"] == 0)
    assert np.mean(a["iters"] == b["iters"]) >= 0.95
    assert np.abs(a["x"] - b["x"]).max() <= 1e-8
    assert np.abs(a["f"] - b["f"]).max() <= 1e-9 * (1 + np.abs(a["f"]).max())
    assert _same(b, c)


def test_restoration_and_one_sided_damping_oracle_vs_device_headers():
    """Round 2: feasibility restoration (the w0 = 0 batch of SURVEY Appendix E has no failure left) and the kappa_d
    damping of one-sided bounds, oracle (null-space condensing) against the device headers (Riccati), both schedules."""
    sp = S.unicycle_multiple_shooting()
    x0s, p = common.unicycle_batch(160, seed=77)
    lbx, ubx = problems.unicycle_bounds(sp)
    a, b, c = O.solve(sp, None, lbx, ubx, p), H.solve(sp, None, lbx, ubx, p), H.solve(sp, None, lbx, ubx, p, phased=True)
    assert np.all(a["status"] == 0) and np.all(b["status"] == 0)
    assert a["iters"].max() > 40                                   # the restoration cases are in the sample
    same = a["iters"] == b["iters"]
    assert same.mean() > 0.95
    assert np.abs(a["x"][same] - b["x"][same]).max() <= 1e-8
    assert _same(b, c)
    x0s, p = common.unicycle_batch(100)
    lbx, ubx = (np.array(v, dtype=float) for v in problems.unicycle_bounds(sp, x_box=20.0))
    nz = sp.nx + sp.nu
    for k in range(sp.N):
        ubx[k * nz + sp.nx] = np.inf
        lbx[k * nz + sp.nx + 1] = -np.inf
    w0 = problems.cold_start(sp, x0s)
    a, b = O.solve(sp, w0, lbx, ubx, p), H.solve(sp, w0, lbx, ubx, p)
    assert np.all(a["status"] == 0) and np.array_equal(a["iters"], b["iters"])
    assert np.abs(a["x"] - b["x"]).max() <= 1e-9


def test_filter_follows_ipopt_beyond_eight_entries():
    """the device filter used to hold 8 entries and evict the oldest, IPOPT's is unbounded.  Now both the
    oracle and the device headers prune dominated entries the way IPOPT's Filter::AddEntry does; the oracle stays
    unbounded, the device filter holds 16 and counts overflows.  On the hard all-zeros-guess batch the oracle's
    filters grow past 8 (that is what the old cap truncated) and the two implementations still walk the same path."""
    sp = S.unicycle_multiple_shooting()
    x0s, p = common.unicycle_batch(768, seed=77)
    lbx, ubx = problems.unicycle_bounds(sp)
    a = O.solve(sp, None, lbx, ubx, p, want_stats=True)
    before = H.filter_overflows()
    b = H.solve(sp, None, lbx, ubx, p)
    big = a["stats"][:, 6] > 8
    assert big.sum() >= 5                                      # the sample does exercise filters beyond the old cap
    assert np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["iters"][big], b["iters"][big])
    assert H.filter_overflows() - before <= int((a["stats"][:, 6] > 16).sum()) + 1
    # the BASELINE workload never comes near the cap
    x0s, p = common.unicycle_batch(1024)
    lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
    c = O.solve(sp, problems.cold_start(sp, x0s), lbx, ubx, p, want_stats=True)
    assert c["stats"][:, 6].max() <= 8
