"""The oracle pinned against every golden artefact the reference holds for this path
(SURVEY.md §4 / §8c).  CPU only."""
import math

import numpy as np
import pytest

from mpc_verde_b200 import problems
from mpc_verde_b200 import spec as S
from tests import reference_loops
from oracle import mpc_oracle as O
from tests import common


def test_rk4_known_answer_and_replay_of_1exemplo_2exemplo():
    sp = S.unicycle_multiple_shooting()
    X, _ = O.rollout(sp, np.array([0, 0, 0, 10, 10, 0.0]), np.tile([1.0, math.pi / 4], 10))
    assert np.allclose(X[0, 1], [0.1991785472131909, 0.01567569162852597, 0.1570796326794897], rtol=1e-14, atol=0)
    for name in ("unicycle_ms_1exemplo.csv", "unicycle_ss_2exemplo.csv"):
        g = common.golden(name)
        xs, us = g[1:84, 0:3], g[0:83, 3:5]
        p = np.concatenate([xs, np.tile([10, 10, 0.0], (83, 1))], 1)
        Xb, _ = O.rollout(sp, p, np.tile(us, (1, 10)))
        assert np.abs(Xb[:, 1, :] - g[2:85, 0:3]).max() <= 1e-12


def test_ms_first_solve():
    sp = S.unicycle_multiple_shooting()
    lbx, ubx = problems.unicycle_bounds(sp)
    r = O.solve(sp, None, lbx, ubx, np.array([0, 0, 0, 10, 10, 0.0]), want_stats=True)
    assert r["status"][0] == 0 and r["iters"][0] == 18
    assert r["x"][0, 3] == 1.0 and r["x"][0, 4] == math.pi / 4
    assert abs(r["f"][0] - 1081.5439729) < 1e-6
    assert r["stats"][0, 0] > 0      # inertia correction fires: the exact Hessian is indefinite early on


def test_ms_closed_loop_reproduces_1exemplo():
    g = common.golden("unicycle_ms_1exemplo.csv")
    sp = S.unicycle_multiple_shooting()
    lbx, ubx = problems.unicycle_bounds(sp)
    # fed the script's own (scrambled) warm start, the restated IPOPT follows the reference's
    # iterates: agreement is at round-off level, far inside the 1e-5 / 1e-4 bars
    r = O.closed_loop(sp, [0, 0, 0], [10, 10, 0], None, lbx, ubx, 100, S.WARM_REFERENCE, 0.1)
    assert r["steps"][0] == 84 and r["status"][0] == 0
    assert np.abs(r["controls"][0, :84] - g[:84, 3:5]).max() <= 1e-10
    assert np.abs(r["states"][0, :84] - g[1:, 0:3]).max() <= 1e-10
    assert np.allclose(r["states"][0, 83], [9.896029357313932, 9.999999615669999, 0.001753695673014654], atol=1e-10)
    for mode in (S.WARM_SHIFT, S.WARM_COLD):
        r = O.closed_loop(sp, [0, 0, 0], [10, 10, 0], None, lbx, ubx, 100, mode, 0.1)
        assert r["steps"][0] == 84
        assert np.abs(r["controls"][0, :84] - g[:84, 3:5]).max() <= 1e-5
        assert np.abs(r["states"][0, :84] - g[1:, 0:3]).max() <= 1e-4


def test_ss_closed_loops():
    g = common.golden("unicycle_ss_2exemplo.csv")
    sp = S.unicycle_single_shooting_rk4()
    lbx, ubx = problems.unicycle_bounds(sp)
    r = O.closed_loop(sp, [0, 0, 0], [10, 10, 0], None, lbx, ubx, 100, S.WARM_REFERENCE, 0.1)
    assert r["steps"][0] == 84
    assert np.abs(r["controls"][0, :84] - g[:84, 3:5]).max() <= 1e-5
    assert np.abs(r["states"][0, :84] - g[1:, 0:3]).max() <= 1e-4
    sp1 = S.unicycle_single_shooting_euler()
    r = O.closed_loop(sp1, [0, 0, 0], [10, 10, 0], None, lbx, ubx, 100, S.WARM_REFERENCE, 0.1)
    assert r["steps"][0] == 84          # hard-coded reshape((85,2)) in single_shooting_v1.py:232
    assert abs(np.linalg.norm(r["states"][0, 84] - [10, 10, 0]) - 0.0868) < 1e-3


def test_mpctools_unicycle_3exemplo():
    g = common.golden("unicycle_mpctools_3exemplo.csv")
    sp = S.unicycle_tracking(N=10, T=0.2, M=1, Q=(1, 5, 0.1), R=(1, 1))
    lbx, ubx = problems.unicycle_bounds(sp)
    nst = 87
    ptraj = np.tile([10, 10, 0, 0, 0.0], (1, nst + 10, 1))
    r = O.closed_loop(sp, [0, 0, 0], None, ptraj, lbx, ubx, nst, S.WARM_COLD, 0.0)
    assert np.abs(r["controls"][0, :nst] - g[:nst, 3:5]).max() <= 1e-5
    # plant there is CVODES; with the exact unicycle flow the recorded states agree to its tolerance
    err = 0.0
    for t in range(nst):
        v, w = g[t, 3], g[t, 4]
        x = g[t, 0:3].copy()          # one-step replay from the recorded state
        th = x[2]
        if abs(w) > 1e-12:
            x = x + [v / w * (math.sin(th + 0.2 * w) - math.sin(th)), -v / w * (math.cos(th + 0.2 * w) - math.cos(th)), 0.2 * w]
        else:
            x = x + [0.2 * v * math.cos(th), 0.2 * v * math.sin(th), 0.0]
        err = max(err, np.abs(x - g[t + 1, 0:3]).max())
    assert err < 2e-6


def test_pendulum_c2d_and_closed_loop_vs_golden():
    g = common.golden("pendulum_invertpend.csv")
    sp, lbx, ubx, pglob, A, Bd = common.pendulum_setup(N=50, ntu=5)
    assert np.allclose(Bd.ravel(), [4.83820374e-05, 9.51937011e-03, 9.67801053e-05, 1.90451212e-02], rtol=1e-8)
    assert abs(A[1, 1] - 0.904806299) < 1e-9 and abs(A[3, 2] - 0.383159406) < 1e-9
    # ZOH replay x+ = A x + B u of the whole file
    xr = g[:-1, :4] @ A.T + g[:-1, 4:5] * Bd.ravel()
    assert np.abs(xr - g[1:, :4]).max() <= 1e-12
    nst = 1000
    ptraj = np.tile([10, 0, 0, 0, 0.0], (1, nst + 50, 1))
    r = O.closed_loop(sp, [0, 0, 0, 0, 0], pglob, ptraj, lbx, ubx, nst, S.WARM_REFERENCE, 0.0)
    assert r["status"][0] == 0
    assert abs(r["controls"][0, 0, 0] - (-60.84425718936204)) <= 1e-6
    assert np.abs(r["controls"][0, :nst, 0] - g[:nst, 4]).max() <= 1e-5
    assert np.abs(r["states"][0, :, :4] - g[:, :4]).max() <= 1e-4


def test_lateral_error_lti_and_ltv_closed_loops_vs_dados():
    """Trajectory Tracking/dados2.csv (LTI, Phiref.py:379-381) and dados.csv (LTV variant): the reference's
    own parameter builder (columns yref..deltaref) and the 500-step closed loops.  The reference plant is
    CVODES (integrator tolerance ~1e-6), here exact ZOH."""
    solve = lambda sp, w0, lbx, ubx, p: O.solve(sp, w0, lbx, ubx, p)["x"]
    g2 = common.golden("lateral_lti_dados2.csv")
    u, x, par = common.lateral_error_closed_loop(solve, ltv=False)
    assert np.abs(par[:, 0, :].T - g2[:, 6:10]).max() <= 1e-12          # the `par` builder, bit-level
    assert np.abs(u - g2[:, 3]).max() <= 1e-5
    assert np.abs(x[1:] - g2[:, 0:3]).max() <= 1e-4
    g1 = common.golden("lateral_ltv_dados.csv")
    u, x, _ = common.lateral_error_closed_loop(solve, ltv=True)
    assert np.abs(u - g1[:, 3]).max() <= 1e-5
    assert np.abs(x[1:] - g1[:, 0:3]).max() <= 1e-4


def test_reference_path_generators():
    """Host-side reference builders next to the hot path: lane_change.py's extended path is
    reproduced exactly (its committed output out.csv), the circle builder against its closed form."""
    g = common.golden("lane_change.csv")
    o = common.golden("lane_change_out.csv")
    x, y, c = reference_loops.lane_change_extended(g[:, 0], g[:, 1], g[:, 2])
    assert x.size == o.shape[0] == 2210
    assert np.abs(x - o[:, 0]).max() <= 1e-13 and np.abs(y - o[:, 1]).max() <= 1e-13 and np.array_equal(c, o[:, 2])
    par = reference_loops.circle_reference_par(10, 20, 0.2)
    assert np.allclose(par[:, 3, 5], [np.cos(0.1 * 1.6), np.sin(0.1 * 1.6), np.pi / 2 + 0.16, 1, 1])


def test_result_sinks_and_error_log():
    """The table the unicycle scripts dump, rebuilt from closed_loop-shaped arrays, equals 1exemplo.xlsx; the
    LTV tracker's error log runs on the dados.csv loop; the test2.py builder keeps its swapped entries."""
    from mpc_verde_b200 import sinks
    g = common.golden("unicycle_ms_1exemplo.csv")
    states, controls = g[1:, 0:3], g[:84, 3:5]            # what closed_loop returns for this run (84 MPC steps)
    tab = sinks.unicycle_table(np.vstack([states[0:1], g[2:, 0:3], g[-1:, 0:3]]), controls, 0.2, 84)
    assert tab.shape == (85, 6)
    assert np.abs(tab[:, 0:5] - g[:, 0:5]).max() <= 1e-15 and np.abs(tab[:, 5] - g[:, 5]).max() <= 1e-9
    solve = lambda sp, w0, lbx, ubx, p: O.solve(sp, w0, lbx, ubx, p)["x"]
    u, x, par = common.lateral_error_closed_loop(solve, ltv=True, nsim=120)
    lc = common.golden("lane_change.csv")
    e = sinks.lateral_tracking_errors(x, u, par, lc[:, 0], lc[:, 1], lc[:, 2], 0.05)
    assert e["path"].shape == (2, 120) and 0 <= e["mse"] < 5 and e["max"] >= e["dist"] > 0
    p = reference_loops.frenet_reference_par(lc[:, 0], lc[:, 1], lc[:, 2], 20, 0.05, t=10)
    assert p.shape == (20, 4) and np.allclose(p[:, 2], lc[10:30, 2]) and np.all(p[:, 3] >= 0)
