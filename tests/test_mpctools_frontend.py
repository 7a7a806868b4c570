"""The MPCTools-shaped front end (mpc_verde_b200/mpctools.py) driven the way the reference scripts drive
MPCTools.  Host logic on the CPU (the oracle stands in for the GPU solver: the tests patch solver.nlpsol);
the same loops on the GPU in test_gpu_parity-style tests below (marked gpu)."""
import numpy as np
import pytest

from mpc_verde_b200 import mpctools as mpc
from mpc_verde_b200 import solver as solver_mod
from mpc_verde_b200 import problems
from oracle import mpc_oracle as O
from tests import common


class _OracleSolver:
    """NlpSolver interface over the CPU oracle — test infrastructure only."""

    def __init__(self, name, plugin, prob, opts=None, device=None):       # the signature of solver.nlpsol
        self.spec = prob["spec"].copy()

    def __call__(self, x0, lbx, ubx, p, outputs=("x", "f")):
        self._r = O.solve(self.spec, x0, lbx, ubx, p)
        return {"x": self._r["x"], "f": self._r["f"]}

    def stats(self):
        st = self._r["status"]
        return {"return_status": ["Solve_Succeeded" if s == 0 else "Failed" for s in st], "iter_count": self._r["iters"],
                "success": bool(np.all(st == 0))}


def pendulum_script(nsim, batch=None):
    """Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:10-79, line by line."""
    Nx, Nu, T, Nt = 4, 1, 0.01, 50
    A, B = mpc.util.c2d(problems.PENDULUM_AC, problems.PENDULUM_BC, T)
    umax = 200
    Dulb, Duub, Dub = np.tile(-np.inf, (5, 1)), np.tile(np.inf, (5, 1)), np.tile(0, (45, 1))
    lb = {"u": np.array([-umax]), "Du": np.vstack((Dulb, Dub))}
    ub = {"u": np.array([umax]), "Du": np.vstack((Duub, Dub))}
    # stage cost (1.2 (x1-10))^2 + (1 x3)^2 + (0.01 du)^2  (:47-55; the weights are squared inside)
    model = problems.linear_tracking(4, Nt, Q=(1.2 ** 2, 0.0, 1.0, 0.0), R=0.0, T=T, R1=0.01 ** 2, ntu=5)
    x0 = np.array([0, 0, 0, 0.0]) if batch is None else np.tile([0, 0, 0, 0.0], (batch, 1))
    N = {"x": Nx, "u": Nu, "t": Nt}
    solver = mpc.nmpc(model, None, N, x0, lb, ub, p=np.tile([10.0, 0, 0, 0, 0], (Nt, 1)), uprev=np.array([0]),
                      pglob=np.concatenate([A.ravel(), B.ravel()]), isQP=True, verbosity=0)
    xcl = np.zeros((Nx, nsim + 1))
    ucl = np.zeros((Nu, nsim))
    for k in range(nsim):
        solver.fixvar("x", 0, x0)
        sol = mpc.callSolver(solver)
        assert (sol["status"] if batch is None else sol["status"][0]) == "Solve_Succeeded"
        xs, us = (sol["x"], sol["u"]) if batch is None else (sol["x"][0], sol["u"][0])
        xcl[:, k] = xs[0, :]
        ucl[:, k] = us[0, :]
        step = lambda x, u: A @ x + B[:, 0] * u
        x0 = step(x0, ucl[0, k]) if batch is None else np.tile(step(x0[0], ucl[0, k]), (batch, 1))
    xcl[:, nsim] = x0 if batch is None else x0[0]
    return xcl, ucl


def test_pendulum_script_through_the_front_end_cpu(monkeypatch):
    monkeypatch.setattr(solver_mod, "nlpsol", _OracleSolver)
    g = common.golden("pendulum_invertpend.csv")
    xcl, ucl = pendulum_script(40)
    assert abs(ucl[0, 0] - (-60.84425718936204)) <= 1e-5
    assert np.abs(ucl[0] - g[:40, 4]).max() <= 1e-5
    assert np.abs(xcl.T - g[:41, :4]).max() <= 1e-4
    xb, ub = pendulum_script(3, batch=4)                  # leading batch dimension
    assert np.abs(ub[0] - g[:3, 4]).max() <= 1e-5


def test_controlsolver_surface_cpu(monkeypatch):
    """par / solve / stats / saveguess / fixvar(x1) / var — the loop of Trajectory Tracking/Trajectory_tracking.py:101-118."""
    monkeypatch.setattr(solver_mod, "nlpsol", _OracleSolver)
    Nt = 10
    model = problems.unicycle_tracking(N=Nt, T=0.2, M=1)
    lb = {"u": np.array([-1, -np.pi / 4]), "x": np.array([-20, -2, -np.inf])}
    ub = {"u": np.array([1, np.pi / 4]), "x": np.array([20, 2, np.inf])}
    solver = mpc.nmpc(model, None, {"x": 3, "u": 2, "t": Nt, "p": 5}, np.array([1.0, 0.0, np.pi / 2]), lb, ub,
                      p=np.zeros((Nt, 5)))
    xs = []
    for t in range(3):
        for k in range(Nt):
            tt = (t + k) * 0.2
            solver.par["p", k] = np.array([np.cos(0.1 * tt), np.sin(0.1 * tt), np.pi / 2 + 0.1 * tt, 0.1, 0.1])
        solver.solve()
        assert solver.stats["status"] == "Solve_Succeeded"
        solver.saveguess()
        solver.fixvar("x", 0, solver.var["x", 1])
        assert solver.var["u", 0, :].shape == (2,) and solver.var["x", :, :].shape == (Nt + 1, 3)
        xs.append(solver.var["x", 1].copy())
    # x0 of step t+1 is the predicted x1 of step t, and the shifted guess keeps the trajectory feasible
    assert np.abs(solver._x0[0] - xs[-1]).max() == 0.0
    with pytest.raises(TypeError):
        mpc.nmpc(lambda x, u: x, None, {"x": 3, "u": 2, "t": Nt}, np.zeros(3), lb, ub)


@pytest.mark.gpu
def test_pendulum_script_through_the_front_end_gpu():
    g = common.golden("pendulum_invertpend.csv")
    xcl, ucl = pendulum_script(300)
    assert np.abs(ucl[0] - g[:300, 4]).max() <= 1e-5
    assert np.abs(xcl.T - g[:301, :4]).max() <= 1e-4
