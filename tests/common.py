"""Shared fixtures of the parity tests: seeded synthetic batches and golden data."""
import math
import os

import numpy as np

from mpc_verde_b200 import problems
from mpc_verde_b200 import spec as S
from tests import reference_loops

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.loadtxt(os.path.join(GOLDEN, name), delimiter=",", skiprows=1)


def unicycle_batch(B, seed=20261):
    """C2: random initial states, target (10,10,0), cold start X_k = x0, U = 0."""
    rng = np.random.default_rng(seed)
    x0s = np.stack([rng.uniform(-2, 12, B), rng.uniform(-2, 12, B), rng.uniform(-math.pi, math.pi, B)], 1)
    p = np.concatenate([x0s, np.tile([10.0, 10.0, 0.0], (B, 1))], 1)
    return x0s, p


def pendulum_setup(N=50, T=0.01, ntu=5, discretisation="c2d"):
    """Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:10-64."""
    if discretisation == "c2d":
        A, Bd = problems.c2d(problems.PENDULUM_AC, problems.PENDULUM_BC, T)
    else:
        A, Bd = problems.rk4_linear(problems.PENDULUM_AC, problems.PENDULUM_BC, T)
    sp = S.linear_tracking(4, N, Q=(1.2 ** 2, 0.0, 1.0, 0.0), R=0.0, T=T, R1=0.01 ** 2, ntu=ntu)
    lbx, ubx = problems.control_box(sp, -200.0, 200.0)
    pglob = np.concatenate([A.ravel(), Bd.ravel()])
    return sp, lbx, ubx, pglob, A, Bd


def pendulum_batch(sp, pglob, B, seed=20262):
    rng = np.random.default_rng(seed)
    x0 = np.stack([rng.uniform(-1, 1, B), rng.uniform(-0.5, 0.5, B), rng.uniform(-0.2, 0.2, B),
                   rng.uniform(-0.5, 0.5, B), np.zeros(B)], 1)
    stage = np.tile([10.0, 0.0, 0.0, 0.0, 0.0], sp.N)
    p = np.concatenate([x0, np.tile(pglob, (B, 1)), np.tile(stage, (B, 1))], 1)
    return x0, p


def lateral_error_closed_loop(solve_fn, ltv, nsim=None):
    """The MPC loop of Trajectory Tracking/Phiref.py:156-200 (LTI; the commented block :157-171 is the LTV
    variant that produced dados.csv): Nt=5, Ntu=1 (one free move, then blocked), Q=diag(10,1,0), R=0.01,
    |delta| <= 0.3491, uprev = 0, solver rebuilt every step from the plant state, exact-ZOH plant.
    solve_fn(spec, w0[B,n], lbx, ubx, p[B,n_p]) -> x[B,n].  Returns (u [Nsim], x [Nsim+1,3], par)."""
    g = golden("lane_change.csv")
    a, b, c = g[:, 0], g[:, 1], g[:, 2]
    Nt, Delta = 5, 0.05
    Nsim = a.size if nsim is None else nsim
    sp = S.linear_tracking(3, Nt, Q=(10.0, 1.0, 0.0), R=0.01, T=Delta, R1=0.0, ntu=1)
    lbx, ubx = problems.control_box(sp, -0.3491, 0.3491)
    par = reference_loops.lateral_error_par(a, b, Nt, Delta)
    x = np.zeros((Nsim + 1, 3))
    u = np.zeros(Nsim)
    for t in range(Nsim):
        Ac, Bc = problems.lateral_error_matrices(c[t] if ltv else c.mean())
        A, Bd = problems.c2d(Ac, Bc, Delta)
        xa = np.concatenate([x[t], [0.0]])                       # state augmented with uprev = 0
        p = np.concatenate([xa, A.ravel(), Bd.ravel(), par[:, :, t].T.ravel()])[None, :]
        sol = solve_fn(sp, problems.cold_start(sp, xa[None, :]), lbx, ubx, p)
        u[t] = sol[0, sp.nx]
        x[t + 1] = A @ x[t] + Bd[:, 0] * u[t]
    return u, x, par
