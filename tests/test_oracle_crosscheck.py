"""Independent cross-checks of the oracle so that it is not trusted on its own word:
SciPy SLSQP on the same NLP (cost), and a dense symmetric-indefinite LDL^T of the assembled KKT
matrix (what IPOPT+MUMPS factorise) against the oracle's condensed step and inertia verdict."""
import math

import numpy as np
import pytest
from scipy.linalg import ldl
from scipy.optimize import minimize

from mpc_verde_b200 import problems
from mpc_verde_b200 import spec as S
from oracle import mpc_oracle as O
from tests import common


def test_scipy_slsqp_agrees_on_single_shooting_cost():
    sp = S.unicycle_single_shooting_rk4()
    lbx, ubx = problems.unicycle_bounds(sp)
    x0s, p = common.unicycle_batch(4, seed=3)
    ref = O.solve(sp, None, lbx, ubx, p)
    for b in range(4):
        fun = lambda U: O.rollout(sp, p[b], U)[1][0]
        best = np.inf
        for start in (ref["x"][b] * 0.9, np.zeros(20)):
            r = minimize(fun, start, method="SLSQP", bounds=list(zip(lbx, ubx)), options={"ftol": 1e-14, "maxiter": 500})
            best = min(best, r.fun)
        assert best >= ref["f"][b] * (1 - 1e-6)            # the IPM point is at least as good a local minimum
        assert abs(best - ref["f"][b]) <= 1e-5 * abs(ref["f"][b])


def _assemble_kkt(sp, w, lam, p, sigma, dw):
    """Dense KKT [[W+Sigma+dw I, J'],[J, 0]] of the multiple-shooting NLP from the oracle's stage blocks."""
    nx, nu, N = sp.nx, sp.nu, sp.N
    nz = nx + nu
    n, m = sp.n_var, sp.n_g
    K = np.zeros((n + m, n + m))
    pg = p[nx:nx + sp.npg]
    for k in range(N):
        z = w[k * nz:(k + 1) * nz][None, :]
        d = O.stage_derivs(sp, z, pg[None, :] if sp.npg else None, lam[(k + 1) * nx:(k + 2) * nx][None, :])
        K[k * nz:(k + 1) * nz, k * nz:(k + 1) * nz] += d["H"][0]
        r0 = n + (k + 1) * nx
        K[r0:r0 + nx, k * nz:k * nz + nx] = d["A"][0]
        K[r0:r0 + nx, k * nz + nx:(k + 1) * nz] = d["B"][0]
        K[r0:r0 + nx, (k + 1) * nz:(k + 1) * nz + nx] -= np.eye(nx)
    K[n:n + nx, 0:nx] = -np.eye(nx)
    K[:n, :n] += np.diag(sigma + dw)
    K[:n, n:] = K[n:, :n].T
    return K


def _inertia(K):
    _, D, _ = ldl(K)
    ev = np.linalg.eigvalsh(D)
    return int((ev > 0).sum()), int((ev < 0).sum())


def test_riccati_condition_equals_kkt_inertia():
    """'Reduced Hessian positive definite' (what oracle and kernel test) <=> KKT inertia (n, m, 0)
    (what IPOPT asks of MUMPS)."""
    sp = S.unicycle_multiple_shooting()
    rng = np.random.default_rng(0)
    n, m = sp.n_var, sp.n_g
    p = np.array([0, 0, 0, 10, 10, 0.0])
    hits = {True: 0, False: 0}
    for trial in range(40):
        w = rng.normal(size=n) * (0.3 if trial % 2 else 3.0)
        lam = rng.normal(size=m) * (0.1 if trial % 2 else 30.0)
        sigma = np.zeros(n)
        for dw in (0.0, 1e-2, 10.0):
            K = _assemble_kkt(sp, w, lam, p, sigma, dw)
            pos, neg = _inertia(K)
            correct = (pos == n and neg == m)
            # reduced Hessian on the null space of J
            J = K[n:, :n]
            _, _, Vt = np.linalg.svd(J)
            Z = Vt[m:].T
            red = Z.T @ K[:n, :n] @ Z
            pd = bool(np.all(np.linalg.eigvalsh(red) > 0))
            assert pd == correct
            hits[correct] += 1
    assert hits[True] > 5 and hits[False] > 5          # both verdicts exercised


def test_oracle_first_iterate_direction_matches_dense_kkt():
    """One Newton step: oracle's solution after max_iter=1 equals the dense-KKT step."""
    sp = S.unicycle_multiple_shooting(opts={"ipopt": {"max_iter": 1}})
    lbx, ubx = problems.unicycle_bounds(sp)
    p = np.array([1.0, 2.0, 0.3, 10, 10, 0.0])
    w0 = problems.cold_start(sp, p[:3])[0]
    r1 = O.solve(sp, w0, lbx, ubx, p)
    n, m = sp.n_var, sp.n_g
    # rebuild iterate 0 in relaxed-bound terms
    relax = lambda b: b + np.sign(b) * 1e-8 * np.maximum(1, np.abs(b))
    hasb = np.isfinite(lbx)
    lo, hi = np.where(hasb, relax(lbx), -np.inf), np.where(hasb, relax(ubx), np.inf)
    w = w0.copy()
    push = np.minimum(1e-2 * np.maximum(1, np.abs(lo[hasb])), 1e-2 * (hi[hasb] - lo[hasb]))
    w[hasb] = np.minimum(np.maximum(w[hasb], lo[hasb] + push), hi[hasb] - push)
    mu = 0.1
    sigma = np.zeros(n)
    sigma[hasb] = 1.0 / (w[hasb] - lo[hasb]) + 1.0 / (hi[hasb] - w[hasb])
    # objective gradient and residuals at the pushed start
    nz = 5
    grad = np.zeros(n)
    c = np.zeros(m)
    c[:3] = p[:3] - w[:3]
    for k in range(10):
        d = O.stage_derivs(sp, w[k * nz:(k + 1) * nz][None, :], p[None, 3:6], np.zeros((1, 3)))
        grad[k * nz:(k + 1) * nz] = d["grad"][0]
        c[3 * (k + 1):3 * (k + 2)] = d["xn"][0] - w[(k + 1) * nz:(k + 1) * nz + 3]
    # least-squares multipliers [I J'; J 0][.; lam] = -[grad f - zl + zu; 0] with z = 1 (cancels for boxes)
    J = _assemble_kkt(sp, w, np.zeros(m), p, np.zeros(n), 0.0)[n:, :n]
    KLS = np.block([[np.eye(n), J.T], [J, np.zeros((m, m))]])
    lam = np.linalg.solve(KLS, -np.concatenate([grad, np.zeros(m)]))[n:]
    assert np.abs(lam).max() <= 1e3
    r = grad.copy()
    r[hasb] += -mu / (w[hasb] - lo[hasb]) + mu / (hi[hasb] - w[hasb])
    dw = 0.0
    for _ in range(30):
        K = _assemble_kkt(sp, w, lam, p, sigma, dw)
        pos, neg = _inertia(K)
        if pos == n and neg == m:
            break
        dw = 1e-4 if dw == 0.0 else dw * 100.0
    sol = np.linalg.solve(K, -np.concatenate([r, c]))
    d = sol[:n]
    alpha = 1.0
    tau = 0.99
    neg_dir = hasb & (d < 0)
    pos_dir = hasb & (d > 0)
    if neg_dir.any():
        alpha = min(alpha, (-tau * (w[neg_dir] - lo[neg_dir]) / d[neg_dir]).min())
    if pos_dir.any():
        alpha = min(alpha, (tau * (hi[pos_dir] - w[pos_dir]) / d[pos_dir]).min())
    w1 = np.clip(w + alpha * d, lbx, ubx)
    # the oracle may have backtracked; accept alpha, alpha/2, alpha/4
    errs = [np.abs(np.clip(w + alpha * 0.5 ** j * d, lbx, ubx) - r1["x"][0]).max() for j in range(4)]
    assert min(errs) < 1e-8, errs
