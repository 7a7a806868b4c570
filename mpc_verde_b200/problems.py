"""`nlp_prob = {'f','x','g','p'}` templates of the reference scripts.

The reference builds these four entries as CasADi SX expressions; here they are structural
descriptions (layout of the decision vector, constraint rows and parameters) bound to a
`Spec`.  The decision-vector layouts are the scripts' own, so `args['x0']`, `lbx/ubx` and the
slicing of `sol['x']` carry over unchanged.
"""
import math

import numpy as np

from . import spec as S


def _prob(spec, f_doc, x_doc, g_doc, p_doc):
    return {"f": f_doc, "x": x_doc, "g": g_doc, "p": p_doc, "spec": spec}


def unicycle_multiple_shooting(N=10, T=0.2, M=4, Q=S.UNICYCLE_Q, R=S.UNICYCLE_R):
    """Casadi/multiple_shooting_casadi.py:116-187."""
    sp = S.unicycle_multiple_shooting(N, T, M, Q, R)
    return _prob(sp, "sum_k RK4-quadrature of (x-ref)'Q(x-ref)+u'Ru over interval k",
                 "[X0(3),U0(2),X1(3),U1(2),...,U_{N-1}(2),X_N(3)]",
                 "[P[:3]-X0 ; F(X_k,U_k)-X_{k+1}]", "[x_init(3); x_ref(3)]")


def unicycle_single_shooting_rk4(N=10, T=0.2, M=4, Q=S.UNICYCLE_Q, R=S.UNICYCLE_R):
    """Casadi/single_shooting_v2.py:115-167."""
    sp = S.unicycle_single_shooting_rk4(N, T, M, Q, R)
    return _prob(sp, "sum_k RK4-quadrature cost along the rollout", "[U0(2),...,U_{N-1}(2)]",
                 "predicted states (inert rows, bounds +-inf)", "[x_init(3); x_ref(3)]")


def unicycle_single_shooting_euler(N=10, T=0.2, Q=S.UNICYCLE_Q, R=S.UNICYCLE_R):
    """Casadi/single_shooting_v1.py:97-119."""
    sp = S.unicycle_single_shooting_euler(N, T, Q, R)
    return _prob(sp, "sum_{k<N} (x_k-ref)'Q(x_k-ref)+u_k'Ru_k along the Euler rollout", "[U0(2),...,U_{N-1}(2)]",
                 "vec(X) (inert rows, bounds +-inf)", "[x_init(3); x_ref(3)]")


def unicycle_tracking(N=10, T=0.2, M=1, Q=(1.0, 1.0, 0.1), R=S.UNICYCLE_R):
    """Trajectory Tracking/Trajectory_tracking.py:54-72 (MPCTools nmpc with per-stage p)."""
    sp = S.unicycle_tracking(N, T, M, Q, R)
    return _prob(sp, "sum_{k<N} (x_k-p_k[:3])'Q(.)+(u_k-p_k[3:])'R(.)", "[X0,U0,...,X_N] (interleaved)",
                 "[x0-X0 ; RK4(X_k,U_k)-X_{k+1}]", "[x0(3); p_0(5) ... p_{N-1}(5)]")


def frenet_bicycle(N=20, T=0.05, M=1, L=3.5, lam=(2.5, 1.75, 2.5, 0.4, 10.0)):
    """Trajectory Tracking/test2.py:20-122 (Frenet kinematic bicycle, RK4 M=1, per-stage p = (yt, phit, kappat, vdes))."""
    return _prob(S.frenet_bicycle(N, T, M, L, lam),
                 "sum_k (l1 (v-vdes)^2 + l2 (y-yt)^2 + l3 (phi-phit)^2 + l4 a^2 + l5 (tan delta - L kappat)^2)/(Nt+1)",
                 "[X0,U0,...,X_N], X = (y, phi, v, delta_prev), U = (d_delta, a)", "[xbar - X0 ; F(X_k,U_k) - X_{k+1}]",
                 "[xbar ; p_0 .. p_{N-1}]")


def frenet_bounds(spec, delta_max=0.384, a_max=2.0, ddelta_max=0.1225):
    """test2.py:31-36,55-59: |delta| <= 0.384 (box on the delta_prev state of stages 1..N), |a| <= 2,
    |d delta| <= 0.1225 (MPCTools' Du bound = box on the control)."""
    lbx, ubx = control_box(spec, (-ddelta_max, -a_max), (ddelta_max, a_max),
                           (-np.inf, -np.inf, -np.inf, -delta_max), (np.inf, np.inf, np.inf, delta_max))
    nz = spec.nx + spec.nu
    lbx[3], ubx[3] = -np.inf, np.inf          # stage 0: delta_prev is data (pinned by the x0 equality)
    return lbx, ubx


def linear_tracking(nx, N, Q, R, T=0.0, R1=None, ntu=0):
    """MPCTools linear trackers (lateral-error bicycle nx=3, dynamic bicycle / cart-pendulum nx=4)."""
    sp = S.linear_tracking(nx, N, Q, R, T, R1, ntu)
    return _prob(sp, "sum_{k<N} sum_i Q_i(x_i-r_i)^2 + R(u-r_u)^2 [+ R1 (u-u_prev)^2]",
                 "[X0,U0,...,X_N] (interleaved; X carries u_prev last for Du models)",
                 "[x0-X0 ; A X_k + B U_k - X_{k+1}]", "[x0; A(row-major); B; (r_k, r_u,k) k<N]")


def control_box(spec, u_lo, u_hi, x_lo=None, x_hi=None):
    """lbx/ubx vectors in the decision-vector layout of `spec` (e.g. multiple_shooting_casadi.py:132-168)."""
    n = spec.n_var
    lb, ub = np.full(n, -np.inf), np.full(n, np.inf)
    nx, nu, N = spec.nx, spec.nu, spec.N
    u_lo, u_hi = np.broadcast_to(u_lo, (nu,)), np.broadcast_to(u_hi, (nu,))
    for k in range(N):
        o = k * nu if spec.single else k * (nx + nu) + nx
        lb[o:o + nu], ub[o:o + nu] = u_lo, u_hi
    if not spec.single and x_lo is not None:
        x_lo, x_hi = np.broadcast_to(x_lo, (nx,)), np.broadcast_to(x_hi, (nx,))
        for k in range(N + 1):
            o = k * (nx + nu)
            lb[o:o + nx], ub[o:o + nx] = x_lo, x_hi
    return lb, ub


def unicycle_bounds(spec, x_box=None):
    """v in [-1,1], omega in [-pi/4,pi/4] (single_shooting_v1.py:39-42); optional x,y box."""
    if x_box is None:
        return control_box(spec, (-S.V_MAX, -S.OMEGA_MAX), (S.V_MAX, S.OMEGA_MAX))
    return control_box(spec, (-S.V_MAX, -S.OMEGA_MAX), (S.V_MAX, S.OMEGA_MAX),
                       (-x_box, -x_box, -np.inf), (x_box, x_box, np.inf))


def cold_start(spec, x_init):
    """X_k = x_init for all k, U = 0 — the repmat(state_init,1,N+1) guess of MS:213."""
    x_init = np.atleast_2d(np.asarray(x_init, dtype=np.float64))
    B = x_init.shape[0]
    w0 = np.zeros((B, spec.n_var))
    if not spec.single:
        nz = spec.nx + spec.nu
        for k in range(spec.N + 1):
            w0[:, k * nz:k * nz + spec.nx] = x_init
    return w0


def c2d(Ac, Bc, dt):
    """Exact zero-order hold, mpc.util.c2d: expm([[Ac,Bc],[0,0]] dt) (Inverted_pendulum/...:24)."""
    from scipy.linalg import expm
    Ac, Bc = np.atleast_2d(Ac).astype(float), np.asarray(Bc, dtype=float).reshape(len(Ac), -1)
    n, m = Ac.shape[0], Bc.shape[1]
    M = np.zeros((n + m, n + m))
    M[:n, :n], M[:n, n:] = Ac, Bc
    E = expm(M * dt)
    return E[:n, :n], E[:n, n:]


def rk4_linear(Ac, Bc, dt):
    """Classic RK4 (one step) of xdot = Ac x + Bc u with u held: the discrete pair (A,B)."""
    Ac, Bc = np.atleast_2d(Ac).astype(float), np.asarray(Bc, dtype=float).reshape(len(Ac), -1)
    n = Ac.shape[0]
    h = dt
    A2, A3, A4 = Ac @ Ac, Ac @ Ac @ Ac, Ac @ Ac @ Ac @ Ac
    A = np.eye(n) + h * Ac + h ** 2 / 2 * A2 + h ** 3 / 6 * A3 + h ** 4 / 24 * A4
    B = (h * np.eye(n) + h ** 2 / 2 * Ac + h ** 3 / 6 * A2 + h ** 4 / 24 * A3) @ Bc
    return A, B


# Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:19-22
PENDULUM_AC = np.array([[0, 0, 0, 0], [1, -10, 0, -20], [0, 9.81, 0, 39.24], [0, 0, 1, 0.0]]).T
PENDULUM_BC = np.array([[0.0], [1.0], [0.0], [2.0]])


def dynamic_bicycle_matrices(v, m=1200.0, a=1.5, b=2.0, Ca=55000.0, Jz=1350.0):
    """Trajectory_tracking_dynamic_model.py:37-43,119-128, formulas exactly as written
    (including the operator precedence of A34 at :120)."""
    A33 = -4 * Ca / (m * v)
    A34 = (2 * Ca * (b - a) / m * v) - v
    A43 = 2 * Ca * ((b - a) / (Jz * v))
    A44 = -2 * Ca * (a ** 2 + b ** 2) / (Jz * v)
    B31 = 2 * Ca / m
    B41 = 2 * Ca * a / Jz
    Ac = np.array([[0, v, 1, 0], [0, 0, 0, 1], [0, 0, A33, A34], [0, 0, A43, A44]], dtype=float)
    Bc = np.array([[0.0], [0.0], [B31], [B41]])
    return Ac, Bc


# ---- lateral-error bicycle (Trajectory Tracking/Phiref.py, Trjectory_tracking_le_LTV.py) ---------------
# (the reference-trajectory tables of these trackers are built on the device: mpc_verde_b200/reference.py)
LATERAL_AR, LATERAL_BR = -23.55, 61.99           # Phiref.py:47-48


def lateral_error_matrices(uref, ar=LATERAL_AR, br=LATERAL_BR):
    """Ac = [[0,u_ref,0],[0,0,1],[0,0,a_r]], Bc = [0,0,b_r]' (Phiref.py:87-90; LTV: u_ref = c[t], :158)."""
    return np.array([[0.0, uref, 0.0], [0.0, 0.0, 1.0], [0.0, 0.0, ar]]), np.array([[0.0], [0.0], [br]])
