"""TEST INFRASTRUCTURE — the reference scripts' own nested-loop builders of the per-stage parameter tables and of
the extended lane-change path, restated one scenario at a time with their sequential read-backs, as the pin for the
device pipeline (mpc_verde_b200/reference.py -> csrc/mpcv_ref.cu, which computes every entry independently in
closed form).  Checked against the reference's committed outputs (dados2.csv columns, out.csv) in
tests/test_oracle_golden.py; nothing in the product imports this module."""
import numpy as np

LATERAL_AR, LATERAL_BR = -23.55, 61.99           # Phiref.py:47-48


def lateral_error_par(a, b, Nt, Delta, ar=LATERAL_AR, br=LATERAL_BR):
    """Reference builder of the lateral-error trackers, exactly as written (Phiref.py:124-155): per MPC step t
    and horizon stage k the parameters (y_ref, phi_ref, r_ref, delta_ref) from finite differences of the path
    (a, b).  Quirks kept: `par[1, k, t-1]` reads the previous step's slice — at t = 0 that is the still-zero
    LAST slice — and `par[1, k-1, t-1]` wraps to the last stage at k = 0.  Returns par [4, Nt, Nsim]."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    Nsim = a.size
    par = np.zeros((4, Nt, Nsim))
    p = np.zeros((Nt, 4))
    for t in range(Nsim):
        for k in range(Nt):
            if t + k > Nsim - 1:
                p[k, 0] = b[Nsim - 1]
                p[k, 1] = np.arctan2(b[Nsim - 1] - b[Nsim - 2], a[Nsim - 1] - a[Nsim - 2])
            elif t + k == 0:
                p[k, 0] = b[k + t]
                p[k, 1] = 0.0
            else:
                p[k, 0] = b[k + t]
                p[k, 1] = np.arctan2(b[k + t] - b[k + t - 1], a[k + t] - a[k + t - 1])
            if t + k < 2:
                plus = np.arctan2(b[k + 1 + t] - b[k + t], a[k + 1 + t] - a[k + t])
                plus2 = np.arctan2(b[k + 2 + t] - b[k + 1 + t], a[k + 2 + t] - a[k + 1 + t])
                p[k, 2] = (plus - p[k, 1]) / Delta
                p[k, 3] = (((plus2 - 2 * plus + p[k, 1]) / Delta ** 2) - ar * p[k, 2]) / br
            elif t + k > Nsim - 3:
                p[k, 2] = (p[k, 1] - par[1, k, t - 1]) / Delta
                p[k, 3] = (((p[k, 1] - 2 * par[1, k, t - 1] + par[1, k - 1, t - 1]) / Delta ** 2) - ar * p[k, 2]) / br
            else:
                plus = np.arctan2(b[k + 1 + t] - b[k + t], a[k + 1 + t] - a[k + t])
                p[k, 2] = (plus - par[1, k, t - 1]) / (2 * Delta)
                p[k, 3] = (((plus - 2 * p[k, 1] + par[1, k, t - 1]) / Delta ** 2) - ar * p[k, 2]) / br
            par[:, k, t] = p[k, :]
    return par


def lane_change_extended(a, b, c, v=0.6, dt=0.05):
    """Reference-path generator Trajectory Tracking/lane_change.py:5-79, as written: the 500-sample lane change
    (a, b, c) = (x, y, uref) followed by a half turn, a straight, two quarter-radius half turns, a straight back
    to x = 0 and a closing half turn; `uref` = v on the appended part.  (`np.linspace(..., num=float)` of the
    script's numpy era truncates: int(k).)  Returns (x_t, y_t, c2) — the columns of `out.csv`."""
    a, b, c = (np.asarray(q, dtype=np.float64) for q in (a, b, c))
    k = 500
    w = np.pi / (k * dt)
    r = v / w
    t = np.linspace(1.5 * np.pi, 2.5 * np.pi, k)
    x_2, y_2 = a[499] + r * np.cos(t), b[499] + r + r * np.sin(t)
    ds = 10
    k = ds / (v * dt)
    t = np.linspace(0, ds, int(k))
    x_3, y_3 = x_2[-1] - t * v, y_2[-1] + np.zeros(int(k))
    w = v / (r / 2)
    k = np.pi / (w * dt)
    t = np.linspace(np.pi / 2, 1.5 * np.pi, int(k))
    x_4, y_4 = x_3[-1] + (r / 2) * np.cos(t), y_3[-1] - r / 2 + (r / 2) * np.sin(t)
    t = np.linspace(np.pi / 2, -np.pi / 2, int(k))
    x_5, y_5 = x_4[-1] + (r / 2) * np.cos(t), y_4[-1] - 0.5 * r + (r / 2) * np.sin(t)
    d = x_5[-1]
    k = d / (v * dt)
    t = np.linspace(0, k * dt, int(k))
    x_6, y_6 = d - v * t, y_5[-1] + np.zeros(int(k))
    r = y_6[-1] / 2
    w = v / r
    k = np.pi / (w * dt)
    t = np.linspace(np.pi / 2, 1.5 * np.pi, int(k))
    x_7, y_7 = x_6[-1] + r * np.cos(t), y_6[-1] - r + r * np.sin(t)
    x_t = np.hstack((a, x_2[1:], x_3[1:], x_4[1:], x_5[1:], x_6[1:], x_7[1:]))
    y_t = np.hstack((b, y_2[1:], y_3[1:], y_4[1:], y_5[1:], y_6[1:], y_7[1:]))
    c2 = np.zeros(x_t.size)
    c2[0:500] = c
    c2[500:] = v
    return x_t, y_t, c2


def circle_reference_par(Nt, Nsim, Delta):
    """Trajectory Tracking/Trajectory_tracking.py:84-97: per-stage p = (x, y, theta, v, omega) of the unit circle
    x = cos 0.1 t, y = sin 0.1 t, theta = pi/2 + 0.1 t, v_ref = 1, omega_ref = 1.  Returns par [5, Nt, Nsim]."""
    par = np.zeros((5, Nt, Nsim))
    for t in range(Nsim):
        for k in range(Nt):
            tt = (t + k) * Delta
            par[:, k, t] = (np.cos(0.1 * tt), np.sin(0.1 * tt), np.pi / 2 + 0.1 * tt, 1.0, 1.0)
    return par


def frenet_reference_par(xtraj, ytraj, vdes, Nt, Delta, t):
    """Per-stage parameters of Trajectory Tracking/test2.py:79-100 at MPC step t, as written: p[k] =
    (y_ref, phi_ref, p2, p3) with p2 = vdes and p3 = ||(xdd, ydd)|| from central differences — which the cost
    and the model then unpack as (kappat, vdes) = (p[2], p[3]): swapped, and kept that way.  Returns p [Nt, 4]."""
    xtraj, ytraj, vdes = (np.asarray(q, dtype=np.float64) for q in (xtraj, ytraj, vdes))
    Nsim = 500                                   # test2.py:61 (the builder indexes against it)
    p = np.zeros((Nt, 4))
    for k in range(Nt):
        if t + k > Nsim - 1:
            p[k, 0] = ytraj[Nsim - 1]
            p[k, 1] = np.arctan2(ytraj[Nsim - 1] - ytraj[Nsim - 2], xtraj[Nsim - 1] - xtraj[Nsim - 2])
        elif t + k == 0:
            p[k, 0] = ytraj[k + t]
            p[k, 1] = 0.0
        else:
            p[k, 0] = ytraj[k + t]
            p[k, 1] = np.arctan2(ytraj[k + t] - ytraj[k + t - 1], xtraj[k + t] - xtraj[k + t - 1])
        if t + k < 2:
            p[k, 3] = 1.0
            p[k, 2] = vdes[t + k]
        elif t + k > Nsim - 2:
            p[k, 3] = p[k - 1, 3]
            p[k, 2] = vdes[Nsim - 1]
        else:
            ddx = (xtraj[k + t - 1] - 2 * xtraj[k + t] + xtraj[k + t + 1]) / Delta ** 2
            ddy = (ytraj[k + t - 1] - 2 * ytraj[k + t] + ytraj[k + t + 1]) / Delta ** 2
            p[k, 3] = np.hypot(ddx, ddy)
            p[k, 2] = vdes[t + k]
    return p
