"""Result sinks next to the hot path: the table layouts the reference scripts dump
(and their overlay scripts read back) and the tracking-error metrics of the lateral-error trackers, from the
arrays `closed_loop` returns.  Plain CSV (the reference writes .xlsx through pandas/openpyxl, which this image
does not have; `Casadi/plot.py` / `leitordados.py` read the same columns).
"""
import numpy as np


def unicycle_table(states, controls, T, n_steps):
    """Rows of `1exemplo.xlsx` / `2exemplo.xlsx` (Casadi/multiple_shooting_casadi.py:316-334): columns
    x, y, theta, v, w, t with q[0] = q[1] = x_init (the scripts stack the initial prediction first), the control
    column shifted one row up and its last row repeated, t = [0, 0, T, 2T, ...].
    states [n_steps+1, 3], controls [n_steps, 2] of ONE scenario as returned by closed_loop."""
    n = int(n_steps)
    q = np.vstack([states[0:1], states[: n]])                       # q[0] = q[1] = x_init, ..., state at iteration n-1
    w = np.vstack([controls[:n], controls[n - 1:n]])                # shifted: row i = control applied at iteration i
    t = np.concatenate([[0.0], np.arange(n) * T])
    return np.column_stack([q, w, t])


def write_csv(path, table, header):
    np.savetxt(path, table, delimiter=",", header=",".join(header), comments="", fmt="%.17g")


def lateral_tracking_errors(x, u, par, a, b, c, Delta):
    """The error log of Trajectory Tracking/Trjectory_tracking_le_LTV.py:173-188, as written: dead-reckoned path
    (xz, yz), mean squared deviations from the stage-0 references and the running path distance `dist`,
    its running mean `mse` and maximum `max`.  x [Nsim+1, 3], u [Nsim], par [4, Nt, Nsim]."""
    Nsim = u.shape[0]
    xz, yz = [], []
    mean = np.zeros(4)
    dist = mse = mx = 0.0
    for t in range(Nsim):
        xz.append(0.0 if t == 0 else xz[t - 1] + c[t] * np.cos(x[t, 1]) * Delta)
        yz.append(x[t, 0])
        mean[0] += (x[t, 0] - par[0, 0, t]) ** 2 / Nsim
        mean[1] += (x[t, 1] - par[1, 0, t]) ** 2 / Nsim
        mean[2] += (x[t, 2] - par[2, 0, t]) ** 2 / Nsim
        mean[3] += (u[t] - par[3, 0, t]) ** 2 / Nsim
        dist += np.hypot(xz[t] - a[t], yz[t] - b[t]) / Nsim
        mse += dist / (t + 1)
        mx = max(mx, abs(dist))
    return {"path": np.array([xz, yz]), "mean_sq": mean, "dist": dist, "mse": mse, "max": mx}
