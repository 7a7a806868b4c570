#!/usr/bin/env python3
"""Timing + oracle spot-check of the BASELINE.json configurations other than the headline one
(C3 pendulum N=40, C4 tracker closed loop N=20, C5 dynamic bicycle N=50).  Not the driver's bench:
bench.py measures the headline metric; this prints one JSON line per configuration.

    python scripts/bench_configs.py [--scale 1.0] [--configs c3,c4,c5]
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import mpc_verde_b200 as mv  # noqa: E402
from mpc_verde_b200 import problems  # noqa: E402
from mpc_verde_b200 import spec as S  # noqa: E402
from oracle import mpc_oracle as O  # noqa: E402
from tests import common  # noqa: E402

OPTS = {"ipopt": {"max_iter": 2000, "print_level": 0}, "print_time": 0}


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, min(ts)


def c3(scale, layout):
    """pendulum (Inverted_pendulum/...:19-64 dynamics), N=40, T=0.01, RK4, |u|<=200, B=262,144, one solve each"""
    B = int(262144 * scale)
    sp0, lbx, ubx, pglob, _, _ = common.pendulum_setup(N=40, ntu=0, discretisation="rk4")
    solver = mv.nlpsol("solver", "ipopt", {"spec": sp0}, dict(OPTS, layout=layout))
    sp = solver.spec
    x0, p = common.pendulum_batch(sp, pglob, B)
    w0 = problems.cold_start(sp, x0)
    w0d, pd = torch.as_tensor(w0).cuda(), torch.as_tensor(p).cuda()
    sol, ms = timed(lambda: solver(x0=w0d, lbx=lbx, ubx=ubx, p=pd, outputs=("x", "f")))
    st = solver.stats()
    idx = np.random.default_rng(0).choice(B, 128, replace=False)
    ref = O.solve(sp, w0[idx], lbx, ubx, p[idx], nthreads=os.cpu_count() or 1)
    x = sol["x"].cpu().numpy()
    return {"config": "C3 pendulum N=40", "B": B, "ms": ms, "solves_per_s": B / ms * 1e3, "success": st["success"],
            "mean_iters": float(np.mean(st["iter_count"])), "max_dx_vs_oracle": float(np.abs(x[idx] - ref["x"]).max()),
            "sweeps": solver.phase_sweeps() if layout in (0, 3) else None}


def c5(scale, layout):
    """dynamic bicycle (Trajectory_tracking_dynamic_model.py:37-55,119-134), N=50, c2d per scenario, one solve each"""
    B = int(131072 * scale)
    N, dt = 50, 0.05
    rng = np.random.default_rng(7)
    solver = mv.nlpsol("solver", "ipopt", problems.linear_tracking(4, N, Q=(1, 1, 1, 1), R=1.0, T=dt), dict(OPTS, layout=layout))
    sp = solver.spec
    lbx, ubx = problems.control_box(sp, -20.0, 20.0)
    x0 = rng.normal(size=(B, 4)) * np.array([0.2, 0.05, 0.1, 0.05])
    vs = rng.uniform(0.4, 0.8, 64)
    AB = []
    for v in vs:
        Ac, Bc = problems.dynamic_bicycle_matrices(v)
        A, Bd = problems.c2d(Ac, Bc, dt)
        AB.append(np.concatenate([A.ravel(), Bd.ravel()]))
    AB = np.array(AB)[rng.integers(0, 64, B)]
    stage = np.zeros((B, N, 5))
    stage[:, :, 0] = np.linspace(0, 1, N)[None, :] * rng.uniform(0.5, 1.5, (B, 1))
    p = np.concatenate([x0, AB, stage.reshape(B, -1)], 1)
    w0 = problems.cold_start(sp, x0)
    w0d, pd = torch.as_tensor(w0).cuda(), torch.as_tensor(p).cuda()
    sol, ms = timed(lambda: solver(x0=w0d, lbx=lbx, ubx=ubx, p=pd, outputs=("x", "f")))
    st = solver.stats()
    idx = np.random.default_rng(0).choice(B, 64, replace=False)
    ref = O.solve(sp, w0[idx], lbx, ubx, p[idx], nthreads=os.cpu_count() or 1)
    x = sol["x"].cpu().numpy()
    return {"config": "C5 dynamic bicycle N=50", "B": B, "ms": ms, "solves_per_s": B / ms * 1e3, "success": st["success"],
            "mean_iters": float(np.mean(st["iter_count"])), "max_dx_vs_oracle": float(np.abs(x[idx] - ref["x"]).max()),
            "sweeps": solver.phase_sweeps() if layout in (0, 3) else None}


def c4(scale, layout):
    """unicycle tracker (Trajectory_tracking.py:40-112) on lane-change-style references, N=20, closed loop with
    shifted warm start: scenarios x steps closed-loop solves"""
    nsc, nst = int(4096 * scale), 100
    solver = mv.nlpsol("solver", "ipopt", problems.unicycle_tracking(N=20, T=0.05, M=1), dict(OPTS, layout=layout))
    sp = solver.spec
    rng = np.random.default_rng(20264)
    g = common.golden("lane_change.csv")        # x, y, uref  (Trajectory Tracking/lane_change.csv)
    T = nst + sp.N
    lat, spd = rng.uniform(0.5, 1.5, nsc), rng.uniform(0.75, 1.25, nsc)
    xs = g[:T, 0][None, :] * spd[:, None]
    ys = g[:T, 1][None, :] * lat[:, None]
    th = np.arctan2(np.gradient(ys, axis=1), np.gradient(xs, axis=1))
    v = np.hypot(np.gradient(xs, axis=1), np.gradient(ys, axis=1)) / sp.T
    w = np.gradient(th, axis=1) / sp.T
    ptraj = np.stack([xs, ys, th, np.clip(v, -1, 1), np.clip(w, -math.pi / 4, math.pi / 4)], 2)
    x_init = np.stack([xs[:, 0], ys[:, 0] + rng.normal(size=nsc) * 0.05, th[:, 0]], 1)
    lbx, ubx = problems.control_box(sp, (-1, -math.pi / 4), (1, math.pi / 4), (-20, -20, -np.inf), (20, 20, np.inf))
    t0 = time.perf_counter()
    r = solver.closed_loop(x_init, None, ptraj, lbx, ubx, n_steps=nst, warm_mode=S.WARM_SHIFT)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    r = solver.closed_loop(x_init, None, ptraj, lbx, ubx, n_steps=nst, warm_mode=S.WARM_SHIFT)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t1) * 1e3
    ro = O.closed_loop(sp, x_init[:8], None, ptraj[:8], lbx, ubx, nst, S.WARM_SHIFT, 0.0)
    return {"config": "C4 unicycle tracker N=20 closed loop", "scenarios": nsc, "steps": nst, "ms": ms,
            "solves_per_s": nsc * nst / ms * 1e3, "ok": bool(np.all(r["status"] == 0)),
            "mean_iters_per_solve": float(r["iters"].mean() / nst),
            "max_du_vs_oracle": float(np.abs(r["controls"][:8] - ro["controls"]).max()), "first_call_ms": (t1 - t0) * 1e3}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--configs", default="c3,c4,c5")
    ap.add_argument("--layout", type=int, default=0)
    a = ap.parse_args()
    for name in a.configs.split(","):
        print(json.dumps(globals()[name](a.scale, a.layout)), flush=True)
