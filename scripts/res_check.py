#!/usr/bin/env python3
"""Development check of the CTA-resident layout (mpcv_resident.cuh) on a GPU box: results against the CPU oracle and
against the phase pipeline, then the C2 batch timed in both layouts.  One JSON line per check.

    python scripts/res_check.py [--batch 65536] [--reps 3] [--skip-parity]
"""
import argparse
import json
import math
import os
import sys

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
NCPU = os.cpu_count() or 1


def solve_dev(solver, w0, lbx, ubx, p):
    sol = solver(x0=None if w0 is None else torch.as_tensor(w0).cuda(), lbx=lbx, ubx=ubx, p=torch.as_tensor(p).cuda())
    torch.cuda.synchronize()
    st = solver.stats()
    return {k: v.cpu().numpy() for k, v in sol.items()}, st


def parity(name, prob, w0, lbx, ubx, p, nref=256):
    out = {"check": name}
    res = {}
    for lay, tag in ((S.LAYOUT_RESIDENT, "res"), (S.LAYOUT_PHASED, "ph")):
        solver = mv.nlpsol("solver", "ipopt", prob, dict(OPTS, layout=lay))
        res[tag] = solve_dev(solver, w0, lbx, ubx, p)
    sp = prob["spec"]
    idx = np.arange(min(nref, p.shape[0]))
    ref = O.solve(sp, None if w0 is None else w0[idx], lbx, ubx, p[idx], nthreads=NCPU)
    (a, sa), (b, sb) = res["res"], res["ph"]
    out["status_res"] = {int(k): int(v) for k, v in zip(*np.unique(sa["status_code"], return_counts=True))}
    out["status_ph"] = {int(k): int(v) for k, v in zip(*np.unique(sb["status_code"], return_counts=True))}
    out["status_eq_oracle"] = bool(np.array_equal(sa["status_code"][idx], ref["status"]))
    ok = (sa["status_code"][idx] == 0) & (ref["status"] == 0)
    same = ok & (sa["iter_count"][idx] == ref["iters"])
    out["iters_eq_oracle"] = float(np.mean(sa["iter_count"][idx] == ref["iters"]))
    out["iters_eq_phased"] = float(np.mean(sa["iter_count"] == sb["iter_count"]))
    out["max_dx_oracle_sameiters"] = float(np.abs(a["x"][idx][same] - ref["x"][same]).max()) if same.any() else None
    out["max_dx_oracle_ok"] = float(np.abs(a["x"][idx][ok] - ref["x"][ok]).max()) if ok.any() else None
    out["max_df_rel_oracle"] = float((np.abs(a["f"][idx][ok] - ref["f"][ok]) / np.abs(ref["f"][ok]).max()).max()) if ok.any() else None
    both = (sa["status_code"] == 0) & (sb["status_code"] == 0) & (sa["iter_count"] == sb["iter_count"])
    out["max_dx_phased"] = float(np.abs(a["x"][both] - b["x"][both]).max())
    out["max_dlamg_phased"] = float(np.abs(a["lam_g"][both] - b["lam_g"][both]).max())
    out["mean_iters"] = float(np.mean(sa["iter_count"]))
    print(json.dumps(out), flush=True)


def timing(B, reps, layouts):
    prob = problems.unicycle_multiple_shooting()
    sp = prob["spec"]
    x0s, p = common.unicycle_batch(B)
    lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
    w0 = problems.cold_start(sp, x0s)
    w0d, pd = torch.as_tensor(w0).cuda(), torch.as_tensor(p).cuda()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    for lay in layouts:
        solver = mv.nlpsol("solver", "ipopt", prob, dict(OPTS, layout=lay))
        for _ in range(2):
            solver(x0=w0d, lbx=lbx, ubx=ubx, p=pd, outputs=("x", "f"))
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            solver(x0=w0d, lbx=lbx, ubx=ubx, p=pd, outputs=("x", "f"))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        st = solver.stats()
        print(json.dumps({"timing": "C2", "layout": lay, "B": B, "ms": ts, "best_ms": min(ts),
                          "solves_per_s": B / min(ts) * 1e3, "success": st["success"],
                          "mean_iters": float(np.mean(st["iter_count"])), "max_iters": int(np.max(st["iter_count"]))}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--layouts", default="4,3")
    a = ap.parse_args()
    if not a.skip_parity:
        prob = problems.unicycle_multiple_shooting()
        sp = prob["spec"]
        x0s, p = common.unicycle_batch(4096)
        lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
        parity("C2 cold start", prob, problems.cold_start(sp, x0s), lbx, ubx, p)
        x0s, p = common.unicycle_batch(512, seed=77)
        lbx, ubx = problems.unicycle_bounds(sp)
        parity("C2 zeros guess (slow path, failures)", prob, None, lbx, ubx, p)
        # tracker N = 20 with per-stage references
        prob = problems.unicycle_tracking(N=20, T=0.05, M=1)
        sp = prob["spec"]
        rng = np.random.default_rng(5)
        B = 1024
        tt = np.arange(sp.N)[None, :] * sp.T + rng.uniform(0, 50, (B, 1))
        stage = np.stack([np.cos(0.1 * tt), np.sin(0.1 * tt), math.pi / 2 + 0.1 * tt, np.ones_like(tt) * 0.1, np.ones_like(tt) * 0.1], 2)
        x0 = stage[:, 0, :3] + rng.normal(size=(B, 3)) * 0.1
        p = np.concatenate([x0, stage.reshape(B, -1)], 1)
        lbx, ubx = problems.control_box(sp, (-1, -math.pi / 4), (1, math.pi / 4), (-20, -2, -np.inf), (20, 2, np.inf))
        parity("tracker N=20", prob, problems.cold_start(sp, x0), lbx, ubx, p)
        # linear with move blocking (pendulum N = 10)
        sp, lbx, ubx, pglob, _, _ = common.pendulum_setup(N=10, ntu=3)
        x0, p = common.pendulum_batch(sp, pglob, 512)
        parity("pendulum N=10 ntu=3", {"spec": sp}, problems.cold_start(sp, x0), lbx, ubx, p)
    timing(a.batch, a.reps, [int(x) for x in a.layouts.split(",")])
