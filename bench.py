#!/usr/bin/env python3
"""bench.py — batched MPC solves/sec (BASELINE.json metric) on every BASELINE.json configuration.

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path, headline config C2
    python bench.py --config c1|c2|c3|c4|c4f|c5 ...                # any configuration as the headline line
    python bench.py --impl reference --gpus N --steps K ...        # CPU arm (oracle port of CasADi/IPOPT)

Configurations (BASELINE.json `configs`, synthetic inputs of SURVEY.md 8d; per-GPU sizes, weak scaling):
  c1   unicycle single shooting, Euler, N=10, the script's own 84-step closed loop, ONE problem (latency pair)
       Casadi/single_shooting_v1.py:29-214
  c2   unicycle multiple shooting, N=10, RK4 M=4 + quadrature, 65,536 random initial states, cold start   [headline]
       Casadi/multiple_shooting_casadi.py:29-298
  c3   cart-pendulum (linear, Du cost), N=40, T=0.01, RK4-of-linear, |u|<=200, 262,144 problems
       Inverted_pendulum/inverted_pendulum_single_shooting_mpctools.py:10-64
  c4   unicycle tracker N=20 on lane-change references, warm-started closed loop, 1,024 scenarios x 128 steps
       Trajectory Tracking/Trajectory_tracking.py:15-118, lane_change.csv
  c4f  Frenet kinematic bicycle N=20 (nonlinear), same closed loop, 512 scenarios x 64 steps
       Trajectory Tracking/test2.py:20-122
  c5   dynamic bicycle N=50, LTV in v_ref[t] (c2d per scenario and step on the device), warm-started closed loop,
       131,072 scenarios x 2 steps      Trajectory Tracking/Trajectory_tracking_dynamic_model.py:17-141
A "step" is one pass of the hot path over the per-GPU batch (closed-loop configs: the whole n_steps loop of every
scenario).  Every rank owns its own problems (sharded by problem index, no collective on the hot path); after each
step NCCL gathers the results and all-reduces a small statistics vector.

The default run prints ONE JSON line for C2 and attaches short runs of the other configurations under
`other_configs` (--no-others skips them), so that the driver's record holds every configuration.
Keys beyond the contract:
  roofline     — FP64 vector-pipe roofline (this path is neither HBM- nor tensor-bound: ~360 FLOP/B; tensor cores
                 deliberately unused); algorithmic FLOPs = measured IPM iterations x W_iter(config), SURVEY 8d,
                 capped by the FP64 operations the kernels actually execute (ncu, profiles/) where that is known
  cpu_baseline — the oracle timed on the box's host cores on a bounded sample
  p50_solve_us — median batch-residency latency per problem (device %globaltimer); lone_problem_ms — one problem alone
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mpc_verde_b200 import problems  # noqa: E402
from mpc_verde_b200 import spec as S  # noqa: E402

SEED = 20260          # + config index
METRIC = "batched MPC solves/sec (RK4 NLP to IPOPT tol)"
OPTS = {"ipopt": {"max_iter": 2000, "print_level": 0, "acceptable_tol": 1e-8, "acceptable_obj_change_tol": 1e-6},
        "print_time": 0}
GOLDEN = os.path.join(ROOT, "tests", "golden")


def lane_change_csv():
    """Trajectory Tracking/lane_change.csv (x, y, uref; committed input fixture)."""
    return np.loadtxt(os.path.join(GOLDEN, "lane_change.csv"), delimiter=",", skiprows=1)


# =====================================================================================================================
# workloads: one class per BASELINE.json configuration
# =====================================================================================================================
class Workload:
    key = ""
    index = 0
    workload = ""
    flop_per_iter = 0.0           # ALGORITHMIC FLOPs per IPM iteration
    closed_loop = False
    # batches in flight when --inflight is left at 0: 4 where stragglers / latency-bound closed loops leave the GPU
    # idle (C2, C4), 1 for the HBM-bound linear configurations (C3: 138 ms one at a time, 143 with four in flight)
    default_inflight = 4

    def __init__(self, batch=None, rank=0):
        self.rank = rank
        self.B = int(batch or self.default_batch)
        self.seed = SEED + self.index + rank             # every rank owns different problems (C2: 20261 + rank, as in round 1)

    # -- host-side description (no GPU needed: the CPU arm uses it too) --
    def problem(self):
        raise NotImplementedError

    # -- device side --
    def setup(self, mv, dev, layout=S.LAYOUT_AUTO, pipes=None):
        import torch
        self.mv, self.dev, self.torch, self.layout, self.pipes = mv, dev, torch, layout, pipes
        extra = {"layout": layout}
        if pipes:
            extra["pipes"] = pipes
        self.solver = mv.nlpsol("solver", "ipopt", self.problem(), dict(OPTS, **extra))
        self.spec = self.solver.spec
        self._setup_inputs()

    def solves_per_step(self):
        return self.B

    def algorithmic_bytes_per_solve(self):
        sp = self.spec
        return 8 * (sp.n_p + 2 * sp.n_var + 1) + 8


class C2(Workload):
    key, index, default_batch = "c2", 1, 65536
    workload = ("C2 unicycle multiple shooting N=10 T=0.2 RK4(M=4)+quadrature, cold start, "
                "x,y in [-20,20], v in [-1,1], w in [-pi/4,pi/4]")
    #   N*(C_HJ + C_ric + C_bar) + n_ls*N*C_val + C_norm = 10*(1868+242+50) + 10*283 + 300
    flop_per_iter = 24730.0

    def problem(self):
        return problems.unicycle_multiple_shooting()

    def host_inputs(self, n=None):
        sp = self.problem()["spec"]
        n = self.B if n is None else n
        rng = np.random.default_rng(self.seed)
        x0s = np.stack([rng.uniform(-2, 12, self.B), rng.uniform(-2, 12, self.B), rng.uniform(-math.pi, math.pi, self.B)], 1)[:n]
        p = np.concatenate([x0s, np.tile([10.0, 10.0, 0.0], (n, 1))], 1)
        lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
        return problems.cold_start(sp, x0s), lbx, ubx, p

    def _setup_inputs(self):
        t = self.torch
        self.w0_h, self.lbx, self.ubx, self.p_h = self.host_inputs()
        self.w0, self.p = t.as_tensor(self.w0_h).to(self.dev), t.as_tensor(self.p_h).to(self.dev)
        self.lb, self.ub = t.as_tensor(self.lbx).to(self.dev), t.as_tensor(self.ubx).to(self.dev)
        self.w0_pin, self.p_pin = t.as_tensor(self.w0_h).pin_memory(), t.as_tensor(self.p_h).pin_memory()

    def step(self):
        sol = self.solver(x0=self.w0, lbx=self.lb, ubx=self.ub, p=self.p, outputs=("x", "f"))
        self.status, self.iters = self.solver._last
        return [sol["x"], sol["f"]]

    def host_step(self):
        """public API with HOST buffers: H2D of the step's inputs from page-locked memory, solve, D2H of x, f, status, iters"""
        sol = self.solver(x0=self.w0_pin, lbx=self.lbx, ubx=self.ubx, lbg=0, ubg=0, p=self.p_pin, outputs=("x", "f"))
        t = self.torch
        return [t.as_tensor(sol["x"]), t.as_tensor(sol["f"])]

    def io_bytes(self):
        sp = self.spec
        return 8 * (self.w0_h.size + self.p_h.size + 2 * sp.n_var), 8 * (self.B * sp.n_var + self.B) + 8 * self.B

    def cpu_sample(self, n, threads):
        from oracle import mpc_oracle
        sp = self.problem()["spec"]
        S.ipopt_defaults(sp, OPTS)
        w0, lbx, ubx, p = self.host_inputs(n)
        mpc_oracle.solve(sp, w0[:threads], lbx, ubx, p[:threads], nthreads=threads)    # warm the library
        t0 = time.perf_counter()
        r = mpc_oracle.solve(sp, w0, lbx, ubx, p, nthreads=threads)
        dt = time.perf_counter() - t0
        return n, dt, r

    def check(self, outs, n=64):
        """spot check against the oracle: max |dx| over the first n problems"""
        _, _, r = self.cpu_sample(n, os.cpu_count() or 1)
        return float(np.abs(outs[0][:n].cpu().numpy() - r["x"]).max())


class C3(C2):
    default_inflight = 1
    key, index, default_batch = "c3", 2, 262144
    workload = ("C3 cart-pendulum (linear, Du cost R1=1e-4, Q=(1.44,0,1,0)) N=40 T=0.01 RK4-of-linear, |u|<=200, "
                "set-point x=10, cold start")
    flop_per_iter = 16900.0

    def problem(self):
        return problems.linear_tracking(4, 40, Q=(1.2 ** 2, 0.0, 1.0, 0.0), R=0.0, T=0.01, R1=0.01 ** 2, ntu=0)

    def host_inputs(self, n=None):
        sp = self.problem()["spec"]
        n = self.B if n is None else n
        A, Bd = problems.rk4_linear(problems.PENDULUM_AC, problems.PENDULUM_BC, 0.01)
        pglob = np.concatenate([A.ravel(), Bd.ravel()])
        rng = np.random.default_rng(self.seed)
        x0 = np.stack([rng.uniform(-1, 1, self.B), rng.uniform(-0.5, 0.5, self.B), rng.uniform(-0.2, 0.2, self.B),
                       rng.uniform(-0.5, 0.5, self.B), np.zeros(self.B)], 1)[:n]
        stage = np.tile([10.0, 0.0, 0.0, 0.0, 0.0], sp.N)
        p = np.concatenate([x0, np.tile(pglob, (n, 1)), np.tile(stage, (n, 1))], 1)
        lbx, ubx = problems.control_box(sp, -200.0, 200.0)
        return problems.cold_start(sp, x0), lbx, ubx, p


class LoopWorkload(Workload):
    """warm-started closed loops: a step = the whole n_steps loop of every scenario (B x n_steps solves)"""
    closed_loop = True

    def solves_per_step(self):
        return self.B * self.n_steps

    def step(self):
        r = self._loop(self.dev_in)
        self.status, self.iters = r["status"], r["iters"]
        return [r["controls"], r["states"]]

    def host_step(self):
        """public API from HOST inputs: scenario descriptions host -> device, reference tables built on the device,
        the closed loop, histories device -> host"""
        if not hasattr(self, "host_pin"):
            self.host_pin = {k: self.torch.as_tensor(v).pin_memory() for k, v in self.host_in.items()}
        dev_in = {k: v.to(self.dev, non_blocking=True) for k, v in self.host_pin.items()}
        r = self._loop(dev_in)
        return [r["controls"].cpu(), r["states"].cpu()]

    def io_bytes(self):
        sp = self.spec
        h2d = sum(v.nbytes for v in self.host_in.values() if isinstance(v, np.ndarray))
        d2h = 8 * self.B * (self.n_steps * sp.nu + (self.n_steps + 1) * sp.nx)
        return h2d, d2h

    def algorithmic_bytes_per_solve(self):
        sp = self.spec
        return 8 * (sp.nx + sp.npg + sp.nps + sp.nu + sp.nx)     # x0, this step's model / window entry in; u0, x+ out

    def check(self, outs, n=8):
        from oracle import mpc_oracle
        n = min(n, self.B)
        a = self._oracle_args(n)
        ro = mpc_oracle.closed_loop(self.spec, *a)
        return float(np.abs(outs[0][:n].cpu().numpy() - ro["controls"]).max())


class C4(LoopWorkload):
    key, index, default_batch, n_steps = "c4", 3, 1024, 128
    workload = ("C4 unicycle tracker N=20 T=0.05 RK4(M=1), per-stage (x,y,theta,v,w) references cut on the device from "
                "lane_change.csv scaled per scenario (speed U[0.75,1.25], lateral U[0.5,1.5]), warm-started closed loop")
    flop_per_iter = 9700.0

    def problem(self):
        return problems.unicycle_tracking(N=20, T=0.05, M=1)

    def _setup_inputs(self):
        sp = self.spec
        g = lane_change_csv()
        T = self.n_steps + sp.N
        rng = np.random.default_rng(self.seed)
        scale = np.stack([rng.uniform(0.75, 1.25, self.B), rng.uniform(0.5, 1.5, self.B)], 1)
        x_init = np.stack([g[0, 0] * scale[:, 0], g[0, 1] * scale[:, 1] + rng.normal(size=self.B) * 0.05, np.zeros(self.B)], 1)
        self.lbx, self.ubx = problems.control_box(sp, (-1, -math.pi / 4), (1, math.pi / 4), (-20, -20, -np.inf), (20, 20, np.inf))
        self.host_in = {"x": np.ascontiguousarray(g[:T, 0]), "y": np.ascontiguousarray(g[:T, 1]), "scale": scale, "x_init": x_init}
        self.dev_in = {k: self.torch.as_tensor(v).to(self.dev) for k, v in self.host_in.items()}

    def _refs(self, d):
        from mpc_verde_b200 import reference as R
        return R.unicycle_path_reference(d["x"], d["y"], self.spec.T, scale=d["scale"])

    def _loop(self, d):
        return self.solver.closed_loop(d["x_init"], None, self._refs(d), self.lbx, self.ubx, n_steps=self.n_steps,
                                       warm_mode=S.WARM_SHIFT)

    def _oracle_args(self, n):
        ptraj = self._refs(self.dev_in)[:n].cpu().numpy()
        return self.host_in["x_init"][:n], None, ptraj, self.lbx, self.ubx, self.n_steps, S.WARM_SHIFT, 0.0


class C4F(LoopWorkload):
    key, index, default_batch, n_steps = "c4f", 3, 512, 64
    workload = ("C4 (Frenet) kinematic bicycle of test2.py, N=20 T=0.05 RK4(M=1), |delta|<=0.384 |a|<=2 |d delta|<=0.1225, "
                "per-stage (y,phi,p2,p3) tables built on the device from lane_change.csv scaled per scenario, "
                "warm-started closed loop")
    flop_per_iter = 58700.0

    def problem(self):
        return problems.frenet_bicycle(N=20, T=0.05, M=1)

    def _setup_inputs(self):
        sp = self.spec
        g = lane_change_csv()
        rng = np.random.default_rng(self.seed + 7)
        scale = np.stack([rng.uniform(0.9, 1.1, self.B), rng.uniform(0.5, 1.2, self.B)], 1)
        x_init = np.stack([rng.normal(size=self.B) * 0.02, np.zeros(self.B), g[0, 2] * np.ones(self.B), np.zeros(self.B)], 1)
        self.lbx, self.ubx = problems.frenet_bounds(sp)
        self.host_in = {"x": np.ascontiguousarray(g[:, 0]), "y": np.ascontiguousarray(g[:, 1]),
                        "v": np.ascontiguousarray(g[:, 2]), "scale": scale, "x_init": x_init}
        self.dev_in = {k: self.torch.as_tensor(v).to(self.dev) for k, v in self.host_in.items()}

    def _refs(self, d):
        from mpc_verde_b200 import reference as R
        return R.frenet_windows(d["x"], d["y"], d["v"], self.spec.N, self.spec.T, self.n_steps, scale=d["scale"])

    def _loop(self, d):
        return self.solver.closed_loop(d["x_init"], None, self._refs(d), self.lbx, self.ubx, n_steps=self.n_steps,
                                       warm_mode=S.WARM_SHIFT, windows=True)

    def check(self, outs, n=4):
        # the oracle's closed loop reads sliding windows only: check the first solve of each scenario instead
        from oracle import mpc_oracle
        n = min(n, self.B)
        sp = self.spec
        win = self._refs(self.dev_in)[:n, 0].cpu().numpy().reshape(n, -1)
        x0 = self.host_in["x_init"][:n]
        p = np.concatenate([x0, win], 1)
        r = mpc_oracle.solve(sp, problems.cold_start(sp, x0), self.lbx, self.ubx, p)
        return float(np.abs(outs[0][:n, 0].cpu().numpy() - r["x"][:, sp.nx:sp.nx + sp.nu]).max())


class C5(LoopWorkload):
    default_inflight = 1
    key, index, default_batch, n_steps = "c5", 4, 131072, 2
    workload = ("C5 dynamic bicycle (m=1200,a=1.5,b=2,Ca=55000,Jz=1350; A34 as written) N=50 T=0.05, LTV in v_ref[t]~U[0.4,0.8] "
                "with exact ZOH per scenario and step on the device, Q=I R=1, |delta|<=20, (y,phi) references from "
                "lane_change.csv scaled per scenario, warm-started closed loop")
    flop_per_iter = 21200.0

    def problem(self):
        return problems.linear_tracking(4, 50, Q=(1, 1, 1, 1), R=1.0, T=0.05)

    def _setup_inputs(self):
        sp = self.spec
        g = lane_change_csv()
        T = self.n_steps + sp.N
        rng = np.random.default_rng(self.seed)
        lat = rng.uniform(0.5, 1.5, self.B)
        # start mid-manoeuvre so that the references move: samples 137.. of the lane change
        o = 137
        yref = g[o:o + T, 1][None, :] * lat[:, None]
        phi = np.arctan2(np.gradient(yref, axis=1), np.gradient(g[o:o + T, 0])[None, :])
        ptraj = np.zeros((self.B, T, 5))
        ptraj[:, :, 0], ptraj[:, :, 1] = yref, phi
        x_init = np.stack([yref[:, 0] + rng.normal(size=self.B) * 0.05, phi[:, 0], np.zeros(self.B), np.zeros(self.B)], 1)
        v = rng.uniform(0.4, 0.8, (self.B, self.n_steps))
        self.lbx, self.ubx = problems.control_box(sp, -20.0, 20.0)
        self.host_in = {"x_init": x_init, "ptraj": ptraj, "v": v}
        self.dev_in = {k: self.torch.as_tensor(v_).to(self.dev) for k, v_ in self.host_in.items()}

    def _loop(self, d):
        from mpc_verde_b200 import reference as R
        pgt = R.ltv_dynamic_bicycle(d["v"], self.spec.T, self.n_steps)
        return self.solver.closed_loop(d["x_init"], None, d["ptraj"], self.lbx, self.ubx, n_steps=self.n_steps,
                                       warm_mode=S.WARM_SHIFT, pglob_traj=pgt)

    def _first_solves(self, n, threads=1):
        # first solve of every scenario's loop on the oracle (its closed loop takes a constant model only)
        from oracle import mpc_oracle
        n = min(n, self.B)
        sp = self.spec
        x0, v0 = self.host_in["x_init"][:n], self.host_in["v"][:n, 0]
        AB = []
        for vv in v0:
            A, Bd = problems.c2d(*problems.dynamic_bicycle_matrices(vv), sp.T)
            AB.append(np.concatenate([A.ravel(), Bd.ravel()]))
        p = np.concatenate([x0, np.array(AB), self.host_in["ptraj"][:n, :sp.N].reshape(n, -1)], 1)
        w0 = problems.cold_start(sp, x0)
        t0 = time.perf_counter()
        r = mpc_oracle.solve(sp, w0, self.lbx, self.ubx, p, nthreads=threads)
        return n, time.perf_counter() - t0, r

    def cpu_first_solves(self, n, threads):
        """CPU arm of this configuration: the oracle on the first (cold) solve of n scenarios, all host threads"""
        self._first_solves(threads, threads)
        return self._first_solves(n, threads)

    def check(self, outs, n=8):
        n, _, r = self._first_solves(n)
        return float(np.abs(outs[0][:n, 0, 0].cpu().numpy() - r["x"][:, self.spec.nx]).max())


class C1(LoopWorkload):
    default_inflight = 1
    key, index, default_batch, n_steps = "c1", 0, 1, 100
    workload = ("C1 unicycle single shooting, Euler, N=10 T=0.2, (0,0,0)->(10,10,0), the script's own closed loop "
                "(84 MPC steps, its scrambled warm start), ONE problem: latency pair GPU / CPU")
    flop_per_iter = 7000.0

    def problem(self):
        return problems.unicycle_single_shooting_euler()

    def solves_per_step(self):
        return self.B * 84

    def _setup_inputs(self):
        sp = self.spec
        self.lbx, self.ubx = problems.unicycle_bounds(sp)
        self.host_in = {"x_init": np.zeros((self.B, 3)), "target": np.tile([10.0, 10.0, 0.0], (self.B, 1))}
        self.dev_in = {k: self.torch.as_tensor(v).to(self.dev) for k, v in self.host_in.items()}

    def _loop(self, d):
        return self.solver.closed_loop(d["x_init"], d["target"], None, self.lbx, self.ubx, n_steps=self.n_steps,
                                       warm_mode=S.WARM_REFERENCE, stop_radius=0.1)

    def _oracle_args(self, n):
        return self.host_in["x_init"][:n], self.host_in["target"][:n], None, self.lbx, self.ubx, self.n_steps, S.WARM_REFERENCE, 0.1

    def cpu_loop_ms_per_solve(self):
        from oracle import mpc_oracle
        sp = self.problem()["spec"]
        S.ipopt_defaults(sp, OPTS)
        lbx, ubx = problems.unicycle_bounds(sp)
        args = (np.zeros((1, 3)), np.array([[10.0, 10.0, 0.0]]), None, lbx, ubx, 100, S.WARM_REFERENCE, 0.1)
        mpc_oracle.closed_loop(sp, *args)
        t0 = time.perf_counter()
        r = mpc_oracle.closed_loop(sp, *args)
        dt = time.perf_counter() - t0
        return dt * 1e3 / int(r["steps"][0]), int(r["steps"][0])


CONFIGS = {c.key: c for c in (C1, C2, C3, C4, C4F, C5)}


# =====================================================================================================================
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region: ONE background `nvidia-smi -lms 50` process
    (the profiling recipe's clocks line), started before the timed region and stopped after it."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows, self.stop_flag = index, None, [], False

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return
        time.sleep(0.25)                      # at least one sample after short regions
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 7:
                self.rows.append(f)
        self.proc = None

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, nm in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"), (6, "sw_power_cap")):
            if any(r[i].lower().startswith("active") for r in self.rows):
                reasons.append(nm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def config_block(wl, world, extra=None):
    c = {"workload": wl.workload, "config": wl.key, "batch_per_gpu": wl.B, "global_batch": world * wl.B, "seed": SEED + wl.index,
         "solves_per_step_per_gpu": wl.solves_per_step()}
    if wl.closed_loop:
        c["closed_loop_steps"] = wl.n_steps
    if extra:
        c.update(extra)
    return c


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  CasADi/IPOPT cannot
    be installed here (no network; not in the wheelhouse), so this arm times the oracle port with
    every host thread; each step is a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wl = CONFIGS[args.config](batch=args.batch)
    if wl.key == "c1":
        ms, steps = wl.cpu_loop_ms_per_solve()
        value, sample, dt, iters = 1e3 / ms, "%d-step closed loop of one problem, single thread" % steps, ms * steps * 1e-3, None
        threads = 1
    elif wl.closed_loop:
        raise SystemExit("--impl reference supports c1, c2, c3 (the closed-loop configs c4/c4f/c5 report their CPU "
                         "baseline inside the b200 line)")
    else:
        n = (256 if wl.key == "c2" else 64) * threads
        for _ in range(max(args.warmup, 1)):
            wl.cpu_sample(threads * 4, threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _, _, r = wl.cpu_sample(n, threads)
        dt = (time.perf_counter() - t0) / args.steps
        value, sample, iters = n / dt, "%d problems per step" % n, float(r["iters"].mean())
    ref_f = 1 if wl.key == "c1" else (args.inflight if args.inflight > 0 else wl.default_inflight)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the same keys and values as the GPU arm's line for this configuration; what the CPU arm actually ran per
        # step (a bounded sample of the workload) is said in cpu_baseline.sample
        "config": config_block(wl, args.gpus, {
            "l2": "flushed between steps (256 MB write)", "batches_in_flight": ref_f,
            "pipes_per_batch": args.pipes or (4 if ref_f <= 1 else 1 if ref_f >= 3 else 2),
            "layout": "auto",
            "parallelism": "problem-index sharding x%d, ONE NCCL all-gather of results, statuses and iteration counts per step (inline)" % args.gpus}),
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port",
                         "sample": "%s x %d steps, %d host threads (CasADi/IPOPT not installable offline; oracle port "
                                   "of IPOPT's algorithm)" % (sample, args.steps, threads)},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_ipm_iters": iters,
    }
    print(json.dumps(line), flush=True)


def counted_flops_per_iter(key):
    """FP64 operations the kernels actually execute per IPM iteration, from the committed ncu capture of this
    configuration (profiles/r2_counted_flops.json: dadd + dmul + 2 dfma thread instructions of one solve / its
    iterations).  None when no capture exists."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_counted_flops.json")))
        return float(d[key]["flop_per_iter_executed"])
    except Exception:
        return None


def measure(wl, args, world, rank, dev, gather_mode, steps, warmup, sampler=None, with_latency=False, inflight=1):
    """timed region of one workload: K steps, CUDA events on the launching stream, L2 flushed between steps, NCCL
    gather of the results after every step; then the end-to-end leg through the public API from host buffers.

    inflight = F > 1: the K steps alternate over F solver handles on F streams (batch k goes to handle k mod F), so
    that the sparse end of one batch (late sweeps, straggler tail: a few warps busy) runs underneath the dense sweeps
    of the next one.  Still exactly K solves of the full batch between the two synchronisation points; the serial
    figure (one batch at a time, F = 1) is measured in the same run and reported beside it."""
    from concurrent.futures import ThreadPoolExecutor

    import torch
    import torch.distributed as dist
    from mpc_verde_b200 import dist as mdist

    F = max(1, int(inflight))
    wls = [wl]
    for _ in range(F - 1):
        w = type(wl)(batch=wl.B, rank=wl.rank)
        w.setup(wl.mv, dev, wl.layout, wl.pipes)
        wls.append(w)
    main = torch.cuda.current_stream(dev)
    streams = [main] if F == 1 else [torch.cuda.Stream(dev) for _ in wls]
    flush = [torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev) for _ in streams]   # > 126 MB L2

    # --gather overlap: the collectives of step k run on a high-priority side stream (and NCCL's own high-priority
    # stream) underneath the solve of step k + 1; the region ends when the last gather has
    gstream = torch.cuda.Stream(dev, priority=-1) if (world > 1 and gather_mode == "overlap") else None

    def gather(w, outs, overlap=False):
        if world == 1 or gather_mode == "none":
            return None
        if overlap and gstream is not None:
            cur = torch.cuda.current_stream(dev)
            gstream.wait_stream(cur)
            with torch.cuda.stream(gstream):
                for t in list(outs) + [w.status, w.iters]:
                    t.record_stream(gstream)
                return mdist.gather_packed(outs, w.status, w.iters)
        return mdist.gather_packed(outs, w.status, w.iters)

    for w, st in zip(wls, streams):
        with torch.cuda.stream(st):
            for _ in range(max(warmup, 3) if w is wl else 2):
                gather(w, w.step())
        torch.cuda.synchronize()
    outs = wl.step()
    torch.cuda.synchronize()
    ok = bool((wl.status == 0).all())
    iters_sum = float(wl.iters.sum().item())

    def timed(n_steps, lanes, wls=wls):
        """n_steps solves round-robin over the first `lanes` handles / streams; returns (region ms, sum of the
        per-launch durations, flush ms that can be subtracted)"""
        evk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        fev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record(main)
        for st in streams[:lanes]:
            if st is not main:
                st.wait_event(t_start)
        o = None
        for k in range(n_steps):
            w, st = wls[k % lanes], streams[k % lanes]
            with torch.cuda.stream(st):
                fev[k][0].record()
                if not os.environ.get("BENCH_NO_FLUSH"):
                    flush[k % lanes].fill_(float(k))            # evict L2
                fev[k][1].record()
                evk[k][0].record()
                o = w.step()
                evk[k][1].record()
                gather(w, o, overlap=True)
        for st in streams[:lanes]:
            if st is not main:
                main.wait_stream(st)
        if gstream is not None:
            main.wait_stream(gstream)
        t_end.record(main)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        # one batch at a time: the flush is a separate interval of the region and is taken out; overlapped: it
        # runs underneath the other batch and stays in
        flush_ms = sum(a.elapsed_time(b) for a, b in fev) if lanes == 1 else 0.0
        return t_start.elapsed_time(t_end) - flush_ms, sum(a.elapsed_time(b) for a, b in evk), o

    serial = None
    if F > 1:
        # one batch at a time FIRST (nothing else in flight, no clock sampler running), on a handle with the library's
        # own pipe count (the in-flight handles run one pipe each); mean over its steps
        ws = wl
        if wl.pipes:
            ws = type(wl)(batch=wl.B, rank=wl.rank)
            ws.setup(wl.mv, dev, wl.layout, None)
            for _ in range(3):
                gather(ws, ws.step())
            torch.cuda.synchronize()
        s_steps = max(3, min(steps, 5))
        s_ms, sk_ms, _ = timed(s_steps, 1, [ws])
        serial = {"ms_per_step": s_ms / s_steps, "kernel_ms": sk_ms / s_steps, "steps": s_steps}
        del ws
    launches0 = sum(w.solver.kernel_count() for w in wls)
    if sampler is not None:
        sampler.start()
    t_ms, tk_ms, outs = timed(steps, F)
    if sampler is not None:
        sampler.stop()
    launches = sum(w.solver.kernel_count() for w in wls) - launches0
    if F > 1:
        tk_ms = t_ms                    # overlapped launches: the average launch duration is the region / K
    print("rank %d [%s]: %.3f ms per step in the timed region (%d in flight), %.3f ms per solve launch%s" %
          (rank, wl.key, t_ms / steps, F, tk_ms / steps,
           "" if serial is None else "; one batch at a time: %.3f ms per step" % serial["ms_per_step"]), file=sys.stderr)
    tt = torch.tensor([t_ms, tk_ms] + ([serial["ms_per_step"], serial["kernel_ms"]] if serial else []),
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, tk_ms = float(tt[0]), float(tt[1])
    n_solves = wl.solves_per_step()
    value = world * n_solves * steps / (t_ms * 1e-3)
    if serial:
        serial["ms_per_step"], serial["kernel_ms"] = float(tt[2]), float(tt[3])
        serial["value"] = world * n_solves / (serial["ms_per_step"] * 1e-3)

    # ---- end to end: host buffers in, host results out, the NCCL gather of the results included.  F callers (one
    # host thread per handle, as F workers of a service would) keep F batches in flight: the copies of one batch run
    # underneath the solve of the other.  Results are consumed (and gathered) in step order. ----
    for w in wls:
        w.host_step()
    e2e_steps = max(3, min(steps, 5)) * F
    pools = [ThreadPoolExecutor(1) for _ in wls]

    def host_job(w, st, k):
        with torch.cuda.device(dev), torch.cuda.stream(st):
            if not os.environ.get("BENCH_NO_FLUSH"):
                flush[k % F].fill_(float(k))                 # the same L2 eviction as in the device-timed region
                st.synchronize()
            res = w.host_step()
            if world > 1 and gather_mode != "none":
                res = [r.to(dev, non_blocking=True) for r in res]      # (the page-locked result arrays are reused)
                st.synchronize()
            return res
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    futs = [pools[k % F].submit(host_job, wls[k % F], streams[k % F], k) for k in range(e2e_steps)]
    for k, fu in enumerate(futs):
        res = fu.result()
        if world > 1 and gather_mode != "none":
            gather(wls[k % F], res)
            torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    for pl in pools:
        pl.shutdown()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_solves * e2e_steps / float(te[0])
    h2d, d2h = wl.io_bytes()

    out = {"value": value, "ms_per_step": t_ms / steps, "kernel_ms": tk_ms / steps, "launches": int(launches),
           "iters_sum": iters_sum, "n_solves": n_solves, "ok": ok, "inflight": F, "serial": serial,
           "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "steps": e2e_steps, "batches_in_flight": F,
                   "includes_gather": bool(world > 1 and gather_mode != "none")},
           "outs": outs}
    if with_latency and not wl.closed_loop:
        lat = wl.solver.enable_latency(wl.B)
        wl.step()
        torch.cuda.synchronize()
        lat_us = lat.cpu().numpy() / 1e3
        wl.solver.disable_latency()
        out["p50_solve_us"], out["p99_solve_us"] = float(np.median(lat_us)), float(np.percentile(lat_us, 99))
    del wls[1:]
    return out


def roofline_of(wl, m, peak_tf, hbm_peak, hbm_src):
    flop_formula = wl.flop_per_iter
    counted = counted_flops_per_iter(wl.key)
    per_iter = min(flop_formula, counted) if counted else flop_formula
    flops = m["iters_sum"] * per_iter
    achieved = flops / (m["kernel_ms"] * 1e-3) / 1e12
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))[wl.key]["dram_bytes_per_step"]
    except Exception:
        pass
    one = None
    if m.get("serial"):
        a1 = flops / (m["serial"]["kernel_ms"] * 1e-3) / 1e12
        one = {"kernel_ms": m["serial"]["kernel_ms"], "achieved": a1, "frac": a1 / peak_tf}
    return {
        "bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
        # with several batches in flight the launches overlap: kernel_ms is the timed region / K (what one launch costs
        # the GPU); `one_batch_at_a_time` is the same launch timed alone with CUDA events around it
        "kernel_ms_is": "timed region / steps" if m.get("inflight", 1) > 1 else "CUDA events around each launch",
        "one_batch_at_a_time": one,
        "traffic": traffic,
        "traffic_source": "committed ncu capture profiles/r2_traffic.json (not measured in this run)" if traffic else None,
        "peak_source": "FP64 FMA peak measured in this run by mpcv_fp64_peak (MEASURED_PEAKS.json has no FP64 figure)",
        "flop_per_iter": {"formula": flop_formula, "executed_ncu": counted, "used": per_iter},
        "flop_per_launch": flops, "kernel_ms": m["kernel_ms"],
        "hbm": {"achieved_gbs": m["n_solves"] * wl.algorithmic_bytes_per_solve() / (m["kernel_ms"] * 1e-3) / 1e9,
                "peak_gbs": hbm_peak, "peak_source": hbm_src},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="problems / scenarios per GPU (default: the config's)")
    ap.add_argument("--layout", type=int, default=S.LAYOUT_AUTO)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="do not attach short runs of the other configurations")
    ap.add_argument("--inflight", type=int, default=0,
                    help="batches in flight: the K steps alternate over this many solver handles / streams.  Batches in "
                         "flight sit in different phases of the solve (a DRAM-bound Riccati sweep of one under the FP64-bound "
                         "derivative sweep of another, the one-warp straggler tail of one under the dense sweeps of the "
                         "next), which the lock-step pipes of ONE batch cannot.  Measured on one B200, C2, ms per batch "
                         "(in flight x pipes per batch): 1x4 14.1..15.0, 2x2 12.1..12.5, 3x1 11.75, 4x1 11.5..11.8, 8x1 "
                         "11.5; four B200s with rank 2's 119-iteration straggler: 1x4 20.9..28.6, 2x4 14.3.  1 = one batch at "
                         "a time (the round-1 arrangement; measured in the same run and reported as `serial`); 0 = the "
                         "configuration's own default (4 for c2 / c4 / c4f; 1 for the HBM-bound linear c3 / c5 and for the "
                         "single-problem latency config c1)")
    ap.add_argument("--seed-rank", type=int, default=None,
                    help="diagnostic: give this process the synthetic batch of another rank (rank 2's C2 batch holds the "
                         "119-iteration straggler)")
    ap.add_argument("--pipes", type=int, default=0,
                    help="pipes per batch (mpcv_set_knob phase_pipes); 0 = 1 pipe when 3 or more batches are in flight, 2 "
                         "with two, the library default (4) with one")
    ap.add_argument("--gather", default="inline", choices=["inline", "overlap", "none"],
                    help="N>1: NCCL gather of the results in stream order after each step; overlap: on a high-priority "
                         "side stream underneath the next step's solve; none: diagnostic")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import mpc_verde_b200 as mv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU; there is no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, pg_options=opts)

    def pipes_for(f):
        return args.pipes or (1 if f >= 3 else 2 if f == 2 else None)

    wl = CONFIGS[args.config](batch=args.batch, rank=rank)
    def inflight_for(w):
        return 1 if w.key == "c1" else (args.inflight if args.inflight > 0 else w.default_inflight)

    inflight = inflight_for(wl)
    wl.setup(mv, dev, args.layout, pipes_for(inflight))
    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("BENCH_NO_SAMPLER") else None
    m = measure(wl, args, world, rank, dev, args.gather, args.steps, args.warmup, sampler, with_latency=True,
                inflight=inflight)
    assert m["ok"], "solver failures in the benchmark batch"

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = peaks.get("hbm_gbs", 6650.0), ("MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback")
    peak_tf, _ = mv.fp64_peak()

    # ---- the other BASELINE.json configurations, short runs (same timing rules, 3 warm-up + 2 timed steps) ----
    others = {}
    if not args.no_others:
        for key in ("c1", "c3", "c4", "c4f", "c5"):
            if key == args.config:
                continue
            try:
                w2 = CONFIGS[key](rank=rank)
                f2 = inflight_for(w2)
                w2.setup(mv, dev, S.LAYOUT_AUTO, pipes_for(f2))
                m2 = measure(w2, args, world, rank, dev, args.gather, 2 * f2, 3, inflight=f2)
                o = {"config": config_block(w2, world, {"batches_in_flight": f2, "pipes_per_batch": pipes_for(f2) or 4}),
                     "value": m2["value"], "unit": "solves/s",
                     "ms_per_step": m2["ms_per_step"], "serial": m2["serial"],
                     "all_succeeded": m2["ok"], "mean_ipm_iters": m2["iters_sum"] / m2["n_solves"], "e2e": m2["e2e"],
                     "gpu_launches": m2["launches"], "roofline": roofline_of(w2, m2, peak_tf, hbm_peak, hbm_src)}
                if rank == 0:
                    o["max_abs_diff_vs_oracle_sample"] = w2.check(m2["outs"])
                    if key == "c1":
                        ms_cpu, st = w2.cpu_loop_ms_per_solve()
                        o["latency_pair"] = {"gpu_ms_per_solve": m2["ms_per_step"] / 84.0, "cpu_oracle_ms_per_solve": ms_cpu,
                                             "mpc_steps": st}
                others[key] = o
                del w2, m2
                torch.cuda.empty_cache()
            except Exception as e:      # a failing side configuration must not take the headline line down
                others[key] = {"error": "%s: %s" % (type(e).__name__, e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # one problem alone (the latency the batch figures hide)
    lone = None
    if not wl.closed_loop:
        w1 = CONFIGS[wl.key](batch=1)
        w1.setup(mv, dev, args.layout)
        for _ in range(3):
            w1.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            w1.step()
        e1.record()
        torch.cuda.synchronize()
        lone = e0.elapsed_time(e1) / 5

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        if wl.key == "c1":
            ms_cpu, st = wl.cpu_loop_ms_per_solve()
            cpu = {"value": 1e3 / ms_cpu, "unit": "solves/s", "cores": 1, "kind": "port",
                   "sample": "the %d-step closed loop of the one problem, single thread" % st}
        elif not wl.closed_loop:
            n_cpu = (512 if wl.key == "c2" else 64) * threads
            n, dt, _ = wl.cpu_sample(n_cpu, threads)
            cpu = {"value": n / dt, "unit": "solves/s", "cores": threads, "kind": "port",
                   "sample": "first %d problems of the same synthetic batch, %d host threads, %.1f s wall" % (n, threads, dt)}
        else:
            from oracle import mpc_oracle
            n = min(wl.B, 16)
            a = wl._oracle_args(n) if hasattr(wl, "_oracle_args") else None
            if a is None and hasattr(wl, "cpu_first_solves"):
                n, dt, _ = wl.cpu_first_solves(256 * threads, threads)
                cpu = {"value": n / dt, "unit": "solves/s", "cores": threads, "kind": "port",
                       "sample": "first (cold) solve of the loops of %d scenarios, %d host threads, %.1f s wall" % (n, threads, dt)}
            if a is not None:
                t0 = time.perf_counter()
                mpc_oracle.closed_loop(wl.spec, *a)
                dt = time.perf_counter() - t0
                cpu = {"value": n * wl.n_steps / dt, "unit": "solves/s", "cores": 1, "kind": "port",
                       "sample": "first %d scenarios x %d steps, one host thread, %.1f s wall" % (n, wl.n_steps, dt)}

    line = {
        "metric": METRIC, "value": m["value"], "unit": "solves/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(wl, world, {
            "l2": "flushed between steps (256 MB write)", "batches_in_flight": inflight,
            "pipes_per_batch": pipes_for(inflight) or 4,
            "layout": {0: "auto", 1: "thread-per-problem", 2: "warp-per-problem", 3: "phase kernels", 4: "CTA-resident"}[args.layout],
            "parallelism": "problem-index sharding x%d, ONE NCCL all-gather of results, statuses and iteration counts per step (%s)" % (world, args.gather)}),
        "e2e": m["e2e"],
        "serial": m["serial"],
        "serial_note": "the same batch with ONE batch in flight (one handle with the library's default four pipes, one stream: the round-1 arrangement)",
        "gpu_launches": m["launches"],
        "clocks": sampler.summary() if sampler is not None else None,
        "roofline": roofline_of(wl, m, peak_tf, hbm_peak, hbm_src),
        "cpu_baseline": cpu,
        "mean_ipm_iters": m["iters_sum"] / m["n_solves"],
        "p50_solve_us": m.get("p50_solve_us"), "p99_solve_us": m.get("p99_solve_us"),
        "p50_note": "batch-residency latency (entry to convergence inside the batch-synchronous pipeline), not a lone-solve latency",
        "lone_problem_ms": lone,
        "batch_us_per_solve": m["ms_per_step"] * 1e3 / m["n_solves"],
        "other_configs": others,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
