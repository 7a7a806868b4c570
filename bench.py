#!/usr/bin/env python3
"""bench.py — headline benchmark: batched MPC solves/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU arm (oracle port of CasADi/IPOPT)

Workload (BASELINE.json configs[1]): unicycle multiple shooting (Casadi/multiple_shooting_casadi.py),
N=10, T=0.2, RK4 M=4 + quadrature cost, batch of 65,536 random initial states per GPU with
state/control bounds, cold start X_k = x0, U = 0, one solve per problem, FP64, tol 1e-8.
A "step" is one batched solve of the whole per-GPU batch.  Weak scaling: every rank owns its own
65,536 problems (sharded by problem index, no collective on the hot path); after the solve NCCL
gathers the results and all-reduces a small statistics vector.

Printed keys (one JSON line on rank 0): see the contract in the task description; additionally
  roofline     — FP64 vector-pipe roofline of the solve kernel (this path is neither HBM- nor
                 tensor-bound: ~360 FLOP/B; tensor cores deliberately unused)
  cpu_baseline — the oracle timed on the box's host cores on a bounded sample
  p50_solve_us — median per-problem latency from device %globaltimer stamps
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mpc_verde_b200 import problems  # noqa: E402
from mpc_verde_b200 import spec as S  # noqa: E402

B_PER_GPU = 65536
SEED = 20261
# ALGORITHMIC work per IPM iteration of the C2 problem:
#   N*(C_HJ + C_ric + C_bar) + n_ls*N*C_val + C_norm = 10*(1868+242+50) + 10*283 + 300 = 24,730 FLOP
FLOP_PER_ITER_C2 = 24730.0
# algorithmic HBM bytes per solve: 8*(n_p + n_var_in + n_var_out + 1) + 8 (status, iters)
BYTES_PER_SOLVE_C2 = 8 * (6 + 53 + 53 + 1) + 8

METRIC = "batched MPC solves/sec (RK4 NLP to IPOPT tol)"


def make_batch(spec, B, seed):
    rng = np.random.default_rng(seed)
    x0s = np.stack([rng.uniform(-2, 12, B), rng.uniform(-2, 12, B), rng.uniform(-math.pi, math.pi, B)], 1)
    p = np.concatenate([x0s, np.tile([10.0, 10.0, 0.0], (B, 1))], 1)
    lbx, ubx = problems.unicycle_bounds(spec, x_box=20.0)
    return problems.cold_start(spec, x0s), lbx, ubx, p


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.1)

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


def cpu_baseline(spec, n_problems, threads):
    """The oracle (port of the reference's CasADi/IPOPT path) on the host cores, bounded sample."""
    from oracle import mpc_oracle
    w0, lbx, ubx, p = make_batch(spec, n_problems, SEED)
    mpc_oracle.solve(spec, w0[:threads], lbx, ubx, p[:threads], nthreads=threads)   # warm the library
    t0 = time.perf_counter()
    r = mpc_oracle.solve(spec, w0, lbx, ubx, p, nthreads=threads)
    dt = time.perf_counter() - t0
    assert np.all(r["status"] == 0)
    return n_problems / dt, dt, float(r["iters"].mean())


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  CasADi/IPOPT cannot
    be installed here (no network; not in the wheelhouse), so this arm times the oracle port with
    every host thread; each step is a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = S.unicycle_multiple_shooting()
    threads = os.cpu_count() or 1
    sample = 256 * threads
    from oracle import mpc_oracle
    w0, lbx, ubx, p = make_batch(spec, sample, SEED)
    for _ in range(max(args.warmup, 1)):
        mpc_oracle.solve(spec, w0[:threads * 8], lbx, ubx, p[:threads * 8], nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = mpc_oracle.solve(spec, w0, lbx, ubx, p, nthreads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 unicycle multiple shooting N=10 T=0.2 RK4(M=4)+quadrature, cold start, "
                               "x,y in [-20,20], v in [-1,1], w in [-pi/4,pi/4]",
                   "batch_per_step": sample, "seed": SEED},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port",
                         "sample": "%d problems per step x %d steps, %d host threads (CasADi/IPOPT not installable "
                                   "offline; oracle port of IPOPT's algorithm)" % (sample, args.steps, threads)},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_ipm_iters": float(r["iters"].mean()),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="problems per GPU per step")
    ap.add_argument("--layout", type=int, default=S.LAYOUT_AUTO)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="inline", choices=["inline", "overlap", "none"],
                    help="N>1: NCCL gather of the results in stream order after each solve, or on a side stream under the next solve")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import mpc_verde_b200 as mv
    from mpc_verde_b200 import dist as mdist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU; there is no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # (NCCL channel caps were tried for the side-stream gather: NCCL_MAX_NCHANNELS=4 helps at N=2,
        # 21.8 vs 24.4 ms/step, but hurts at N=8, 27.1 vs 24.0 — the defaults stay.)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    prob = problems.unicycle_multiple_shooting()
    solver = mv.nlpsol("solver", "ipopt", prob, {"ipopt": {"max_iter": 2000, "print_level": 0, "acceptable_tol": 1e-8,
                                                           "acceptable_obj_change_tol": 1e-6},
                                                  "print_time": 0, "layout": args.layout})
    spec = solver.spec
    B = args.batch
    w0_h, lbx, ubx, p_h = make_batch(spec, B, SEED + rank)       # every rank owns different problems
    w0 = torch.as_tensor(w0_h).to(dev)
    p = torch.as_tensor(p_h).to(dev)
    lb, ub = torch.as_tensor(lbx).to(dev), torch.as_tensor(ubx).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # > 126 MB L2

    def step_device():
        sol = solver(x0=w0, lbx=lb, ubx=ub, p=p, outputs=("x", "f"))
        return sol

    # result gather + statistics over NCCL, no host synchronisation inside; the timed region ends only after the
    # last gather has finished.  "inline": in stream order right after the solve (0.3-0.8 ms per step).  "overlap":
    # on a side stream underneath the next step's solve — measured slower since the solve runs several pipes
    # at once: the NCCL kernels of the two ranks wait for SM slots behind the other rank's sweeps (N=2: 33.0 ms
    # per step against 17.4 ms at N=1).
    gstream = torch.cuda.Stream(device=dev) if (world > 1 and args.gather == "overlap") else None

    def gather(sol):
        if world == 1 or args.gather == "none":      # "none": diagnostic (solve time per rank without any collective)
            return None
        st, it = solver._last
        if gstream is None:
            return mdist.gather_rows_equal(sol["x"]), mdist.gather_rows_equal(sol["f"]), mdist.reduce_stats_device(st, it)
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(gstream):
            gstream.wait_event(done)
            for t in (sol["x"], sol["f"], st, it):
                t.record_stream(gstream)
            xs = mdist.gather_rows_equal(sol["x"])
            fs = mdist.gather_rows_equal(sol["f"])
            stats = mdist.reduce_stats_device(st, it)
        return xs, fs, stats

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        sol = step_device()
        gather(sol)
    torch.cuda.synchronize()
    st, it = solver._last
    assert bool((st == 0).all()), "solver failures in the benchmark batch"
    iters_sum = float(it.sum().item())
    iters_mean = iters_sum / B

    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = solver.kernel_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush_ms = 0.0
    fev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start.record()
    for k in range(args.steps):
        fev[k][0].record()
        flush.fill_(float(k))                       # evict L2; its duration is subtracted below
        fev[k][1].record()
        evk[k][0].record()
        sol = step_device()
        evk[k][1].record()
        last = gather(sol)
    if gstream is not None:
        torch.cuda.current_stream().wait_stream(gstream)
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = solver.kernel_count() - launches0
    flush_ms = sum(a.elapsed_time(b) for a, b in fev)
    t_ms = t_start.elapsed_time(t_end) - flush_ms
    tk_ms = sum(a.elapsed_time(b) for a, b in evk)
    if rank == 0:
        sampler.stop_flag = True
    print("rank %d: %.3f ms per step in the timed region, %.3f ms per solve" % (rank, t_ms / args.steps, tk_ms / args.steps),
          file=sys.stderr)
    tt = torch.tensor([t_ms, tk_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, tk_ms = float(tt[0]), float(tt[1])
    value = world * B * args.steps / (t_ms * 1e-3)

    # ---- end-to-end through the public API with HOST buffers: every step copies its inputs host->device
    # from page-locked memory and its results (x, f, status, iters) device->host, inside the timed region ----
    w0_p, p_p = torch.as_tensor(w0_h).pin_memory(), torch.as_tensor(p_h).pin_memory()
    solver(x0=w0_p, lbx=lbx, ubx=ubx, lbg=0, ubg=0, p=p_p, outputs=("x", "f"))     # warm the staging buffers
    e2e_steps = max(3, min(args.steps, 5))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sol_h = solver(x0=w0_p, lbx=lbx, ubx=ubx, lbg=0, ubg=0, p=p_p, outputs=("x", "f"))
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(te[0])
    h2d = 8 * (w0_h.size + p_h.size + lbx.size + ubx.size)
    d2h = 8 * (B * spec.n_var + B) + 4 * 2 * B

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-problem latency (separate, untimed pass) ----
    lat = solver.enable_latency(B)
    step_device()
    torch.cuda.synchronize()
    lat_us = lat.cpu().numpy() / 1e3
    solver.disable_latency()

    # ---- roofline of the solve launch: ONE CUDA graph per step (init chain + a conditional WHILE node holding
    # the 12-kernel iteration sweep, one graph per pipe); its duration is taken with CUDA events on the launching stream ----
    peak_tf, _ = mv.fp64_peak()
    kernel_ms = tk_ms / args.steps
    flops = iters_sum * FLOP_PER_ITER_C2
    achieved_tf = flops / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None      # DRAM bytes of one solve launch from the committed ncu launch list (profiles/)
    try:
        if B == B_PER_GPU and args.layout in (S.LAYOUT_AUTO, S.LAYOUT_PHASED):
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1d_traffic.json")))["dram_bytes_per_solve_launch"]
    except Exception:
        pass
    roofline = {
        "bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
        "traffic": traffic,
        "peak_source": "FP64 FMA peak measured in this run by mpcv_fp64_peak (MEASURED_PEAKS.json has no FP64 figure)",
        "flop_per_launch": flops, "kernel_ms": kernel_ms,
        "hbm": {"achieved_gbs": B * BYTES_PER_SOLVE_C2 / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        n_cpu = 512 * threads
        v, dt, _ = cpu_baseline(spec, n_cpu, threads)
        cpu = {"value": v, "unit": "solves/s", "cores": threads, "kind": "port",
               "sample": "first %d problems of the same synthetic batch, %d host threads, %.1f s wall" % (n_cpu, threads, dt)}

    line = {
        "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 unicycle multiple shooting N=10 T=0.2 RK4(M=4)+quadrature, cold start, "
                               "x,y in [-20,20], v in [-1,1], w in [-pi/4,pi/4]",
                   "batch_per_gpu": B, "global_batch": world * B, "seed": SEED, "l2": "flushed between steps (256 MB write)",
                   "layout": {0: "auto (phase kernels)", 1: "thread-per-problem", 2: "warp-per-problem", 3: "phase kernels"}[args.layout],
                   "parallelism": "problem-index sharding x%d, NCCL all-gather of results + stats (%s)" % (world, args.gather)},
        "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "mean_ipm_iters": iters_mean,
        "p50_solve_us": float(np.median(lat_us)), "p99_solve_us": float(np.percentile(lat_us, 99)),
        "batch_us_per_solve": t_ms * 1e3 / (args.steps * B),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
