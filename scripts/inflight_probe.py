#!/usr/bin/env python3
"""Batches in flight: K solves of the C2 batch issued on one stream through one handle (serial) against the same K
solves alternating over F handles on F streams, so that the sparse end of one batch (late sweeps, straggler tail)
runs underneath the dense sweeps of the next.  One JSON line per (seed, F).

    python scripts/inflight_probe.py [--seeds 20261,20263] [--inflight 1,2,3] [--steps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

import bench         # noqa: E402
import mpc_verde_b200 as mv   # noqa: E402


def run(seed, F, steps, B):
    dev = torch.device("cuda", 0)
    wls = []
    for f in range(F):
        wl = bench.C2(batch=B)
        wl.seed = seed
        wl.setup(mv, dev)
        wls.append(wl)
    streams = [torch.cuda.Stream() for _ in range(F)]
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    for wl, st in zip(wls, streams):
        with torch.cuda.stream(st):
            for _ in range(3):
                wl.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    done = [torch.cuda.Event() for _ in range(F)]
    main = torch.cuda.current_stream()
    e0.record()
    for k in range(steps):
        st, wl = streams[k % F], wls[k % F]
        st.wait_stream(main) if k < F else None
        with torch.cuda.stream(st):
            flush.fill_(float(k)) if F == 1 else None
            outs = wl.step()
    for st in streams:
        main.wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ok = all(bool((wl.status == 0).all()) for wl in wls)
    print(json.dumps({"seed": seed, "inflight": F, "steps": steps, "ms_per_step": ms / steps,
                      "solves_per_s": B * steps / ms * 1e3, "ok": ok, "max_iters": int(wls[0].iters.max())}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", default="20261,20263")
    ap.add_argument("--inflight", default="1,2,3")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--batch", type=int, default=65536)
    a = ap.parse_args()
    for seed in [int(s) for s in a.seeds.split(",")]:
        for F in [int(s) for s in a.inflight.split(",")]:
            run(seed, F, a.steps, a.batch)
