#!/usr/bin/env python3
"""One BASELINE configuration's step, a few times, for ncu launch lists (run with MPCV_PHASE_HOSTLOOP=1 so that
every phase kernel is a plain launch):

    MPCV_PHASE_HOSTLOOP=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/X_launches.csv python scripts/profile_config.py --config c3
    python scripts/launch_sweeps.py gpurun_out/X_launches.csv
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                     # noqa: E402

import bench                     # noqa: E402
import mpc_verde_b200 as mv      # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--layout", type=int, default=0)
    a = ap.parse_args()
    wl = bench.CONFIGS[a.config](batch=a.batch)
    wl.setup(mv, torch.device("cuda", 0), a.layout)
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        wl.step()
        e1.record()
        torch.cuda.synchronize()
        print("%s: %.3f ms, all ok %s, mean iters %.2f" % (a.config, e0.elapsed_time(e1), bool((wl.status == 0).all()),
                                                          float(wl.iters.float().mean())), flush=True)
