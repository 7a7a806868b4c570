#!/usr/bin/env python3
"""Fresh handles, same batch: are the results the same bits?  (rank 2's C2 batch holds problem 6,236, whose 119-iteration
path amplifies any difference)   python scripts/determinism_probe.py [--rank 2] [--reps 4]"""
import argparse
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np   # noqa: E402
import torch         # noqa: E402

import bench         # noqa: E402
import mpc_verde_b200 as mv   # noqa: E402
from mpc_verde_b200 import spec as S   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rank", type=int, default=2)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
dev = torch.device("cuda", 0)
for pipes in (None, 1):
    for rep in range(a.reps):
        # dirty the allocator's memory so that a fresh slab does not come back zeroed or with the last run's contents
        junk = torch.full((300 * 1024 * 1024 // 8,), float(rep + 1) * 1e300, dtype=torch.float64, device=dev)
        del junk
        wl = bench.C2(rank=a.rank)
        wl.setup(mv, dev, S.LAYOUT_AUTO, pipes)
        outs = wl.step()
        torch.cuda.synchronize()
        x = outs[0].cpu().numpy()
        it = wl.iters.cpu().numpy()
        print("pipes", pipes, "rep", rep, "max iters", int(it.max()), "at", int(it.argmax()), "iters[6236]", int(it[6236]),
              "sha", hashlib.sha1(x.tobytes()).hexdigest()[:12], "sha iters", hashlib.sha1(it.tobytes()).hexdigest()[:12], flush=True)
        del wl
        torch.cuda.empty_cache()
