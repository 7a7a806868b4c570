"""Latency of small batches through the phase pipeline (B = 1, 32, 1024): a batch this small goes straight to the
straggler tail kernel, so ms per solve / iterations of the longest problem is the lone-problem iteration latency.

    python scripts/latency_probe.py        (needs a GPU)
"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from tests import common
import mpc_verde_b200 as mv
from mpc_verde_b200 import problems, spec as S
from tests.test_gpu_parity import _solver
solver = _solver(mv, problems.unicycle_multiple_shooting(), layout=S.LAYOUT_PHASED)
sp = solver.spec
lbx, ubx = problems.unicycle_bounds(sp, x_box=20.0)
for B in (1, 32, 1024):
    x0s, p = common.unicycle_batch(B, seed=5)
    w0 = torch.as_tensor(problems.cold_start(sp, x0s)).cuda(); pd = torch.as_tensor(p).cuda()
    lb, ub = torch.as_tensor(lbx).cuda(), torch.as_tensor(ubx).cuda()
    for _ in range(3): solver(x0=w0, lbx=lb, ubx=ub, p=pd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): solver(x0=w0, lbx=lb, ubx=ub, p=pd)
    e1.record(); torch.cuda.synchronize()
    it = np.asarray(solver.stats()["iter_count"]).reshape(-1)
    ms = e0.elapsed_time(e1) / 20
    print("B=%d: %.3f ms per solve, iterations max %d mean %.1f -> %.1f us per iteration of the longest problem" % (B, ms, it.max(), it.mean(), 1e3 * ms / it.max()))
