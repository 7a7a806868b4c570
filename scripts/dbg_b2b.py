import sys, os, time, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench, mpc_verde_b200 as mv
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
wl = bench.C2(); wl.setup(mv, dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
for _ in range(4): wl.step()
torch.cuda.synchronize()
for mode in ("sync", "nosync", "nosync_noflush"):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    t0 = time.perf_counter()
    for k in range(8):
        if mode != "nosync_noflush": flush.fill_(float(k))
        ev[k][0].record(); wl.step(); ev[k][1].record()
        if mode == "sync": torch.cuda.synchronize()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(os.environ.get("MPCV_RES_TAIL"), mode, ["%.2f" % a.elapsed_time(b) for a, b in ev], "host enqueue %.1f ms total %.1f ms" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3), flush=True)
