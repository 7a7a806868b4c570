import sys, os, json
sys.path.insert(0, '/root/repo')
import torch, bench
import mpc_verde_b200 as mv
from mpc_verde_b200 import spec as S
dev = torch.device("cuda", 0)
class A: pass
args = A()
order = sys.argv[1].split(",")
for key in order:
    w = bench.CONFIGS[key]()
    w.setup(mv, dev, S.LAYOUT_AUTO)
    m = bench.measure(w, args, 1, 0, dev, "inline", 2, 3)
    print(sys.argv[1], key, round(m["ms_per_step"], 2), flush=True)
    if len(sys.argv) > 2 and key != order[-1]:
        w.check(m["outs"])
    del w, m
    torch.cuda.empty_cache()
