import sys; sys.path.insert(0,'/root/repo')
import torch, bench
import mpc_verde_b200 as mv
from mpc_verde_b200 import spec as S
dev = torch.device("cuda", 0)
def t(w, n=5, stream=None):
    st = stream or torch.cuda.current_stream()
    with torch.cuda.stream(st):
        for _ in range(3): w.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): w.step()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
w4 = bench.C2(); w4.setup(mv, dev, S.LAYOUT_AUTO, None)
print("default handle alone", t(w4))
ws = [bench.C2() for _ in range(4)]
for w in ws: w.setup(mv, dev, S.LAYOUT_AUTO, 1)
print("pipes=1 handle alone", t(ws[0]))
print("default handle again (5 handles alive)", t(w4))
w5 = bench.C2(); w5.setup(mv, dev, S.LAYOUT_AUTO, None)
print("new default handle (6 alive)", t(w5))
print("new default handle on side stream", t(w5, stream=torch.cuda.Stream()))
print("mem", torch.cuda.mem_get_info())
