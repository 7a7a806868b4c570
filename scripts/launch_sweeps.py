#!/usr/bin/env python3
"""Per-sweep and per-kernel view of an ncu launch list of one phase-pipeline solve (host-driven loop):
    python scripts/launch_sweeps.py gpurun_out/XXX_launches.csv [solve_index]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
L = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < 15:
        continue
    name = re.sub(r"[<(].*", "", r[4]).replace("mpcv::", "").replace("void ", "")
    d = L.setdefault(int(r[0]), {"name": name})
    d[r[12]] = float(r[14].replace(",", ""))
begins = [i for i in L if L[i]["name"].startswith("ph_begin")]
start = begins[which]
end = begins[which + 1] if which != -1 and which + 1 < len(begins) else max(L) + 1
sw, cur, tot = [], None, collections.defaultdict(lambda: [0, 0.0, 0.0])
for i in L:
    if i < start or i >= end:
        continue
    d = L[i]
    t = d["gpu__time_duration.sum"] / 1e3
    mb = (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6
    k = d["name"].replace("ph_", "").replace("_kernel", "")
    tot[k][0] += 1; tot[k][1] += t; tot[k][2] += mb
    if d["name"] == "ph_pre_kernel":
        cur = {"t": 0.0, "mb": 0.0, "k": collections.OrderedDict()}
        sw.append(cur)
    if cur is not None:
        cur["t"] += t; cur["mb"] += mb; cur["k"][k] = (t, mb)
for j, s in enumerate(sw):
    print("%2d %5.0f us %5.0f MB | " % (j, s["t"], s["mb"]) + " ".join("%s=%.0f/%.0f" % (k, v[0], v[1]) for k, v in s["k"].items() if v[0] > 8))
T = sum(v[1] for v in tot.values())
print("total %.2f ms, %.2f GB" % (T / 1e3, sum(v[2] for v in tot.values()) / 1e3))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-14s %3d launches %8.3f ms %5.1f %% %7.2f GB" % (k, v[0], v[1] / 1e3, 100 * v[1] / T, v[2] / 1e3))
