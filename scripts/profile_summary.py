#!/usr/bin/env python3
"""Turn the ncu outputs of one profiling call into the tracked summaries under profiles/.

    python scripts/profile_summary.py launches gpurun_out/launches_<tag>.csv <tag>
        -> profiles/<tag>_launches_one_solve.csv, <tag>_kernel_shares.md, <tag>_traffic.json
           (the FIRST complete solve of the log: ph_begin_kernel ... ph_tail_kernel)
    python scripts/profile_summary.py full gpurun_out/prof_<tag>.ncu-rep <tag>
        -> profiles/<tag>_full_metrics.csv (one column per captured kernel)

The launch list comes from `ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,
smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,
smsp__thread_inst_executed_per_inst_executed.ratio --csv` in host-loop mode (MPCV_PHASE_HOSTLOOP=1: ncu does not
step into the conditional WHILE node of the solve graph; the host loop issues the same kernels, one pipe).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL_METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def short(name):
    k = name.split("(")[0].replace("void ", "").replace("mpcv::", "")
    return k.replace("Unicycle<(int)0>", "Unicycle<0>")


def launches(path, tag):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    byid = collections.OrderedDict()
    for r in rows:
        d = byid.setdefault(r["ID"], {"Kernel": short(r["Kernel Name"]), "Block": r["Block Size"], "Grid": r["Grid Size"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    L = list(byid.values())
    begins = [i for i, d in enumerate(L) if d["Kernel"].startswith("ph_begin")]
    tails = [i for i, d in enumerate(L) if d["Kernel"].startswith("ph_tail")]
    lo = begins[0]
    hi = [t for t in tails if t > lo][0]
    L = L[lo:hi + 1]
    out = os.path.join(ROOT, "profiles", tag + "_launches_one_solve.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel", "Block", "Grid", "gpu__time_duration.sum [ns]", "registers", "smsp__inst_executed.sum",
                    "active_threads_per_inst", "dram_read_bytes", "dram_write_bytes"])
        for i, d in enumerate(L):
            w.writerow([i, d["Kernel"], d["Block"], d["Grid"], int(d.get("gpu__time_duration.sum", 0)),
                        int(d.get("launch__registers_per_thread", 0)), int(d.get("smsp__inst_executed.sum", 0)),
                        "%.1f" % d.get("smsp__thread_inst_executed_per_inst_executed.ratio", 0.0),
                        int(d.get("dram__bytes_read.sum", 0)), int(d.get("dram__bytes_write.sum", 0))])
    agg = collections.OrderedDict()
    for d in L:
        k = d["Kernel"].split("<")[0]
        a = agg.setdefault(k, [0, 0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] = max(a[1], int(d.get("launch__registers_per_thread", 0)))
        a[2] += d.get("gpu__time_duration.sum", 0) / 1e3
        a[3] += d.get("smsp__inst_executed.sum", 0) / 1e6
        a[4] += (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6
    tot = sum(a[2] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", tag + "_kernel_shares.md"), "w") as f:
        f.write("| kernel | launches / solve | registers | total us | share | warp-instr (M) | DRAM bytes (MB) |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
            f.write("| `%s` | %d | %d | %.0f | %.1f%% | %.1f | %.0f |\n" % (k, a[0], a[1], a[2], 100 * a[2] / tot, a[3], a[4]))
        f.write("| **sum** | %d | | %.0f | 100%% | %.1f | %.0f |\n" % (len(L), tot, sum(a[3] for a in agg.values()),
                                                                     sum(a[4] for a in agg.values())))
    json.dump({"source": "profiles/%s_launches_one_solve.csv (ncu --metrics ..., host-loop mode, one 65,536-problem solve)" % tag,
               "dram_bytes_per_solve_launch": sum(a[4] for a in agg.values()) * 1e6, "kernel_us_sum": tot, "launches": len(L)},
              open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w"), indent=1)
    print(open(os.path.join(ROOT, "profiles", tag + "_kernel_shares.md")).read())


def full(path, tag):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--metrics", ",".join(FULL_METRICS)],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [short(r[hdr.index("Kernel Name")]).split("<")[0] for r in data]
    with open(os.path.join(ROOT, "profiles", tag + "_full_metrics.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                w.writerow([m, units[i]] + [r[i] for r in data])
    print("kernels:", names)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
