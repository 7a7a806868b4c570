#!/usr/bin/env python3
"""Registers / stack / spills per kernel from the build's ptxas logs (csrc/build/*.ptxas.log); with two directories:
a side-by-side diff.   python scripts/ptxas_summary.py [dir_a] [dir_b] [--model 0]"""
import glob
import os
import re
import subprocess
import sys


def parse(d, model):
    out = {}
    for f in sorted(glob.glob(os.path.join(d, "*_%s.ptxas.log" % model))):
        t = open(f).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, "
                             r"(\d+) bytes spill loads\n.*?Used (\d+) registers", t, re.S):
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void mpcv::", "").replace("mpcv::", "")
            out[name] = (int(m.group(5)), int(m.group(2)), int(m.group(3)))
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    model = sys.argv[sys.argv.index("--model") + 1] if "--model" in sys.argv else "0"
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mpc_verde_b200", "csrc", "build")
    a = parse(args[0] if args else here, model)
    b = parse(args[1], model) if len(args) > 1 else None
    for k in sorted(a):
        line = "%-70s regs %3d stack %4d spill %4d" % (k[:70], *a[k])
        if b is not None and k in b and b[k] != a[k]:
            line += "   ->  regs %3d stack %4d spill %4d" % b[k]
        print(line)
