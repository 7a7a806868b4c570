#!/bin/bash
# Same-box A/B of library builds: scripts/ab.sh <repeats> <bench args or -> <tag> [<tag> ...]   (tag "" = libmpcv.so,
# tag x = libmpcv_x.so).  Runs are interleaved and each is its own process: the slab lands on different physical pages
# from process to process, which alone moves a C2 batch by +-1 ms on some boxes — compare the minima / medians.
rep=$1; shift
args=$1; shift
[ "$args" = "-" ] && args="--no-others --no-cpu-baseline --steps 20"
L="$(cd "$(dirname "$0")/.." && pwd)/mpc_verde_b200"
for i in $(seq $rep); do
  for t in "$@"; do
    lib=$L/libmpcv.so; [ "$t" != "base" ] && lib=$L/libmpcv_$t.so
    ms=$(MPCV_LIB=$lib python "$(dirname "$0")/../bench.py" $args 2>&1 >/dev/null | grep "^rank 0" | head -1 | awk '{print $4}')
    echo "$t $ms"
  done
done
