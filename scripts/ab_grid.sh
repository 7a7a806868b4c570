#!/bin/bash
# (batches in flight) x (pipes per batch) grid on one box: scripts/ab_grid.sh <repeats> "F:P" ...
rep=$1; shift
for i in $(seq $rep); do
  for fp in "$@"; do
    f=${fp%%:*}; p=${fp##*:}
    ms=$(MPCV_PHASE_PIPES=$p python "$(dirname "$0")/../bench.py" --inflight $f --no-others --no-cpu-baseline --steps 24 2>&1 >/dev/null | grep "^rank 0" | head -1 | awk '{print $4}')
    echo "inflight=$f pipes=$p $ms"
  done
done
