#!/bin/bash
# Same-box A/B of run-time knobs: scripts/ab_env.sh <repeats> "VAR=a VAR=b ..." (each word one environment setting; "-" = none)
rep=$1; shift
for i in $(seq $rep); do
  for e in "$@"; do
    if [ "$e" = "-" ]; then ms=$(python "$(dirname "$0")/../bench.py" --no-others --no-cpu-baseline --steps 20 2>&1 >/dev/null | grep "^rank 0" | head -1 | awk '{print $4}');
    else ms=$(env $e python "$(dirname "$0")/../bench.py" --no-others --no-cpu-baseline --steps 20 2>&1 >/dev/null | grep "^rank 0" | head -1 | awk '{print $4}'); fi
    echo "$e $ms"
  done
done
