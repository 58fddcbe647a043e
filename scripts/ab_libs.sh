#!/bin/bash
# A/B of two builds of libgnnb.so on the same box: bash scripts/ab_libs.sh <workload> <libA> <libB> [reps]
W=$1; A=$2; B=$3; R=${4:-3}
for r in $(seq $R); do
  for L in $A $B; do
    GNNB_LIB=$L timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-babsr --no-online --no-queue --no-secondary --no-step 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W', '$L', round(d['value']))"
  done
done
