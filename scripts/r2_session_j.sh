#!/bin/bash
set -u
O=gpurun_out/r03j; mkdir -p $O
for w in wide deep; do bash scripts/sweep_opts.sh $w "chunk=1024" "chunk=2048" "chunk=4096" 2>&1 | tee -a $O/sweep.log; done
timeout 300 python bench.py --domains 8192 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-babsr --no-online --no-queue --no-secondary --no-step --opt chunk=1024 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base8192 chunk1024', round(d['value']))" | tee -a $O/sweep.log
timeout 300 python bench.py --domains 8192 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-babsr --no-online --no-queue --no-secondary --no-step --opt chunk=4096 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('base8192 chunk4096', round(d['value']))" | tee -a $O/sweep.log
