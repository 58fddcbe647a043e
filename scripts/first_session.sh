#!/bin/bash
# First GPU session after code was written without a GPU (end of round 1): validate it, then measure.  Run under gpurun, one GPU:
#   gpurun --timeout 1500 -- 'bash scripts/first_session.sh r02a'
set -u
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
# 1. the whole GPU suite (the isolated checks run in their own processes)
python -m pytest tests -m gpu -q -rxX > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/${TAG}_pytest.log
# 2. the isolated checks again, verbosely, so that a failure shows its assertion
for a in base deep wide; do
  for c in gather_prefetch kw_bounds; do
    timeout 600 python tests/gpu_isolated.py $c $a > $O/${TAG}_iso_${c}_$a.log 2>&1; echo "$c $a rc=$? $(tail -1 $O/${TAG}_iso_${c}_$a.log | cut -c1-200)"
  done
done
# 3. the propagation variant with prefetched gather indices against the default
for pf in 0 1 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-babsr --no-online --no-queue --opt gather_prefetch=$pf \
      > $O/${TAG}_bench_base_pf$pf.json 2> $O/${TAG}_bench_base_pf$pf.err; echo "bench pf=$pf rc=$?"
done
# 4. the full default bench lines (with every secondary object and both baselines)
for w in base wide deep; do
  python bench.py --workload $w --steps 10 --warmup 3 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; echo "bench $w rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'], 3), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('kernel_ms'),
              (d.get('cpu_baseline') or {}).get('eager_pytorch_on_gpu'))
    except Exception as e:
        print(f, 'ERR', e)
PY
