#!/bin/bash
# One GPU session: parity tests, bench lines for the three workloads, ncu launch list and full captures.
# Usage (under gpurun): bash scripts/gpu_session.sh <tag>
set -u
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
for w in base wide deep; do
  python bench.py --workload $w --steps 10 --warmup 3 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; echo "bench $w rc=$?"
done
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>&1
CMD="python bench.py --domains 256 --chunk 256 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-babsr --no-online --no-queue"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 33 -c 40 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launches.log 2>&1
$CMD > $O/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_tc_(update|prop|relax)" -s 26 -c 26 -o $O/${TAG}_prof -f $CMD > $O/${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/*_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('roofline',{}).get('kernel_ms'), d.get('clocks'))
    except Exception as e:
        print(f, 'ERR', e)
PY
