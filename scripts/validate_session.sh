#!/bin/bash
# Full validation on one GPU (under gpurun): the whole GPU test suite, then the default bench line.  TAG=<name> bash scripts/validate_session.sh
set -u
export TAG=${TAG:-r03o}; O=gpurun_out/$TAG; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
tail -c 600 $O/bench_default.err
python - <<'PY'
import json, os
d=json.loads(open('gpurun_out/' + os.environ.get('TAG', 'r03o') + '/bench_default.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'wide',d['wide'].get('value'),d['wide'].get('e2e',{}).get('value'),'deep',d['deep'].get('value'),d['deep'].get('e2e',{}).get('value'))
print('step', d['frontier_step'].get('value'), d['frontier_step'].get('child_bounds', {}).get('value'), 'traffic', d['roofline'].get('traffic'), d['roofline']['bound'], round(d['roofline']['frac'], 3))
PY
