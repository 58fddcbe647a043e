#!/bin/bash
# experiment session: early accumulator release + turn taking of the two subdomain chains (GNNB_FUSED_PINGPONG)
set -u
O=gpurun_out/r03a; mkdir -p $O
SEL="scores_and_decisions or every_stage or fused_layer or other_network_shapes or host_buffers or empty_candidate"
for pp in 0 1; do
  GNNB_FUSED_PINGPONG=$pp timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $O/pytest_pp$pp.log 2>&1; echo "pytest pp=$pp rc=$?"; tail -2 $O/pytest_pp$pp.log
done
for w in base wide deep; do
  bash scripts/sweep_opts.sh $w "fused_pingpong=0" "fused_pingpong=1" 2>&1 | tee -a $O/sweep.log
done
for pp in 0 1; do
GNNB_LIB=build/libgnnb_trace.so GNNB_FUSED_PINGPONG=$pp timeout 300 python bench.py --domains 512 --chunk 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-babsr --no-online --no-queue --no-secondary > $O/trace_pp$pp.log 2>&1
grep -c TRACE $O/trace_pp$pp.log
done
