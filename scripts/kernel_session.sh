#!/bin/bash
# Kernel iteration on one GPU (under gpurun): the scoring parity subset, base / wide / deep rates, phase trace of the fused kernel (needs scripts/trace_build.sh).  TAG=<name> bash scripts/kernel_session.sh
set -u
O=gpurun_out/${TAG:-r03i}; mkdir -p $O
SEL="scores_and_decisions or every_stage or fused_layer or other_network_shapes or host_buffers or empty_candidate or other_round or full_size"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for w in base wide deep; do bash scripts/sweep_opts.sh $w "fuse=1" 2>&1 | tee -a $O/sweep.log; done
GNNB_LIB=build/libgnnb_trace.so timeout 300 python bench.py --domains 512 --chunk 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-babsr --no-online --no-queue --no-secondary --no-step > $O/trace.log 2>&1
grep "TRACE fused-mma\|TRACE fused-chain" $O/trace.log | sed -n '19,24p'
