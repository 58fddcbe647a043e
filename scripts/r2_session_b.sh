#!/bin/bash
set -u
O=gpurun_out/r03b; mkdir -p $O
for a in base deep wide; do timeout 300 python tests/gpu_isolated.py child_bounds $a 2>&1 | tail -4; done | tee $O/child.log
for a in base deep; do timeout 300 python tests/gpu_isolated.py kw_bounds $a 2>&1 | tail -2; done | tee -a $O/child.log
timeout 300 python scripts/kw_probe.py base 256 2>&1 | tail -3 | tee $O/probe.log
timeout 300 python scripts/kw_probe.py deep 256 2>&1 | tail -3 | tee -a $O/probe.log
bash scripts/sweep_opts.sh base "fused_mma_group=1" "fused_mma_group=2" "fused_mma_group=3" 2>&1 | tee $O/sweep_group.log
