#!/bin/bash
set -u
O=gpurun_out/r03l; mkdir -p $O
for m in 0 1; do for a in base deep wide; do GNNB_TEST_MATH=$m timeout 300 python tests/gpu_isolated.py child_bounds $a 2>&1 | tail -3; done; done | tee $O/child.log
for a in odd_shapes deep_narrow; do timeout 300 python tests/gpu_isolated.py child_bounds_shapes $a 2>&1 | tail -3; done | tee -a $O/child.log
for a in base deep; do timeout 300 python tests/gpu_isolated.py kw_bounds $a 2>&1 | tail -3; done | tee -a $O/child.log
for a in base deep; do timeout 300 python tests/gpu_isolated.py frontier_step $a 2>&1 | tail -2; done | tee $O/step.log
for a in base deep wide; do timeout 300 python scripts/kw_probe.py $a 512 2>&1 | tail -2; done | tee $O/probe.log
timeout 300 python scripts/step_probe.py base 1024 2>&1 | tail -2 | tee -a $O/probe.log
