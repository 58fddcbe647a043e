#!/bin/bash
set -u
O=gpurun_out/r03m; mkdir -p $O
for a in base deep wide; do timeout 300 python tests/gpu_isolated.py child_bounds $a 2>&1 | tail -2; done | tee $O/child.log
for a in odd_shapes deep_narrow; do timeout 300 python tests/gpu_isolated.py child_bounds_shapes $a 2>&1 | tail -2; done | tee -a $O/child.log
for a in base deep wide; do timeout 300 python scripts/kw_probe.py $a 512 2>&1 | tail -2; done | tee $O/probe.log
for a in base deep wide; do timeout 300 python scripts/step_probe.py $a 512 2>&1 | tail -1; done | tee -a $O/probe.log
