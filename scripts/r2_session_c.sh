#!/bin/bash
set -u
O=gpurun_out/r03c; mkdir -p $O
for a in base deep wide; do timeout 300 python tests/gpu_isolated.py child_bounds $a 2>&1 | tail -4; done | tee $O/child.log
for a in base deep; do timeout 300 python tests/gpu_isolated.py frontier_step $a 2>&1 | tail -6; done | tee $O/step.log
timeout 300 python scripts/step_probe.py base 256 2>&1 | tail -4 | tee $O/probe.log
