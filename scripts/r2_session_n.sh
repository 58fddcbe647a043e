#!/bin/bash
set -u
O=gpurun_out/r03n; mkdir -p $O
for a in base deep; do timeout 300 python tests/gpu_isolated.py frontier_step $a 2>&1 | tail -3; done | tee $O/step.log
for a in base deep wide; do timeout 300 python scripts/step_probe.py $a 512 2>&1 | tail -3; done | tee $O/probe.log
