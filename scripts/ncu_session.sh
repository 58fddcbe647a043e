#!/bin/bash
# ncu captures of one step of the default workload (run under gpurun, one GPU): launch list + full sections of the hot kernels
set -u
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --domains 512 --chunk 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-babsr --no-online --no-queue"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_(tc|amb|output|property|masked)" -s 42 -c 42 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launches.log 2>&1
$CMD > $O/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_tc_(update|prop|relax|input)" -s 28 -c 28 -o $O/${TAG}_prof -f $CMD > $O/${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"
