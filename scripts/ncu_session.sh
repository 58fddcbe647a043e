#!/bin/bash
# ncu captures of one step of the default workload (run under gpurun, one GPU): launch list + full sections of the hot kernels.
# A base step of the tensor-core path is 23 launches: 3 (ambiguous-row compaction) + 1 (relax) + 1 (input embedding) + 2 rounds x
# [3 fused forward layers + output node + property back-propagation + update of the last layer + 2 fused backward layers] + 1 fused
# input-layer update + 1 masked argmax.  The second step of the command is captured.
set -u
TAG=${1:-r02}
N=${2:-23}
O=gpurun_out
mkdir -p $O
CMD="python bench.py --domains 512 --chunk 512 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-babsr --no-online --no-queue --no-secondary"
$CMD > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_(tc|amb|output|property|masked)" -s $N -c $N --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_launches.log 2>&1
$CMD > $O/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_tc_(fused|update|relax|input)" -s 16 -c 16 -o $O/${TAG}_prof -f $CMD > $O/${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"
# the bound producer: two launches of the windowed KW kernel (conv layers) of gnnb_child_bounds
KW="python scripts/kw_probe.py base 256"
$KW > $O/${TAG}_kw_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_kw_cone" -s 2 -c 2 -o $O/${TAG}_kw_prof -f $KW > $O/${TAG}_kw_ncu.log 2>&1
echo "ncu kw rc=$?"
