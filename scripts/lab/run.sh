#!/bin/bash
# runs the experiment harness on the GPU box; output -> gpurun_out/lab_$TAG.txt
TAG=${1:-r02}; shift
GROUPS_="${@:-micro base ring tma4}"
L=scripts/lab/spmv_lab
O=gpurun_out/lab_$TAG.txt
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O
for g in $GROUPS_; do
  case $g in
    micro|tma4|l1cap) timeout 120 $L $g >> $O 2>&1; echo "$g rc=$?" >> $O;;
    *) for m in ${MATS:-cfg2 cfg2a cfg2h banded cfg3 cfg4}; do timeout 180 $L $g $m >> $O 2>&1; echo "rc=$?" >> $O; done;;
  esac
done
cat $O
