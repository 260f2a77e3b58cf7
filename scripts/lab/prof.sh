#!/bin/bash
# ncu --set full capture of ONE lab case:  prof.sh <tag> <group> <matrix> "<label substring>"
TAG=$1; G=$2; M=$3; export LAB_ONLY="$4"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ring_kernel|warp_stream|warp_pipe' -s 3 -c 1 -f -o gpurun_out/prof_$TAG scripts/lab/spmv_lab $G $M > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
