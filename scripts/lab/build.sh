#!/bin/bash
# builds the experiment harness (not product): scripts/lab/spmv_lab
set -e
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false --extended-lambda \
  -Xcudafe --diag_suppress=177 ${PTXAS_V:+-Xptxas -v} scripts/lab/spmv_lab.cu -o scripts/lab/spmv_lab \
  -Lginkgo_b200/lib -lginkgo_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../ginkgo_b200/lib'
