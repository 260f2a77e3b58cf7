#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (warp kernel default)"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -12
for v in warp tma slab; do
  echo "== exp $v"; B200_CSR_KERNEL=$v B200_DEBUG=1 timeout 600 python scripts/exp_spmv.py 2>&1 | grep -v torch_copy | tail -7
done
echo "== pytest csr with tma kernel"; B200_CSR_KERNEL=tma timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "csr or coo" 2>&1 | tail -3
echo "== ncu warp random+banded"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_stream -s 3 -c 1 -f -o gpurun_out/prof_warp_random python scripts/exp_spmv.py cfg2 > gpurun_out/ncu_wr.log 2>&1; tail -1 gpurun_out/ncu_wr.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_stream -s 3 -c 1 -f -o gpurun_out/prof_warp_banded python scripts/exp_spmv.py cfg2_banded > gpurun_out/ncu_wb.log 2>&1; tail -1 gpurun_out/ncu_wb.log
