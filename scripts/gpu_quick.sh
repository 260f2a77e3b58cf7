#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
echo "== exp"; B200_DEBUG=1 timeout 600 python scripts/exp_spmv.py 2>&1 | tail -12
echo "== exp persist 0.6 (cfg2 only)"; B200_L2_PERSIST=0.6 timeout 300 python scripts/exp_spmv.py cfg2 2>&1 | tail -3
echo "== exp persist 1.0 (cfg2 only)"; B200_L2_PERSIST=1.0 timeout 300 python scripts/exp_spmv.py cfg2 2>&1 | tail -3
echo "== ncu banded"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slab_tma -s 3 -c 1 -f -o gpurun_out/prof_banded python scripts/exp_spmv.py cfg2_banded > gpurun_out/ncu_banded.log 2>&1; tail -2 gpurun_out/ncu_banded.log
echo "== ncu random"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slab_tma -s 3 -c 1 -f -o gpurun_out/prof_random python scripts/exp_spmv.py cfg2 > gpurun_out/ncu_random.log 2>&1; tail -2 gpurun_out/ncu_random.log
