#!/bin/bash
mkdir -p gpurun_out
echo "== pytest ring"; B200_CSR_KERNEL=ring timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_solvers_gpu.py tests/test_golden.py -q -m gpu -x --timeout 120 -k "csr or coo or solver or cg or Spmv" 2>&1 | tail -6
echo "== exp ring"; B200_CSR_KERNEL=ring B200_DEBUG=1 timeout 300 python scripts/exp_spmv.py 2>&1 | grep -v torch_copy | tail -7
echo "== exp warp"; timeout 300 python scripts/exp_spmv.py 2>&1 | grep -v torch_copy | tail -5
