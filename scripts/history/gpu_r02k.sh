#!/bin/bash
# r02k: adaptive Jacobi on the device, unrolled long rows (Zipf), drop-in check, sanitizers with the ring kernel
mkdir -p gpurun_out
echo "== pytest new"; timeout 900 python -m pytest tests/test_jacobi_adaptive_gpu.py -q -m gpu --timeout 180 -x 2>&1 | tail -8 | tee gpurun_out/r02k_pytest_jacobi.txt
echo "== pytest all"; timeout 1200 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -8 | tee gpurun_out/r02k_pytest.txt
echo "== dropin"; timeout 300 tests/dropin/_build/dropin_check cuda > gpurun_out/r02k_dropin_check.txt 2>&1; tail -22 gpurun_out/r02k_dropin_check.txt
echo "== kernel table"; EXP_KERNELS_OUT=gpurun_out/r02k_kernels_roofline.json timeout 600 python scripts/exp_kernels.py > gpurun_out/r02k_exp_kernels.log 2>&1; grep -E "zipf|generate|cfg2 " gpurun_out/r02k_exp_kernels.log | cut -c1-260
echo "== sanitizer"; bash scripts/gpu_sanitize.sh
