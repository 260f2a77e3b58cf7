#!/bin/bash
mkdir -p gpurun_out
echo "== zipf / cfg2"; timeout 300 python scripts/zipf_probe.py cfg2_zipf 2>&1 | tail -1; timeout 300 python scripts/zipf_probe.py cfg2 2>&1 | tail -1
echo "== jacobi probe"; timeout 300 python scripts/jacobi_probe.py 2>&1 | tail -4 | tee gpurun_out/r02q_jacobi_probe.txt
echo "== pytest"; timeout 1500 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -3 | tee gpurun_out/r02q_pytest.txt
echo "== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | grep -E "==|SUMMARY"
echo "== dropin"; timeout 300 tests/dropin/_build/dropin_check cuda 2>&1 | tail -2
