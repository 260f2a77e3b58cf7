#!/bin/bash
mkdir -p gpurun_out
echo "== zipf / cfg2"; timeout 300 python scripts/zipf_probe.py cfg2_zipf 2>&1 | tail -1; timeout 300 python scripts/zipf_probe.py cfg2 2>&1 | tail -1
echo "== pytest csr"; timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_solvers_gpu.py -q -m gpu --timeout 180 2>&1 | tail -3
echo "== sanitizer"; bash scripts/gpu_sanitize.sh 2>&1 | grep -E "==|SUMMARY"
