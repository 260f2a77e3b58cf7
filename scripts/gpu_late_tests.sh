#!/bin/bash
# First GPU call of the next round (1 GPU): the tests that were written after round 1's GPU budget
# was spent, file by file so that one failure does not hide the others, then the kernel table.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_late_tests.sh'
mkdir -p gpurun_out
export B200_LATE_GPU_STRICT=1  # the late tests as ordinary gpu tests
for f in tests/test_zz_late_gpu.py tests/test_zzz_dist_assembly_gpu.py; do
    echo "== $f"
    timeout 900 python -m pytest "$f" -q -m gpu 2>&1 | tail -25 | tee "gpurun_out/$(basename "$f" .py).log"
done
echo "== full gpu suite"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
