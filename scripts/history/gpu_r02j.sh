#!/bin/bash
# round-2 session-2 first call: validate the tree (smoke, gpu tests, bench), kernel table, lab micro (L1 capacity),
# one --set full capture of the scattered-gather kernel, sanitizers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/r02j_gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest"; timeout 1200 python -m pytest tests -q -m gpu --timeout 180 -x 2>&1 | tail -12 | tee gpurun_out/r02j_pytest.txt
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; tail -2 gpurun_out/r02j_bench.err; cut -c1-1500 gpurun_out/r02j_bench.json
echo "== kernel table"; EXP_KERNELS_OUT=gpurun_out/r02j_kernels_roofline.json timeout 600 python scripts/exp_kernels.py > gpurun_out/r02j_exp_kernels.log 2>&1; grep -c frac gpurun_out/r02j_exp_kernels.log; tail -2 gpurun_out/r02j_exp_kernels.log | cut -c1-300
echo "== lab"; MATS="cfg2a cfg2" bash scripts/lab/run.sh r02j l1cap base > /dev/null 2>&1; grep -c "" gpurun_out/lab_r02j.txt
echo "== ncu full: warp_stream on cfg2a (lab)"; bash scripts/lab/prof.sh r02j_stream_cfg2a base cfg2a "warp_stream" | cut -c1-200
echo "== sanitizer"; bash scripts/gpu_sanitize.sh
ls -la gpurun_out | tail -20
