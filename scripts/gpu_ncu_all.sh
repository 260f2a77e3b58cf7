#!/bin/bash
# one 1-GPU call: the evidence set of the round.
#  1. bench.py (real numbers)            2. launch list of the bench command (ncu --metrics gpu__time_duration.sum)
#  3. ncu --set full of the cfg2 SpMV launches of bench.py  -> roofline.traffic source
#  4. ncu --set full of ONE launch of every b200:: kernel of scripts/exp_kernels.py (6th invocation each)
#  5. per-kernel roofline table (CUDA events, unprofiled)     6. compute-sanitizer runs
# The .ncu-rep files stay on the box (/tmp): only their summaries come back (gpurun_out/ is capped at 64 MiB).
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest"; timeout 1500 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== dropin"; timeout 300 tests/dropin/_build/dropin_check cuda > gpurun_out/${TAG}_dropin_check.txt 2>&1; tail -1 gpurun_out/${TAG}_dropin_check.txt
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err; cut -c1-300 gpurun_out/${TAG}_bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>gpurun_out/${TAG}_bench_ref.err; cut -c1-400 gpurun_out/${TAG}_bench_ref.json
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'b200::' -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu --small-cg --no-twin > gpurun_out/${TAG}_ncu_launches.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_launches.log | cut -c1-200
echo "== ncu full: cfg2 SpMV of bench.py"; timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'b200::csr::(warp_stream|warp_pipe|ring)_kernel' -s 12 -c 2 -f -o /tmp/${TAG}_prof_spmv_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu --no-cg --no-twin > gpurun_out/${TAG}_ncu_cfg2.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_cfg2.log | cut -c1-200; python scripts/ncu_summary.py /tmp/${TAG}_prof_spmv_cfg2.ncu-rep gpurun_out/${TAG}_spmv_cfg2.json "the two launches of one cfg2 SpMV of bench.py (column-blocked copy), final build" | tail -1
echo "== ncu full: one launch of every kernel family"; timeout 1500 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'b200::' --kernel-id :::6 -f -o /tmp/${TAG}_prof_kernels python scripts/exp_kernels.py > gpurun_out/${TAG}_ncu_kernels.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_kernels.log | cut -c1-200; python scripts/ncu_summary.py /tmp/${TAG}_prof_kernels.ncu-rep gpurun_out/${TAG}_kernels_ncu.json "one launch (the 6th) of every b200:: kernel of scripts/exp_kernels.py, final build" | tail -1
echo "== kernel roofline table"; EXP_KERNELS_OUT=gpurun_out/${TAG}_kernels_roofline.json timeout 600 python scripts/exp_kernels.py > gpurun_out/${TAG}_exp_kernels.log 2>&1; tail -3 gpurun_out/${TAG}_exp_kernels.log | cut -c1-200
echo "== sanitizer"; bash scripts/gpu_sanitize.sh
ls -la gpurun_out | tail -20
