#!/bin/bash
# one gpurun call: smoke, gpu tests, bench, experiments, ncu captures -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
echo "== exp"; timeout 600 python scripts/exp_spmv.py 2>&1 | tail -12
echo "== bench"; timeout 900 python bench.py --steps 200 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json | tail -2; tail -5 gpurun_out/bench.err
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tee gpurun_out/bench_ref.json | tail -1
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:slab_kernel -s 6 -c 2 -f -o gpurun_out/prof_csr_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
