#!/bin/bash
# one 1-GPU gpurun call: smoke, gpu tests, bench (+reference arm), ncu launch list, ncu full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/host.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest"; timeout 900 python -m pytest tests -q -m gpu --timeout 120 2>&1 | tail -8
echo "== bench"; timeout 900 python bench.py --steps 200 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','roofline','e2e','cg','gpu_launches','clocks','cpu_baseline')}, indent=1)[:3500])"; tail -3 gpurun_out/bench.err
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tee gpurun_out/bench_ref.json | cut -c1-400
echo "== ncu launches (bench, no cpu leg, small cg; tuning decisions replayed from an unprofiled run)"
rm -f gpurun_out/tune.log
B200_TUNE_RECORD=gpurun_out/tune.log timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --small-cg > gpurun_out/bench_small.json 2>/dev/null
B200_TUNE_REPLAY=gpurun_out/tune.log timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'warp_stream|warp_pipe|slab|step_|plan_kernel|ew_kernel|reduce_c|init_scalars|extract_diag|multi_|hessenberg|solve_krylov|block_apply|residual_norm|pack_kernel' -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --small-cg > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu full: spmv cfg2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'warp_stream|warp_pipe' -s 22 -c 2 -f -o gpurun_out/prof_spmv_cfg2 python bench.py --steps 10 --warmup 3 --no-cpu --no-cg > gpurun_out/ncu_full1.log 2>&1; tail -1 gpurun_out/ncu_full1.log | cut -c1-200
echo "== ncu full: CG kernels (cfg3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'warp_stream|warp_pipe|step_xr|step_p' -s 338 -c 3 -f -o gpurun_out/prof_cg_cfg3 python scripts/cg_probe.py > gpurun_out/ncu_full2.log 2>&1; tail -2 gpurun_out/ncu_full2.log | cut -c1-200
echo "== exp kernels"; timeout 300 python scripts/exp_kernels.py > gpurun_out/exp_kernels.log 2>&1; tail -3 gpurun_out/exp_kernels.log | cut -c1-200
ls -la gpurun_out | tail -24
