#!/bin/bash
# launch list of the bench command with the tuning decisions of an unprofiled run
mkdir -p gpurun_out
rm -f gpurun_out/tune.log
B200_TUNE_RECORD=gpurun_out/tune.log timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --small-cg > gpurun_out/bench_small.json 2>/dev/null
cat gpurun_out/tune.log
python -c "
import json; d=json.loads(open('gpurun_out/bench_small.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['kernel'][:80], d['gpu_launches'])"
B200_TUNE_REPLAY=gpurun_out/tune.log timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'warp_stream|warp_pipe|slab|step_|plan_kernel|ew_kernel|reduce_c|init_scalars|extract_diag|multi_|hessenberg|solve_krylov|block_apply|residual_norm|pack_kernel|generate_kernel|split_|unsorted' -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --small-cg > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-300
