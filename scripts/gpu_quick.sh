#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -6
echo "== exp"; B200_DEBUG=1 timeout 300 python scripts/exp_spmv.py 2>&1 | grep -v torch_copy | tail -12
echo "== bench (no cpu leg)"; B200_DEBUG=1 timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu 2>gpurun_out/bench_quick.err | tee gpurun_out/bench_quick.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['e2e']['value'], {k:(round(v['iters_per_s'],1)) for k,v in d['cg'].items()})"; grep "b200" gpurun_out/bench_quick.err | tail -8
echo "== ncu full: spmv cfg2 (2 launches)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'warp_stream|warp_pipe' -s 22 -c 2 -f -o gpurun_out/prof_spmv_cfg2_parts python bench.py --steps 10 --warmup 3 --no-cpu --no-cg > gpurun_out/ncu_full3.log 2>&1; tail -1 gpurun_out/ncu_full3.log | cut -c1-200
