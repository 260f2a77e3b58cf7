#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -q -m gpu --timeout 180 2>&1 | tail -6
echo "== bench (no cpu leg)"; timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu 2>gpurun_out/bench_quick.err | tee gpurun_out/bench_quick.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['value'], {k:(round(v['iters_per_s'],1), v.get('first_apply_incl_generate_s')) for k,v in d['cg'].items()})"; tail -2 gpurun_out/bench_quick.err
