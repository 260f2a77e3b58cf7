#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -q -m gpu -x --timeout 120 2>&1 | tail -15
echo "== exp warp"; B200_DEBUG=1 timeout 300 python scripts/exp_spmv.py 2>&1 | grep -v torch_copy | tail -7
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(json.dumps({k:d[k] for k in ('value','ms_per_step','roofline','e2e','cg','gpu_launches','clocks')}, indent=1)[:3000])"; tail -5 gpurun_out/bench.err
