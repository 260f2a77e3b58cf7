#!/bin/bash
mkdir -p gpurun_out
echo "== pytest new"; timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_parity_gpu.py tests/test_solvers_gpu.py -q -m gpu -x --timeout 180 2>&1 | tail -4
echo "== exp tuned"; B200_DEBUG=1 timeout 300 python scripts/exp_spmv.py 2>&1 | grep -v torch_copy | tail -8
echo "== bench tuned"; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-gmres 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('tuned', d['value'], d['roofline']['kernel'], d['cg']['cfg3']['iters_per_s'], d['cg']['cfg5']['iters_per_s'])"
