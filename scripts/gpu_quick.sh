#!/bin/bash
mkdir -p gpurun_out
echo "== pytest pipe"; B200_CSR_KERNEL=pipe timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_solvers_gpu.py tests/test_golden.py -q -m gpu -x --timeout 120 -k "csr or coo or solver or cg or Spmv or hybrid" 2>&1 | tail -4
echo "== exp pipe"; B200_CSR_KERNEL=pipe timeout 300 python scripts/exp_spmv.py 2>&1 | grep -v torch_copy | tail -5
echo "== exp warp"; timeout 300 python scripts/exp_spmv.py 2>&1 | grep -v torch_copy | tail -5
echo "== cg probe pipe"; B200_CSR_KERNEL=pipe timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-gmres 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('pipe', d['value'], d['cg']['cfg3']['iters_per_s'], d['cg']['cfg5']['iters_per_s'])"
