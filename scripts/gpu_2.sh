#!/bin/bash
mkdir -p gpurun_out
echo "== solver tests"; timeout 600 python -m pytest tests/test_solvers_gpu.py -q -m gpu --timeout 120 2>&1 | tail -8
echo "== dist check (2 GPUs)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep -v "^W\|warn" | tail -12
echo "== bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 3 2>gpurun_out/bench2.err | tee gpurun_out/bench2.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','roofline','e2e','cg','gpu_launches')}, indent=1)[:2500])"; tail -8 gpurun_out/bench2.err
echo "== exp warp 2ctas"; B200_DEBUG=1 timeout 300 python scripts/exp_spmv.py 2>&1 | grep -v torch_copy | tail -7
echo "== exp warp persist"; B200_L2_PERSIST=1.0 timeout 300 python scripts/exp_spmv.py cfg2 cfg4 2>&1 | grep -v torch_copy | tail -4
