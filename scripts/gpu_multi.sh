#!/bin/bash
# multi-GPU call: parity check + bench at $1 GPUs
N=$1
mkdir -p gpurun_out
echo "== dist check ($N GPUs)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | grep -E "rank|DIST_CHECK|Error|error" | tail -20
echo "== bench $N GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 5 2>gpurun_out/bench$N.err | tee gpurun_out/bench$N.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','roofline','e2e','cg','gpu_launches')}, indent=1)[:2500])"; grep -v "^W\|OMP_NUM\|\*\*\*\*" gpurun_out/bench$N.err | tail -5
