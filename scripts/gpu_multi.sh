#!/bin/bash
# multi-GPU call: parity check (peer-memory path and NCCL path) + bench at $1 GPUs
N=$1
mkdir -p gpurun_out
for P2P in 1 0; do
echo "== dist check ($N GPUs, B200_P2P=$P2P)"
B200_P2P=$P2P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$P2P tests/multi_gpu_check.py 2>&1 | grep -E "rank|DIST_CHECK|Error|error|gko_b200" | sort | tail -24
done
for P2P in 1 0; do
echo "== bench $N GPUs B200_P2P=$P2P"
B200_P2P=$P2P timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$P2P bench.py --gpus $N --steps 50 --warmup 5 --no-gmres 2>gpurun_out/bench${N}_p2p$P2P.err | tee gpurun_out/bench${N}_p2p$P2P.json | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','roofline','cg','gpu_launches')})[:1500])"; grep -v "^W\|OMP_NUM\|\*\*\*\*" gpurun_out/bench${N}_p2p$P2P.err | tail -5
done
