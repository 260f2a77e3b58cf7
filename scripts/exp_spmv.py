"""GPU experiment: SpMV time / achieved GB/s over the BASELINE matrices (harness only)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads as W
from ginkgo_b200.api import B200Executor, Csr, Dense

ex = B200Executor.create(0)
out = []


def bench(name, rp, ci, va, ncols, reps=50):
    n = rp.numel() - 1
    A = Csr(ex, (n, ncols), va, ci, rp)
    A.plan()
    with torch.cuda.stream(ex.stream):
        x = Dense(ex, W.vector(ncols, xp="torch", device=ex.device).to(va.dtype).reshape(-1, 1))
        y = Dense.create(ex, (n, 1), dtype=va.dtype)
    for _ in range(5):
        A.apply(x, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ex.stream)
    for _ in range(reps):
        A.apply(x, y)
    e1.record(ex.stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    vb = va.element_size()
    nbytes = W.spmv_bytes(n, ncols, va.numel(), vb, 4)
    r = dict(name=name, n=n, nnz=va.numel(), ms=ms, gbs=nbytes / ms / 1e6,
             gflops=2 * va.numel() / ms / 1e6)
    print(json.dumps(r), flush=True)
    out.append(r)


with torch.cuda.stream(ex.stream):
    # stream copy reference point
    a = torch.empty(1 << 28, dtype=torch.float64, device=ex.device)
    b = torch.empty_like(a)
    for _ in range(3):
        b.copy_(a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ex.stream)
    for _ in range(10):
        b.copy_(a)
    e1.record(ex.stream)
    torch.cuda.synchronize()
    print(json.dumps(dict(name="torch_copy_4GiB", gbs=2 * a.numel() * 8 / (e0.elapsed_time(e1) / 10) / 1e6)))
    del a, b

for cfg in sys.argv[1:] or ["cfg1", "cfg2", "cfg2_banded", "cfg3", "cfg4"]:
    with torch.cuda.stream(ex.stream):
        rp, ci, va = W.build(cfg, xp="torch", device=ex.device)
    bench(cfg, rp, ci, va, rp.numel() - 1)
    del rp, ci, va
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/exp_spmv.json", "w"))
