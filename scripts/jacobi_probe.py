"""GPU probe (harness only): block-Jacobi simple_apply on cfg4's block shape (250k x 16x16 fp32), full
precision vs adaptive storage (gko::half), CUDA-event timings; run under ncu for the two kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ginkgo_b200 import api

ex = api.B200Executor.create(0)
dev = ex.device
nb, bs, gp = 250_000, 16, 1
n = nb * bs
with torch.cuda.stream(ex.stream):
    bp = torch.arange(0, n + 1, bs, dtype=torch.int32, device=dev)
    blocks = torch.rand(nb * bs * bs, dtype=torch.float32, device=dev)
    b = torch.rand(n, dtype=torch.float32, device=dev)
    x = torch.zeros(n, dtype=torch.float32, device=dev)
    half_blocks = torch.zeros(nb * bs * bs, dtype=torch.float32, device=dev)
    # every group holds 2 blocks x 256 halfs in the first half of its 512-float slot
    hb = half_blocks.view(torch.float16).view(nb // 2, 1024)
    hb[:, :512] = torch.rand(nb // 2, 512, dtype=torch.float32, device=dev).to(torch.float16)
    prec = {k: torch.full((nb,), v, dtype=torch.uint8, device=dev) for k, v in
            (("full (0,0)", 0x00), ("half (0,2)", 0x02), ("truncated<float,2> (1,0)", 0x10))}


def timeit(fn, reps=30):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ex.stream)
    for _ in range(reps):
        fn()
    e1.record(ex.stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timeit(lambda: ex.run("b200_jacobi_simple_apply_f32_i32", nb, bs, bs, bs * 2 * bs, gp, bp, blocks, b, 1, 1, x, 1))
print("plain entry point, full precision: %.4f ms  %.0f GB/s" % (ms, (nb * bs * bs * 4 + 2 * n * 4) / ms / 1e6))
for name, p in prec.items():
    src = blocks if name.startswith("full") else half_blocks
    w = 4 if name.startswith("full") else 2
    ms = timeit(lambda: ex.run("b200_jacobi_simple_apply_adaptive_f32_i32", nb, bs, bs, bs * 2 * bs, gp, p, bp, src, b,
                               1, 1, x, 1))
    print("adaptive entry point, %s: %.4f ms  %.0f GB/s" % (name, ms, (nb * bs * bs * w + 2 * n * 4) / ms / 1e6))
