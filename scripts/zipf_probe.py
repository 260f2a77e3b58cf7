"""GPU probe (harness only): the Zipf twin of cfg2 through the host layer, a few SpMVs -- run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel breakdown of one skewed SpMV."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import workloads as W
from ginkgo_b200 import api

hx = api.HostExecutor(0)
dev = hx.device
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2_zipf"
with torch.cuda.stream(hx.stream):
    rp, ci, va = W.build(cfg, xp="torch", device=dev)
    n = rp.numel() - 1
    x = W.vector(n, xp="torch", device=dev)
    y = torch.zeros(n, dtype=torch.float64, device=dev)
A = api.host_csr(hx, (n, n), va, ci, rp)
xd, yd = api.host_dense(hx, x), api.host_dense(hx, y)
h = api._host()
for _ in range(6):
    api._hcheck(h.gkob_apply(A.h, xd.h, yd.h))
hx.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(hx.stream)
for _ in range(20):
    api._hcheck(h.gkob_apply(A.h, xd.h, yd.h))
e1.record(hx.stream)
torch.cuda.synchronize()
print("%s: %.4f ms per SpMV, variant %d, parts %d" % (cfg, e0.elapsed_time(e1) / 20, h.gkob_csr_kernel_variant(A.h),
                                                      max(1, h.gkob_csr_plan_parts(A.h))))
