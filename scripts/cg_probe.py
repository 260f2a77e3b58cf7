"""cfg3 fused CG, 100 iterations (harness for ncu captures of the CG kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workloads as W
from ginkgo_b200 import api
ex = api.HostExecutor(0)
dev = ex.device
g = 200
n = g ** 3
with torch.cuda.stream(ex.stream):
    rp, ci, va = W.laplace(g, 3, xp="torch", device=dev)
    b = torch.ones(n, dtype=torch.float64, device=dev)
    x = torch.zeros(n, dtype=torch.float64, device=dev)
A = api.host_csr(ex, (n, n), va, ci, rp)
s = api.HostSolver(ex, "cg", A, precond_max_bs=1, max_iters=120, reduction=1e-30, fused=True, check_every=20)
bd, xd = api.host_dense(ex, b), api.host_dense(ex, x)
s.apply(bd, xd)
ex.synchronize()
print("iterations", s.num_iterations)
