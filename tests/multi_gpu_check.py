"""torchrun --nproc-per-node N tests/multi_gpu_check.py : multi-GPU parity check (run by tests/test_multi_gpu.py).
Distributed SpMV rows must be bit-identical to the single-GPU rows; distributed fused CG must
match the single-GPU fused CG (same iteration count +-1, x to 1e-10)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import workloads as W
from ginkgo_b200 import api
from ginkgo_b200 import distributed as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ex = api.HostExecutor(local)
dev = ex.device
ok = True
OVERLAP = os.environ.get("B200_DIST_OVERLAP", "0") == "1"


def rows_equal(y, yref):
    """bit-equal on the default paths; the pipelined exchange re-associates the row sums"""
    if not OVERLAP:
        return torch.equal(y, yref)
    return bool(((y - yref).abs().max() <= 1e-13 * yref.abs().max()).item())

for name, kw in [("lap3d_40", dict(grid=40, dims=3)), ("cfg2_small", dict(n=200000))]:
    with torch.cuda.stream(ex.stream):
        if name.startswith("lap"):
            rp, ci, va = W.laplace(kw["grid"], kw["dims"], xp="torch", device=dev)
            n = kw["grid"] ** kw["dims"]
        else:
            rp, ci, va = W.build("cfg2", xp="torch", device=dev, n=kw["n"])
            n = kw["n"]
        x = W.vector(n, xp="torch", device=dev)
        offs = D.uniform_offsets(n, world)
        r0, r1 = offs[rank], offs[rank + 1]
        p0, p1 = int(rp[r0]), int(rp[r1])
        lrp = (rp[r0:r1 + 1] - rp[r0]).contiguous()
        lci, lva = ci[p0:p1].contiguous(), va[p0:p1].contiguous()
        # single-GPU result (every rank computes it, for comparison)
        A1 = api.host_csr(ex, (n, n), va, ci, rp)
        y1 = torch.zeros(n, dtype=torch.float64, device=dev)
        _h = api._host()
        xd1, yd1 = api.host_dense(ex, x), api.host_dense(ex, y1)  # keep the handles alive
        api._hcheck(_h.gkob_apply(A1.h, xd1.h, yd1.h))
    A = api.DistMatrix(ex, offs, lrp, lci, lva)
    with torch.cuda.stream(ex.stream):
        x_ext = torch.zeros(A.n_local + A.n_ghost, dtype=torch.float64, device=dev)
        x_ext[:A.n_local] = x[r0:r1]
        y = torch.zeros(A.n_local, dtype=torch.float64, device=dev)
    A.apply(x_ext, y)
    ex.synchronize()
    same = rows_equal(y, y1[r0:r1])
    # repeated exchanges with changing data (epochs / double-buffered landing slots)
    for k in range(1, 6):
        with torch.cuda.stream(ex.stream):
            xk = W.vector(n, stream=20 + k, xp="torch", device=dev)
            x_ext[:A.n_local] = xk[r0:r1]
            xdk = api.host_dense(ex, xk)
        api._hcheck(_h.gkob_apply(A1.h, xdk.h, yd1.h))
        A.apply(x_ext, y)
        ex.synchronize()
        same = same and rows_equal(y, y1[r0:r1])
    ok &= same
    with torch.cuda.stream(ex.stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            A.apply(x_ext, y)
        dist.barrier()
        e0.record(ex.stream)
        for _ in range(100):
            A.apply(x_ext, y)
        e1.record(ex.stream)
    ex.synchronize()
    print("rank %d %s: n_local=%d n_ghost=%d p2p=%d pipelined=%s spmv %s=%s  %.1f us/apply"
          % (rank, name, A.n_local, A.n_ghost, A.p2p, A.pipelined, "1e-13-equal" if OVERLAP else "bit-equal", same,
             e0.elapsed_time(e1) * 10), flush=True)
    # the same matrix through read_distributed: split, renumbering and send lists built by the
    # library's device kernels + its own all-gather (no torch.distributed in the set-up)
    rp_h, ci_h, va_h = rp.cpu().numpy(), ci.cpu().numpy(), va.cpu().numpy()
    rows_h = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp_h))
    part = api.HostPartition.uniform(ex, world, n)
    pb = part.info()["range_bounds"]
    q0, q1 = int(pb[rank]), int(pb[rank + 1])
    mine = (rows_h >= q0) & (rows_h < q1)  # a rank only needs to pass its own rows
    A2 = api.DistMatrix.read(ex, part, (n, n), rows_h[mine], ci_h[mine], va_h[mine])
    with torch.cuda.stream(ex.stream):
        x2 = torch.zeros(A2.n_local + A2.n_ghost, dtype=torch.float64, device=dev)
        y2 = torch.zeros(A2.n_local, dtype=torch.float64, device=dev)
    same2 = A2.n_local == q1 - q0
    for k in range(3):
        with torch.cuda.stream(ex.stream):
            xk = W.vector(n, stream=40 + k, xp="torch", device=dev)
            x2[:A2.n_local] = xk[q0:q1]
            xdk = api.host_dense(ex, xk)
        api._hcheck(_h.gkob_apply(A1.h, xdk.h, yd1.h))
        A2.apply(x2, y2)
        ex.synchronize()
        same2 = same2 and rows_equal(y2, y1[q0:q1])
        same2 = same2 and torch.equal(A2.last_ghosts().cpu(), xk.cpu()[torch.from_numpy(A2.ghost_globals)])
    ok &= same2
    print("rank %d %s: read_distributed n_local=%d n_ghost=%d p2p=%d spmv bit-equal=%s"
          % (rank, name, A2.n_local, A2.n_ghost, A2.p2p, same2), flush=True)
    if name.startswith("lap"):
        # every host-layer solver on distributed::Matrix / distributed::Vector (dots and norms
        # all-reduced, Jacobi from the local block) against the same solver on one GPU
        bfull = W.vector(n, stream=77, xp="torch", device=dev)
        bd_full = api.host_dense(ex, bfull)
        for kind, pre in (("cg", 1), ("bicgstab", 0), ("gmres", 1), ("fcg", 0), ("cgs", 1), ("pipe_cg", 0),
                          ("minres", 0), ("gcr", 1)):
            s1 = api.HostSolver(ex, kind, A1, precond_max_bs=pre, max_iters=600, reduction=1e-9, fused=False,
                                krylov_dim=30)
            with torch.cuda.stream(ex.stream):
                xs = torch.zeros(n, dtype=torch.float64, device=dev)
                xl = torch.zeros(A2.n_local, dtype=torch.float64, device=dev)
                bl = bfull[q0:q1].contiguous()
            xsd = api.host_dense(ex, xs)
            s1.apply(bd_full, xsd)
            it, st = A2.solve(kind, bl, xl, n, precond_max_bs=pre, max_iters=600, reduction=1e-9, krylov_dim=30)
            ex.synchronize()
            err = (xl - xs[q0:q1]).norm().item() / max(xs[q0:q1].norm().item(), 1e-300)
            good = abs(it - s1.num_iterations) <= 2 and err < 1e-7
            ok &= good
            print("rank %d %s: distributed %s(precond %d) iters %d (1 GPU: %d) rel diff %.2e ok=%s"
                  % (rank, name, kind, pre, it, s1.num_iterations, err, good), flush=True)
    if name.startswith("lap"):
        b = torch.ones(n, dtype=torch.float64, device=dev)
        s1 = api.HostSolver(ex, "cg", A1, precond_max_bs=1, max_iters=2000, reduction=1e-9, fused=True)
        x1 = torch.zeros(n, dtype=torch.float64, device=dev)
        bd1, xd2 = api.host_dense(ex, b), api.host_dense(ex, x1)
        s1.apply(bd1, xd2)
        A.make_cg(scalar_jacobi=True, max_iters=2000, reduction=1e-9)
        xd = torch.zeros(A.n_local, dtype=torch.float64, device=dev)
        it, st = A.cg_apply(b[r0:r1].contiguous(), xd)
        ex.synchronize()
        err = (xd - x1[r0:r1]).norm().item() / x1[r0:r1].norm().item()
        good = abs(it - s1.num_iterations) <= 1 and st == s1.stop_status and err < 1e-9
        ok &= good
        print("rank %d %s: dist cg iters %d (1 GPU: %d) status %#x rel diff %.2e ok=%s"
              % (rank, name, it, s1.num_iterations, st, err, good), flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if t.item() == 1 else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
