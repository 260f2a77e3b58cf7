#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native SpMV + Krylov hot path.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one CSR SpMV  y = A x  of BASELINE.json configs[1]: random CSR fp64,
n = 10 000 000, nnz = 150 000 000 (15 distinct uniform columns per row), synthetic data from
workloads.py.  N > 1 (launched by torchrun, one rank per GPU): the matrix is split 1-D by
row ranges (strong scaling), every step all-gathers the x slices over NCCL and runs the local
SpMV; value = 2*nnz_total / max-over-ranks step time.

Printed JSON (one line, rank 0): metric / value / unit as BASELINE.json's metric, plus
  roofline     dominant kernel's algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline the reference's own OMP executor (oracle/_ref) on this box's host cores,
               bounded sample of the same workload
  e2e          same metric through gko_b200::staged_apply with HOST buffers (H2D x, SpMV,
               D2H y every step; consecutive steps overlap on the copy streams)
  cg           CG iterations/s (cfg3 on 1 GPU, cfg5-style row-sharded on N GPUs)
`--impl reference` times the reference's CPU implementation only (no GPU work).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import workloads as W  # noqa: E402

CFG = "cfg2"
N_ROWS = W.CONFIGS[CFG]["n"]
PER_ROW = W.CONFIGS[CFG]["per_row"]
CPU_SAMPLE_ROWS = 2_000_000  # bounded CPU sample: first 2M rows (30M nnz), full-width x


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """polls NVML (SM clock + clock-event reasons) every few ms during the timed region"""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, dev):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(dev)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop and self.nv:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------ CPU arm (reference OMP)
def cpu_reference_spmv(sample_rows, reps):
    """the reference's OMP executor (oracle/_ref) or, if that was not built, the oracle port"""
    rp, ci, va = W.build(CFG, 0, sample_rows, xp="np")
    x = W.vector(N_ROWS)
    nnz = len(va)
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    from oracle import ref
    if ref.available():
        cores = ref.use_physical_cores()
        _, sec = ref.spmv("csr", rp, ci, va, x, N_ROWS, exec_kind=1, reps=reps)
        kind = "reference"
    else:
        from oracle import oracle
        y = np.zeros(sample_rows)
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle.call("orc_csr_spmv_f64_i32", sample_rows, N_ROWS, nnz, rp, ci, va, x, 1, 1, y, 1)
        sec = (time.perf_counter() - t0) / reps
        cores, kind = 1, "port"
    return {"value": 2.0 * nnz / sec / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": kind,
            "sample": "first %d rows (nnz=%d) of %s against the full %d-entry x; %s; %d reps, "
                      "%.3f ms/SpMV" % (sample_rows, nnz, CFG, N_ROWS,
                                        "gko::OmpExecutor Csr(classical)::apply" if kind == "reference"
                                        else "oracle C port, 1 thread", reps, sec * 1e3),
            "ms": sec * 1e3, "nnz": nnz}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    reps = max(1, args.steps)
    # keep the whole run within minutes whatever K is: cap the timed repetitions
    reps = min(reps, 200)
    b = cpu_reference_spmv(CPU_SAMPLE_ROWS, reps)
    line = {
        "impl": "reference", "metric": "csr_spmv_fp64_gflops", "value": b["value"],
        "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": b["ms"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: random CSR n=%d nnz=%d (15/row) fp64/int32; CPU sample = %s"
                               % (CFG, N_ROWS, N_ROWS * PER_ROW, b["sample"])},
        "cpu_baseline": {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": b["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args, rank, world):
    import torch
    import torch.distributed as dist
    from ginkgo_b200 import api
    from ginkgo_b200 import distributed as D

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ex = api.HostExecutor(local)  # C++ host layer (gko_b200.hpp) over the C ABI
    dev = ex.device
    hl = api._host()

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(ex.stream)
        for _ in range(n):
            fn()
        e1.record(ex.stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---------------------------------------------------------------- SpMV, cfg2 (headline)
    offs = D.uniform_offsets(N_ROWS, world)
    r0, r1 = offs[rank], offs[rank + 1]
    with torch.cuda.stream(ex.stream):
        rp, ci, va = W.build(CFG, r0, r1, xp="torch", device=dev)
        x_full = W.vector(N_ROWS, xp="torch", device=dev)
    nnz_loc = va.numel()
    nnz_total = N_ROWS * PER_ROW
    if world == 1:
        A = api.host_csr(ex, (N_ROWS, N_ROWS), va, ci, rp)
        xg = api.host_dense(ex, x_full)
        y_t = torch.empty(N_ROWS, dtype=torch.float64, device=dev)
        yg = api.host_dense(ex, y_t)

        def step():
            api._hcheck(hl.gkob_apply(A.h, xg.h, yg.h))
        kernel_step = step
    else:
        def build_dist():
            A_ = api.DistMatrix(ex, offs, rp, ci, va)
            with torch.cuda.stream(ex.stream):
                xe = torch.zeros(A_.n_local + A_.n_ghost, dtype=torch.float64, device=dev)
                xe[:A_.n_local] = x_full[r0:r1]
                yt = torch.empty(A_.n_local, dtype=torch.float64, device=dev)
            Al = api.host_csr(ex, (A_.n_local, A_.n_local + A_.n_ghost), va, A_.col_idxs, rp)
            return A_, xe, yt, Al, api.host_dense(ex, xe), api.host_dense(ex, yt)

        def halo_delivers(A_, xe, yt):
            """one exchange: every ghost slot must hold the x entry of its global column (all ranks)"""
            A_.apply(xe, yt)
            ex.synchronize()
            with torch.cuda.stream(ex.stream):
                good = A_.n_ghost == 0 or torch.equal(xe[A_.n_local:], x_full[A_.part["ghosts"].to(dev).long()])
                t = torch.tensor([1 if good else 0], device=dev)
            ex.synchronize()
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return int(t.item()) == 1

        A, x_ext, y_t, Aloc, xe_h, y_h = build_dist()
        halo_path = "peer-memory halo exchange over NVLink" if A.p2p & 2 else "NCCL halo exchange"
        if not halo_delivers(A, x_ext, y_t):
            if not A.p2p:
                raise RuntimeError("the halo exchange delivers wrong ghost values")
            # same decision on every rank (all-reduced flag): rebuild on the NCCL path
            os.environ["B200_P2P"] = "0"
            del A, Aloc, xe_h, y_h
            A, x_ext, y_t, Aloc, xe_h, y_h = build_dist()
            halo_path = "NCCL halo exchange (the peer-memory exchange failed its validation on this box)"
            if not halo_delivers(A, x_ext, y_t):
                raise RuntimeError("the halo exchange delivers wrong ghost values")

        def step():  # halo exchange (referenced remote entries only) + local SpMV
            A.apply(x_ext, y_t)

        def kernel_step():
            api._hcheck(hl.gkob_apply(Aloc.h, xe_h.h, y_h.h))
    ex.synchronize()
    kernel_name = {2: "warp_stream_kernel", 4: "warp_pipe_kernel"}.get(
        hl.gkob_csr_kernel_variant((A if world == 1 else Aloc).h), "warp_stream_kernel")
    plan_parts = hl.gkob_csr_plan_parts((A if world == 1 else Aloc).h)
    for _ in range(max(args.warmup, 3)):
        step()
    with ClockSampler(local) as cs:
        l0 = ex.launch_count()
        ms_total = timed(step, args.steps)
        launches = ex.launch_count() - l0
        for _ in range(3):
            kernel_step()
        ms_kernel = timed(kernel_step, args.steps) / args.steps  # dominant kernel alone
    ms_step = ms_total / args.steps
    value = 2.0 * nnz_total / (ms_step * 1e-3) / 1e9

    # ----------------------------- end to end with HOST buffers: H2D x, SpMV, D2H y
    n_e2e = max(5, min(args.steps, 50))
    if world == 1:
        # the public host-buffer call: gko_b200::staged_apply (Csr::apply with HOST b and x;
        # upload, kernel and download of consecutive calls overlap on three streams)
        xh = x_full.cpu().pin_memory()
        yh = torch.empty(N_ROWS, dtype=torch.float64).pin_memory()
        staged = api.StagedApply(A)

        def e2e_step():
            staged.apply(xh, yh)
        e2e_join = staged.join
        h2d, d2h = N_ROWS * 8, N_ROWS * 8
    else:
        xh = x_full[r0:r1].cpu().pin_memory()
        yh = torch.empty(A.n_local, dtype=torch.float64).pin_memory()

        def e2e_step():
            with torch.cuda.stream(ex.stream):
                x_ext[:A.n_local].copy_(xh, non_blocking=True)
            step()
            with torch.cuda.stream(ex.stream):
                yh.copy_(y_t, non_blocking=True)

        def e2e_join():
            pass
        h2d, d2h = N_ROWS * 8, N_ROWS * 8  # summed over ranks
    for _ in range(3):
        e2e_step()
    e2e_join()

    def e2e_run():  # n_e2e calls, then the compute stream waits for the last download
        for _ in range(n_e2e):
            e2e_step()
        e2e_join()
    ms_e2e = timed(e2e_run, 1) / n_e2e
    if world == 1:  # the pipelined result must be the SpMV result
        ex.synchronize()
        assert torch.equal(yh, y_t.cpu()), "staged apply disagrees with the device apply"
    e2e = {"value": 2.0 * nnz_total / (ms_e2e * 1e-3) / 1e9, "unit": "GFLOP/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e}

    peak, peak_src = peaks()
    ncols_loc = N_ROWS if world == 1 else A.n_local + A.n_ghost
    alg_bytes = W.spmv_bytes(r1 - r0, ncols_loc, nnz_loc)
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel": "b200::csr::%s<double,int,1,...>" % kernel_name + (
                    " x %d launches (column-blocked copy, parts applied in order)" % plan_parts
                    if plan_parts > 1 else ""),
                "launches_per_step": max(plan_parts, 1),
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_kernel}
    prof = os.path.join(ROOT, "profiles", "r01_csr_spmv_cfg2.json")
    if os.path.exists(prof) and world == 1:  # the capture is of the 1-GPU workload
        try:
            pj = json.load(open(prof))
            # the capture is of the column-blocked SpMV (2 launches); without the copy one
            # launch moves what profiles/r01g_spmv_cfg2_random.json shows
            roofline["traffic"] = pj.get("dram_bytes_per_launch") if plan_parts > 1 else 4518669040.0
        except Exception:
            pass
    del A, rp, ci, va, x_full
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------------------ CG
    cg = run_cg(args, rank, world, ex, dev, timed_events=True)

    if rank != 0:
        return
    cpu = cpu_reference_spmv(CPU_SAMPLE_ROWS, 10) if world == 1 and not args.no_cpu else None
    line = {
        "metric": "csr_spmv_fp64_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s: CSR SpMV fp64/int32, random n=%d nnz=%d (15 distinct uniform "
                               "cols/row), workloads.py seed 42" % (CFG, N_ROWS, nnz_total),
                   "parallelism": "1-D row split over %d GPU(s)%s" %
                                  (world, ", %s of the referenced x entries per step" % halo_path
                                   if world > 1 else ""),
                   "l2": "inputs (2.0 GB/step) exceed the 126 MB L2; no flush between steps",
                   "gbs": alg_bytes * world / (ms_step * 1e-3) / 1e9},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": cs.summary(),
        "cg": cg,
    }
    if cpu:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))


def run_cg(args, rank, world, ex, dev, timed_events=True):
    """CG iterations/s: N == 1 -> BASELINE configs[2] (7-pt Laplacian 200^3 + scalar Jacobi, to
    1e-8) AND configs[4] (400^3, 200 iterations); N > 1 -> configs[4] row-sharded (z-slabs)."""
    import torch
    import torch.distributed as dist
    from ginkgo_b200 import api
    from ginkgo_b200 import distributed as D
    out = {}

    def wall(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ex.stream)
        fn()
        e1.record(ex.stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    if world == 1 and not args.no_cg:
        g = W.CONFIGS["cfg3"]["grid"]
        n = g ** 3
        with torch.cuda.stream(ex.stream):
            rp, ci, va = W.laplace(g, 3, xp="torch", device=dev)
            b = torch.ones(n, dtype=torch.float64, device=dev)
            x = torch.zeros(n, dtype=torch.float64, device=dev)
        A = api.host_csr(ex, (n, n), va, ci, rp)
        s = api.HostSolver(ex, "cg", A, precond_max_bs=1, max_iters=5000, reduction=1e-8, fused=True)
        bd, xd = api.host_dense(ex, b), api.host_dense(ex, x)
        s.apply(bd, xd)  # warm-up solve (builds the plan + graph)
        x.zero_()
        ms = wall(lambda: s.apply(bd, xd))
        # true relative residual
        r = b.clone()
        rd = api.host_dense(ex, r)
        one = api.host_dense(ex, torch.ones(1, dtype=torch.float64, device=dev))
        neg = api.host_dense(ex, -torch.ones(1, dtype=torch.float64, device=dev))
        api._hcheck(api._host().gkob_apply4(A.h, neg.h, xd.h, one.h, rd.h))
        ex.synchronize()
        nnz = va.numel()
        bytes_it = nnz * 12 + (n + 1) * 4 + 13 * n * 8
        out["cfg3"] = {"workload": "cfg3: CG + scalar Jacobi fp64, 7-pt Laplacian 200^3, b=1, "
                                   "ResidualNorm(rhs_norm) 1e-8", "iterations": s.num_iterations,
                       "ms": ms, "iters_per_s": s.num_iterations / (ms * 1e-3),
                       "true_rel_residual": (r.norm() / b.norm()).item(),
                       "fused": s.used_fused, "stop_status": s.stop_status,
                       "algorithmic_gbs": bytes_it * s.num_iterations / (ms * 1e-3) / 1e9}
        del A, s, rp, ci, va, b, x, r
        torch.cuda.empty_cache()
    if args.no_cg:
        return out
    if world == 1 and not args.no_gmres:
        # configs[3]: GMRES(30, MGS) + block-Jacobi(16) fp32, random nonsymmetric diagonally
        # dominant n=4M nnz=80M, uniform block pointers, rel. residual 1e-6
        c4 = W.CONFIGS["cfg4"]
        n = c4["n"]
        with torch.cuda.stream(ex.stream):
            rp, ci, va = W.build("cfg4", xp="torch", device=dev)
            b = torch.ones(n, dtype=torch.float32, device=dev)
            x = torch.zeros(n, dtype=torch.float32, device=dev)
        A = api.host_csr(ex, (n, n), va, ci, rp)
        bp = np.arange(0, n + 1, 16, dtype=np.int32)
        t0 = time.perf_counter()
        s = api.HostSolver(ex, "gmres", A, precond_max_bs=16, block_ptrs=bp, max_iters=1000,
                           reduction=1e-6, krylov_dim=30, ortho=0)
        bd, xd = api.host_dense(ex, b), api.host_dense(ex, x)
        s.apply(bd, xd)  # warm-up (includes the host-side block inversion of the generate step)
        setup_s = time.perf_counter() - t0
        x.zero_()
        ms = wall(lambda: s.apply(bd, xd))
        r = b.clone()
        rd = api.host_dense(ex, r)
        one = api.host_dense(ex, torch.ones(1, dtype=torch.float32, device=dev))
        neg = api.host_dense(ex, -torch.ones(1, dtype=torch.float32, device=dev))
        api._hcheck(api._host().gkob_apply4(A.h, neg.h, xd.h, one.h, rd.h))
        ex.synchronize()
        out["cfg4"] = {"workload": "cfg4: GMRES(30, MGS) + block-Jacobi(16) fp32, random "
                                   "nonsymmetric diag-dominant n=4M nnz=80M, b=1, 1e-6",
                       "iterations": s.num_iterations, "ms": ms,
                       "iters_per_s": s.num_iterations / (ms * 1e-3),
                       "true_rel_residual": (r.double().norm() / b.double().norm()).item(),
                       "stop_status": s.stop_status, "first_apply_incl_generate_s": setup_s}
        del A, s, rp, ci, va, b, x, r
        torch.cuda.empty_cache()
    # configs[4]: 400^3, fixed 200 iterations (reduction 0 never triggers), strong scaling
    g = W.CONFIGS["cfg5"]["grid"] if not args.small_cg else 160
    n = g ** 3
    offs = [p * g // world * g * g for p in range(world + 1)]  # whole z-slabs per rank
    r0, r1 = offs[rank], offs[rank + 1]
    iters = 200
    with torch.cuda.stream(ex.stream):
        rp, ci, va = W.laplace(g, 3, r0, r1, xp="torch", device=dev)
        b = torch.ones(r1 - r0, dtype=torch.float64, device=dev)
        x = torch.zeros(r1 - r0, dtype=torch.float64, device=dev)
    nnz_loc = va.numel()
    if world == 1:
        A = api.host_csr(ex, (n, n), va, ci, rp)
        s = api.HostSolver(ex, "cg", A, max_iters=iters, reduction=1e-300, fused=True, check_every=20)
        bd, xd = api.host_dense(ex, b), api.host_dense(ex, x)
        s.apply(bd, xd)
        x.zero_()
        ms = wall(lambda: s.apply(bd, xd))
        done, ghosts = s.num_iterations, 0
    else:
        def warm_up():
            """build + one untimed solve; every rank learns whether ALL ranks got through"""
            A_ = api.DistMatrix(ex, offs, rp, ci, va)
            A_.make_cg(scalar_jacobi=False, max_iters=iters, reduction=1e-300, check_every=20)
            good, why = 1, ""
            try:
                good = 1 if A_.cg_apply(b, x)[0] > 0 else 0
            except Exception as e:  # noqa: BLE001 (a peer-memory wait that timed out)
                good, why = 0, repr(e)
            t = torch.tensor([good], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return A_, int(t.item()) == 1, why

        A, good, why = warm_up()
        if not good:
            if not A.p2p:
                raise RuntimeError("distributed CG failed: " + why)
            os.environ["B200_P2P"] = "0"  # same decision on every rank: NCCL collectives
            del A
            A, good, why = warm_up()
            if not good:
                raise RuntimeError("distributed CG failed: " + why)
        x.zero_()
        res = {}
        ms = wall(lambda: res.update(it=A.cg_apply(b, x)[0]))
        done, ghosts = res["it"], A.n_ghost
    nnz_t = torch.tensor([nnz_loc], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz = int(nnz_t.item())
    bytes_it = nnz * 12 + (n + world) * 4 + 13 * n * 8
    out["cfg5"] = {"workload": "cfg5: CG fp64 (unpreconditioned), 7-pt Laplacian %d^3 n=%d nnz=%d, "
                               "b=1, %d iterations, rows split in z-slabs over %d GPU(s)"
                               % (g, n, nnz, iters, world), "iterations": done, "ms": ms,
                   "iters_per_s": done / (ms * 1e-3), "ghosts_per_rank": ghosts,
                   "collectives": ("single GPU" if world == 1 else
                                   "peer memory over NVLink" if A.p2p else "NCCL"),
                   "algorithmic_gbs": bytes_it * done / (ms * 1e-3) / 1e9}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cg", action="store_true", help="skip the CG legs")
    ap.add_argument("--no-gmres", action="store_true", help="skip the GMRES + block-Jacobi leg")
    ap.add_argument("--small-cg", action="store_true", help="160^3 instead of 400^3 for the cfg5 leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    run_gpu_arm(args, rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
