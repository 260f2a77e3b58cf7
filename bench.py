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
  e2e          same metric through the public API with HOST buffers (H2D x, SpMV, D2H y)
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
    from oracle import ref
    if ref.available():
        cores = ref.num_threads()
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
    from ginkgo_b200.api import B200Executor, Csr, Dense

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ex = B200Executor.create(local)
    dev = ex.device

    r0, r1 = rank * N_ROWS // world, (rank + 1) * N_ROWS // world
    with torch.cuda.stream(ex.stream):
        rp, ci, va = W.build(CFG, r0, r1, xp="torch", device=dev)
        x_full = W.vector(N_ROWS, xp="torch", device=dev)
        x_loc = x_full[r0:r1].clone()
    A = Csr(ex, (r1 - r0, N_ROWS), va, ci, rp)
    A.plan()
    nnz_loc = va.numel()
    xg = Dense(ex, x_full.reshape(-1, 1))
    y = Dense.create(ex, (r1 - r0, 1))
    ex.synchronize()

    def step():
        if world > 1:
            with torch.cuda.stream(ex.stream):
                dist.all_gather_into_tensor(x_full, x_loc)
        A.apply(xg, y)

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

    for _ in range(max(args.warmup, 3)):
        step()
    with ClockSampler(local) as cs:
        l0 = ex.launch_count()
        ms_total = timed(step, args.steps)
        launches = ex.launch_count() - l0
        # kernel-only time of the dominant kernel (no collective), same stream, CUDA events
        for _ in range(3):
            A.apply(xg, y)
        ms_kernel = timed(lambda: A.apply(xg, y), args.steps) / args.steps
    ms_step = ms_total / args.steps
    nnz_total = N_ROWS * PER_ROW
    value = 2.0 * nnz_total / (ms_step * 1e-3) / 1e9

    # end to end through the public API with host buffers: H2D x, SpMV, D2H y
    xh = Dense(None, x_full.cpu().pin_memory().reshape(-1, 1))
    yh = Dense(None, torch.empty((r1 - r0, 1), dtype=torch.float64).pin_memory())
    if world == 1:
        def e2e_step():
            A.apply(xh, yh)
        for _ in range(3):
            e2e_step()
        n_e2e = max(5, min(args.steps, 50))
        ms_e2e = timed(e2e_step, n_e2e) / n_e2e
        e2e = {"value": 2.0 * nnz_total / (ms_e2e * 1e-3) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": N_ROWS * 8, "d2h_bytes_per_step": (r1 - r0) * 8,
               "ms_per_step": ms_e2e}
    else:
        xl_h = Dense(None, x_loc.cpu().pin_memory().reshape(-1, 1))

        def e2e_step():
            with torch.cuda.stream(ex.stream):
                x_loc.copy_(xl_h.values.reshape(-1), non_blocking=True)
                dist.all_gather_into_tensor(x_full, x_loc)
            A.apply(xg, y)
            with torch.cuda.stream(ex.stream):
                yh.values.copy_(y.values, non_blocking=True)
        for _ in range(3):
            e2e_step()
        n_e2e = max(5, min(args.steps, 50))
        ms_e2e = timed(e2e_step, n_e2e) / n_e2e
        e2e = {"value": 2.0 * nnz_total / (ms_e2e * 1e-3) / 1e9, "unit": "GFLOP/s",
               "h2d_bytes_per_step": (r1 - r0) * 8 * world, "d2h_bytes_per_step": (r1 - r0) * 8 * world,
               "ms_per_step": ms_e2e}

    peak, peak_src = peaks()
    alg_bytes = W.spmv_bytes(r1 - r0, N_ROWS, nnz_loc)
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "kernel": "b200::csr::slab_kernel<double,int,1,false,true>",
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": ms_kernel}
    prof = os.path.join(ROOT, "profiles", "r01_csr_spmv_cfg2.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

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
                                  (world, ", NCCL all-gather of x per step" if world > 1 else ""),
                   "l2": "inputs (2.0 GB/step) exceed the 126 MB L2; no flush between steps",
                   "gbs": alg_bytes * world / (ms_step * 1e-3) / 1e9},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": cs.summary(),
    }
    if cpu:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
