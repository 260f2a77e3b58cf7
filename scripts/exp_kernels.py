"""GPU experiment (harness only): time every kernel family of the hot path alone and report
its algorithmic GB/s against the measured HBM peak -- the per-row roofline table of DESIGN.md
(SURVEY.md section 8a rows a1-a13).  CUDA events on the executor's stream, 5 warm-ups,
inputs larger than L2.  Writes gpurun_out/exp_kernels.json."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import workloads as W
from ginkgo_b200 import api

PEAK = 6582.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

ex = api.B200Executor.create(0)
hx = api.HostExecutor(0)
dev = ex.device
rows = []


def timeit(fn, reps=30, stream=None):
    stream = stream or ex.stream
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(row, what, nbytes, ms):
    r = dict(row=row, kernel=what, bytes=nbytes, ms=ms, gbs=nbytes / ms / 1e6,
             frac=nbytes / ms / 1e6 / PEAK)
    rows.append(r)
    print(json.dumps(r), flush=True)


# ------------------------------------------------------------------ a3-a6: formats on cfg3
def formats(name, rp, ci, va):
    n = rp.numel() - 1
    nnz = va.numel()
    vb = va.element_size()
    A = api.host_csr(hx, (n, n), va, ci, rp)
    with torch.cuda.stream(hx.stream):
        x = W.vector(n, xp="torch", device=dev).to(va.dtype)
        y = torch.zeros(n, dtype=va.dtype, device=dev)
    xd, yd = api.host_dense(hx, x), api.host_dense(hx, y)
    h = api._host()
    lens = (rp[1:] - rp[:-1]).to(torch.int64)
    width = int(lens.max().item())
    ops = {"csr (a1)": (A, nnz * (vb + 4) + (n + 1) * 4 + 2 * n * vb)}
    ell = api.host_convert(A, "ell")
    ops["ell (a3)"] = (ell, width * n * (vb + 4) + 2 * n * vb)
    sp = api.host_convert(A, "sellp")
    ns = (n + 63) // 64
    # SELL-P stores slice_length * 64 per slice; take it from the per-slice maxima
    pad = torch.zeros(ns * 64, dtype=torch.int64, device=dev)
    pad[:n] = lens
    stored = int(pad.reshape(ns, 64).max(dim=1).values.sum().item()) * 64
    ops["sellp (a4)"] = (sp, stored * (vb + 4) + ns * 16 + 2 * n * vb)
    coo = api.host_convert(A, "coo")
    ops["coo (a5)"] = (coo, nnz * (vb + 8) + 2 * n * vb)
    hyb = api.host_convert(A, "hybrid", strategy="imbalance_limit", percent=0.5)
    k = int(torch.sort(lens).values[int(n * 0.5)].item())
    ell_part = k * n
    coo_part = int(torch.clamp(lens - k, min=0).sum().item())
    ops["hybrid (a6)"] = (hyb, ell_part * (vb + 4) + coo_part * (vb + 8) + 3 * n * vb + n * vb)
    for nm, (op, nbytes) in ops.items():
        ms = timeit(lambda: api._hcheck(h.gkob_apply(op.h, xd.h, yd.h)), stream=hx.stream)
        report(nm, "%s SpMV on %s (n=%d nnz=%d)" % (nm.split()[0], name, n, nnz), nbytes, ms)


with torch.cuda.stream(hx.stream):
    rp, ci, va = W.build("cfg3", xp="torch", device=dev)
formats("cfg3", rp, ci, va)
# several right-hand sides (Missing #6 of VERDICT r01): b is n x 8 row-major
with torch.cuda.stream(hx.stream):
    n3 = rp.numel() - 1
    xb = torch.rand(n3, 8, dtype=torch.float64, device=dev)
    yb = torch.zeros(n3, 8, dtype=torch.float64, device=dev)
report("csr (a1)", "csr SpMV, 8 right-hand sides, on cfg3", int(va.numel()) * 12 + (n3 + 1) * 4 + 2 * n3 * 8 * 8,
       timeit(lambda: ex.run("b200_csr_spmv_f64_i32", None, n3, n3, va.numel(), rp, ci, va, xb, 8, 8, yb, 8)))
del rp, ci, va, xb, yb
torch.cuda.empty_cache()

# CSR on the other input classes: banded twin (ring kernel), cfg2 (warp_stream x 2 column blocks),
# skewed twin (rows split over CTAs)
for cfg in ("cfg2_banded", "cfg2", "cfg2_zipf"):
    with torch.cuda.stream(hx.stream):
        rp, ci, va = W.build(cfg, xp="torch", device=dev)
        n2 = rp.numel() - 1
        x2 = W.vector(n2, xp="torch", device=dev)
        y2 = torch.zeros(n2, dtype=torch.float64, device=dev)
    A2 = api.host_csr(hx, (n2, n2), va, ci, rp)
    xd2, yd2 = api.host_dense(hx, x2), api.host_dense(hx, y2)
    h2 = api._host()
    ms = timeit(lambda: api._hcheck(h2.gkob_apply(A2.h, xd2.h, yd2.h)), stream=hx.stream)
    report("csr (a1)", "csr SpMV on %s (n=%d nnz=%d, kernel variant %d, %d column blocks)"
           % (cfg, n2, va.numel(), h2.gkob_csr_kernel_variant(A2.h), max(1, h2.gkob_csr_plan_parts(A2.h))),
           W.spmv_bytes(n2, n2, va.numel()), ms)
    del A2, rp, ci, va, x2, y2, xd2, yd2
    torch.cuda.empty_cache()

# ------------------------------------------------------------------ a7-a13: vector kernels
n = 32_000_000
with torch.cuda.stream(ex.stream):
    v = [W.vector(n, stream=30 + i, xp="torch", device=dev) for i in range(5)]
    res = torch.zeros(8, dtype=torch.float64, device=dev)
    alpha = torch.tensor([0.5], dtype=torch.float64, device=dev)
    rho = torch.tensor([1.5], dtype=torch.float64, device=dev)
    prev = torch.tensor([1.25], dtype=torch.float64, device=dev)
    stop = torch.zeros(1, dtype=torch.uint8, device=dev)
    dinv = W.vector(n, stream=40, xp="torch", device=dev).abs() + 0.5
V = 8
report("a7", "dense::compute_dot n=32M", 2 * n * V,
       timeit(lambda: ex.run("b200_dense_compute_dot_f64", n, 1, v[0], 1, v[1], 1, res)))
report("a7", "dense::compute_norm2 n=32M", n * V,
       timeit(lambda: ex.run("b200_dense_compute_norm2_f64", n, 1, v[0], 1, res)))
report("a8", "dense::add_scaled n=32M", 3 * n * V,
       timeit(lambda: ex.run("b200_dense_add_scaled_f64", n, 1, alpha, 1, v[0], 1, v[1], 1)))
report("a8", "dense::scale n=32M", 2 * n * V,
       timeit(lambda: ex.run("b200_dense_scale_f64", n, 1, alpha, 1, v[2], 1)))
report("a9", "cg::step_1 n=32M", 3 * n * V,
       timeit(lambda: ex.run("b200_cg_step_1_f64", n, 1, v[2], 1, v[3], 1, rho, prev, stop)))
report("a9", "cg::step_2 n=32M", 6 * n * V,
       timeit(lambda: ex.run("b200_cg_step_2_f64", n, 1, v[0], 1, v[1], 1, v[2], 1, v[3], 1, rho, rho, stop)))
report("a13", "jacobi::simple_scalar_apply n=32M", 3 * n * V,
       timeit(lambda: ex.run("b200_jacobi_simple_scalar_apply_f64", n, 1, dinv, v[0], 1, v[4], 1)))
del v, dinv
torch.cuda.empty_cache()

# GMRES multi_dot / multi_axpy: 30 basis vectors of 4M fp32 (cfg4 shape)
n4, kd = 4_000_000, 30
with torch.cuda.stream(ex.stream):
    basis = torch.rand((kd + 1) * n4, dtype=torch.float32, device=dev)
    w = torch.rand(n4, dtype=torch.float32, device=dev)
    hcol = torch.zeros(kd + 2, dtype=torch.float32, device=dev)
    yv = torch.rand(kd, dtype=torch.float32, device=dev)
    out = torch.zeros(n4, dtype=torch.float32, device=dev)
    fin = torch.full((1,), kd, dtype=torch.int64, device=dev)
    stop = torch.zeros(1, dtype=torch.uint8, device=dev)
report("a11", "gmres::multi_dot 30 x 4M fp32", (kd + 1) * n4 * 4,
       timeit(lambda: ex.run("b200_gmres_multi_dot_f32", n4, 1, kd, basis, 1, w, 1, hcol, 1)))
report("a11", "gmres::multi_axpy 30 x 4M fp32", (kd + 1) * n4 * 4,
       timeit(lambda: ex.run("b200_gmres_multi_axpy_f32", n4, 1, basis, 1, yv, 1, out, 1, fin, stop)))
del basis, w, out
torch.cuda.empty_cache()

# block-Jacobi apply + generate, cfg4 shape: 250k blocks of 16x16 fp32
nb, bs = 250_000, 16
with torch.cuda.stream(ex.stream):
    rp, ci, va = W.build("cfg4", xp="torch", device=dev)
    bp = torch.arange(0, n4 + 1, bs, dtype=torch.int32, device=dev)
    blocks = torch.zeros(nb * bs * bs, dtype=torch.float32, device=dev)
    b = torch.rand(n4, dtype=torch.float32, device=dev)
    x = torch.zeros(n4, dtype=torch.float32, device=dev)
gp = 1  # group_size = 32 / 16 = 2
report("8f-2", "jacobi::generate 250k x 16x16 fp32 (reads the block rows of A)",
       int(va.numel()) * 8 + (n4 + 1) * 4 + nb * bs * bs * 4,
       timeit(lambda: ex.run("b200_jacobi_generate_f32_i32", n4, rp, ci, va, nb, bs, bs, bs * 2 * bs, gp,
                             bp, blocks), reps=10))
report("a13", "jacobi::simple_apply 250k x 16x16 fp32", nb * bs * bs * 4 + 2 * n4 * 4,
       timeit(lambda: ex.run("b200_jacobi_simple_apply_f32_i32", nb, bs, bs, bs * 2 * bs, gp, bp, blocks, b,
                             1, 1, x, 1)))
# adaptive-precision generate (autodetect: conditioning + verification inversions) and apply of the result
with torch.cuda.stream(ex.stream):
    prec = torch.full((nb,), 0xFF, dtype=torch.uint8, device=dev)
    cond = torch.zeros(nb, dtype=torch.float32, device=dev)
    ablocks = torch.zeros(nb * bs * bs, dtype=torch.float32, device=dev)


def gen_adaptive():
    prec.fill_(0xFF)
    ex.run("b200_jacobi_generate_adaptive_f32_i32", n4, rp, ci, va, nb, bs, 0.1, bs, bs * 2 * bs, gp, cond, prec,
           bp, ablocks)


report("8f-2", "jacobi::generate adaptive (autodetect) 250k x 16x16 fp32",
       int(va.numel()) * 8 + (n4 + 1) * 4 + nb * bs * bs * 2, timeit(gen_adaptive, reps=10))
ex.synchronize()
hist = torch.bincount(prec.long(), minlength=256)
stored = int(hist[0].item()) * 4 + int((hist.sum() - hist[0]).item()) * 2
print("# adaptive precisions chosen: " + ", ".join("0x%02x: %d" % (i, int(c)) for i, c in enumerate(hist.tolist()) if c),
      flush=True)
report("a13", "jacobi::simple_apply_adaptive 250k x 16x16 fp32 (stored %.0f %% of full)" % (100.0 * stored / (4 * nb)),
       stored * bs * bs + 2 * n4 * 4,
       timeit(lambda: ex.run("b200_jacobi_simple_apply_adaptive_f32_i32", nb, bs, bs, bs * 2 * bs, gp, prec, bp,
                             ablocks, b, 1, 1, x, 1)))
# many right-hand sides: the block apply as a batch of small GEMMs, SIMT kernel vs fp64 tensor cores
from ginkgo_b200 import _lib as _L
for vt_name, tdt, vb in (("f32", torch.float32, 4), ("f64", torch.float64, 8)):
    for nrhs in (8, 32):
        with torch.cuda.stream(ex.stream):
            blk = torch.rand(nb * bs * bs, dtype=tdt, device=dev)
            bm = torch.rand(n4 * nrhs, dtype=tdt, device=dev)
            xm = torch.zeros(n4 * nrhs, dtype=tdt, device=dev)
        nbytes = nb * bs * bs * vb + 2 * n4 * nrhs * vb
        flops = 2.0 * nb * bs * bs * nrhs
        for mode, label in ((0, "SIMT"), (1, "fp64 tensor cores")):
            _L.lib().b200_jacobi_apply_mode(mode)
            ms = timeit(lambda: ex.run("b200_jacobi_simple_apply_%s_i32" % vt_name, nb, bs, bs, bs * 2 * bs, gp, bp,
                                       blk, bm, nrhs, nrhs, xm, nrhs), reps=10)
            report("a13", "jacobi::simple_apply 250k x 16x16 %s, %d rhs, %s (%.1f TFLOP/s)" % (
                vt_name, nrhs, label, flops / ms / 1e9), nbytes, ms)
        _L.lib().b200_jacobi_apply_mode(-1)
        del blk, bm, xm
        torch.cuda.empty_cache()
# conversions on cfg4
n = n4
with torch.cuda.stream(ex.stream):
    lens = (rp[1:] - rp[:-1])
    width = int(lens.max().item())
    ecols = torch.empty(width * n, dtype=torch.int32, device=dev)
    evals = torch.empty(width * n, dtype=torch.float32, device=dev)
report("8f-1", "csr::convert_to_ell cfg4 (fp32)", int(va.numel()) * 8 + width * n * 8,
       timeit(lambda: ex.run("b200_csr_convert_to_ell_f32_i32", n, rp, ci, va, width, n, ecols, evals), reps=10))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(peak_gbs=PEAK, rows=rows), open(os.environ.get("EXP_KERNELS_OUT", "gpurun_out/exp_kernels.json"), "w"), indent=1)
