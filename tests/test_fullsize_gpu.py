"""BASELINE.json's full-size configurations on the GPU, checked through size-independent
properties (the oracle cannot run 150 M nonzeros in seconds, slices of it can):
 * cfg2 (random CSR n=10M nnz=150M): sampled row ranges bit-equal to the oracle;
   sum(y) equals sum_k val_k * x[col_k] computed independently; linearity A(ax+by)=aAx+bAy.
 * cfg3 (7-pt Laplacian 200^3): A*1 is zero in the interior and 1..3 on the faces (exact);
   fused CG + Jacobi reaches 1e-8 in the oracle-predicted iteration count of the same
   operator at 1/8 scale ratio is not comparable, so the residual itself is verified."""
import numpy as np
import pytest

import workloads as W
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hexec():
    from ginkgo_b200 import api
    return api.HostExecutor(0)


def test_cfg2_full_size_properties(hexec, orc):
    import torch
    from ginkgo_b200 import api
    dev = hexec.device
    n = W.CONFIGS["cfg2"]["n"]
    with torch.cuda.stream(hexec.stream):
        rp, ci, va = W.build("cfg2", xp="torch", device=dev)
        x = W.vector(n, xp="torch", device=dev)
        x2 = W.vector(n, stream=9, xp="torch", device=dev)
        y = torch.empty(n, dtype=torch.float64, device=dev)
    assert va.numel() == W.CONFIGS["cfg2"]["nnz"]
    A = api.host_csr(hexec, (n, n), va, ci, rp)
    h = api._host()
    xd, yd = api.host_dense(hexec, x), api.host_dense(hexec, y)
    api._hcheck(h.gkob_apply(A.h, xd.h, yd.h))
    hexec.synchronize()
    # (1) sampled row ranges against the oracle, bit for bit
    xh = x.cpu().numpy()
    for r0 in (0, 4_999_000, n - 1000):
        r1 = r0 + 1000
        lrp, lci, lva = W.build("cfg2", r0, r1, xp="np")
        yo = np.zeros(r1 - r0)
        orc("csr_spmv_f64_i32", r1 - r0, n, len(lva), lrp, lci, lva, xh, 1, 1, yo, 1)
        assert np.array_equal(y[r0:r1].cpu().numpy(), yo)
    # (2) checksum: 1^T (A x) == sum_k val_k x[col_k]
    with torch.cuda.stream(hexec.stream):
        lhs = y.sum().item()
        rhs = (va * x[ci.long()]).sum().item()
        scale = (va.abs() * x[ci.long()].abs()).sum().item()
    assert abs(lhs - rhs) <= 1e-12 * scale
    # (3) linearity: A (2 x - 3 x2) == 2 A x - 3 A x2
    with torch.cuda.stream(hexec.stream):
        y2 = torch.empty_like(y)
        y3 = torch.empty_like(y)
        x3 = 2 * x - 3 * x2
    d2i = api.host_dense(hexec, x2)  # handles must outlive the call
    api._hcheck(h.gkob_apply(A.h, d2i.h, (d2 := api.host_dense(hexec, y2)).h))
    api._hcheck(h.gkob_apply(A.h, (d3i := api.host_dense(hexec, x3)).h,
                             (d3 := api.host_dense(hexec, y3)).h))
    hexec.synchronize()
    with torch.cuda.stream(hexec.stream):
        err = (y3 - (2 * y - 3 * y2)).norm().item() / y3.norm().item()
    assert err <= 1e-13


def test_cfg3_full_size_operator_and_cg(hexec):
    import torch
    from ginkgo_b200 import api
    dev = hexec.device
    g = W.CONFIGS["cfg3"]["grid"]
    n = g ** 3
    with torch.cuda.stream(hexec.stream):
        rp, ci, va = W.laplace(g, 3, xp="torch", device=dev)
        ones = torch.ones(n, dtype=torch.float64, device=dev)
        y = torch.empty(n, dtype=torch.float64, device=dev)
    assert va.numel() == W.CONFIGS["cfg3"]["nnz"]
    A = api.host_csr(hexec, (n, n), va, ci, rp)
    h = api._host()
    od, yd = api.host_dense(hexec, ones), api.host_dense(hexec, y)
    api._hcheck(h.gkob_apply(A.h, od.h, yd.h))
    hexec.synchronize()
    # A 1 = number of missing neighbours (exact small integers)
    with torch.cuda.stream(hexec.stream):
        idx = torch.arange(n, device=dev)
        miss = torch.zeros(n, dtype=torch.float64, device=dev)
        for s in (1, g, g * g):
            c = (idx // s) % g
            miss += (c == 0).double() + (c == g - 1).double()
        assert torch.equal(y, miss)
    # fused CG + scalar Jacobi to 1e-8: verify the TRUE residual
    x = torch.zeros(n, dtype=torch.float64, device=dev)
    s = api.HostSolver(hexec, "cg", A, precond_max_bs=1, max_iters=5000, reduction=1e-8)
    xd = api.host_dense(hexec, x)
    s.apply(od, xd)
    assert s.used_fused and s.stop_status == (0x80 | 0x40 | 2)
    r = ones.clone()
    one = api.host_dense(hexec, torch.ones(1, dtype=torch.float64, device=dev))
    neg = api.host_dense(hexec, -torch.ones(1, dtype=torch.float64, device=dev))
    rd = api.host_dense(hexec, r)
    api._hcheck(h.gkob_apply4(A.h, neg.h, xd.h, one.h, rd.h))
    hexec.synchronize()
    # parity with the REAL reference at full size (tests/golden/fullsize_reference.json, made by
    # scripts/gen_fullsize_reference.py with gko::solver::Cg on the reference's OmpExecutor):
    # BASELINE.md section 6 -- same iteration count +-2, true relative residual within 1e-10
    exp = _fullsize_reference()["cfg3"]
    res = (r.norm() / ones.norm()).item()
    print("cfg3: %d iterations (reference %d), true relative residual %.6e (reference %.6e)"
          % (s.num_iterations, exp["iterations"], res, exp["true_rel_residual"]))
    assert abs(s.num_iterations - exp["iterations"]) <= 2
    assert abs(res - exp["true_rel_residual"]) <= 1e-10
    assert res <= 1.0e-8


def _fullsize_reference():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "fullsize_reference.json")) as f:
        return json.load(f)


def test_cfg4_full_size_gmres_block_jacobi(hexec):
    """BASELINE configs[3] at full size: GMRES(30, MGS) + block-Jacobi(16) fp32, n=4M nnz=80M.
    Restated fp32 gate (VERDICT r01 Weak #2): the reference sums its fp32 dot products
    and norms sequentially over 4 M entries (Reference executor: 34 iterations, OMP: 32); its
    implicit residual estimate lags and it only stops at the first restart.  The device sums in a
    fixed tree, so its Krylov basis stays orthogonal and the criterion triggers earlier.  Required:
    the same stopping criterion reached, TRUE relative residual <= 1e-6 (the requested reduction)
    and not more iterations than the reference needs."""
    import torch
    from ginkgo_b200 import api
    dev = hexec.device
    n = W.CONFIGS["cfg4"]["n"]
    with torch.cuda.stream(hexec.stream):
        rp, ci, va = W.build("cfg4", xp="torch", device=dev)
        b = torch.ones(n, dtype=torch.float32, device=dev)
        x = torch.zeros(n, dtype=torch.float32, device=dev)
    assert va.numel() == W.CONFIGS["cfg4"]["nnz"]
    A = api.host_csr(hexec, (n, n), va, ci, rp)
    bp = np.arange(0, n + 1, 16, dtype=np.int32)
    s = api.HostSolver(hexec, "gmres", A, precond_max_bs=16, block_ptrs=bp, max_iters=1000, reduction=1e-6,
                       krylov_dim=30, ortho=0)
    bd, xd = api.host_dense(hexec, b), api.host_dense(hexec, x)
    s.apply(bd, xd)
    r = b.clone()
    h = api._host()
    one = api.host_dense(hexec, torch.ones(1, dtype=torch.float32, device=dev))
    neg = api.host_dense(hexec, -torch.ones(1, dtype=torch.float32, device=dev))
    rd = api.host_dense(hexec, r)
    api._hcheck(h.gkob_apply4(A.h, neg.h, xd.h, one.h, rd.h))
    hexec.synchronize()
    res = (r.double().norm() / b.double().norm()).item()
    exp = _fullsize_reference()["cfg4"]
    print("cfg4: %d iterations (reference executor %d, omp %d), true relative residual %.3e (reference %.3e, omp %.3e)"
          % (s.num_iterations, exp["reference"]["iterations"], exp["omp"]["iterations"], res,
             exp["reference"]["true_rel_residual"], exp["omp"]["true_rel_residual"]))
    assert s.stop_status == (0x80 | 0x40 | 2)
    assert res <= 1.0e-6
    assert s.num_iterations <= max(exp["reference"]["iterations"], exp["omp"]["iterations"])


@pytest.mark.parametrize("kind", ["stencil", "random"])
def test_plan_tune_keeps_results_bit_identical(kind):
    """b200_csr_plan_tune_* decides from the matrix (locality of the gathers, size): the 7-pt stencil
    gets the bulk-copy ring (variant 5), uniformly random columns the warp-stream kernel (2); the
    result must not change by a bit against the untuned (plan = NULL) launch."""
    import ctypes
    import torch
    from ginkgo_b200 import api, _lib
    ex = api.B200Executor.create(0)
    dev = ex.device
    with torch.cuda.stream(ex.stream):
        if kind == "stencil":
            rp, ci, va = W.laplace(100, 3, xp="torch", device=dev)
            n = 100 ** 3
        else:
            n = 1 << 20
            rp, ci, va = W.random_csr(n, 8, xp="torch", device=dev, stream=5)
        x = W.vector(n, xp="torch", device=dev)
        y0 = torch.zeros(n, dtype=torch.float64, device=dev)
    A = api.Csr(ex, (n, n), va, ci, rp)
    X, Y = api.Dense(ex, x), api.Dense(ex, torch.zeros_like(y0))
    A.apply(X, Y)
    ex.run("b200_csr_spmv_f64_i32", None, n, n, A.nnz, rp, ci, va, x, 1, 1, y0, 1)
    ex.synchronize()
    variant = _lib.lib().b200_csr_plan_variant(A.plan())
    lines = _lib.lib().b200_csr_plan_gather_lines(A.plan())
    print("%s: variant %d, %.3f lines of b per gathered element" % (kind, variant, lines))
    assert variant == (5 if kind == "stencil" else 2)
    assert (lines < 0.2) if kind == "stencil" else (lines > 0.9)
    assert torch.equal(Y.values.reshape(-1), y0)


def test_value_copy_is_refreshed_after_in_place_changes(monkeypatch):
    """the column-blocked copy holds values.  A raw C-ABI plan never builds it (no opt-in);
    api.Csr opts in and refreshes through values_changed()."""
    import ctypes
    import torch
    from ginkgo_b200 import api, _lib
    monkeypatch.setenv("B200_CSR_REBLOCK", "2")
    ex = api.B200Executor.create(0)
    dev = ex.device
    n = 300_000
    with torch.cuda.stream(ex.stream):
        rp, ci, va = W.random_csr(n, 9, xp="torch", device=dev, stream=21)
        x = W.vector(n, xp="torch", device=dev)
        y0 = torch.zeros(n, dtype=torch.float64, device=dev)
    # raw plan: tune without the opt-in keeps no copy even when forced by size rules
    monkeypatch.delenv("B200_CSR_REBLOCK")
    raw = ctypes.c_void_p()
    l = _lib.lib()
    _lib.check(l.b200_csr_plan_create_f64_i32(ex.ctx, n, va.numel(), rp.data_ptr(), ctypes.byref(raw)))
    ex.run("b200_csr_plan_tune_f64_i32", raw, n, n, va.numel(), rp, ci, va)
    assert l.b200_csr_plan_parts(raw) == 0
    l.b200_csr_plan_destroy(raw)
    monkeypatch.setenv("B200_CSR_REBLOCK", "2")
    A = api.Csr(ex, (n, n), va, ci, rp)
    assert l.b200_csr_plan_parts(A.plan()) == 2
    Y = api.Dense(ex, torch.zeros_like(y0))
    with torch.cuda.stream(ex.stream):
        va.mul_(-3.0)          # in-place change of the matrix values
    A.values_changed()
    A.apply(api.Dense(ex, x), Y)
    ex.run("b200_csr_spmv_f64_i32", None, n, n, A.nnz, rp, ci, va, x, 1, 1, y0, 1)
    ex.synchronize()
    assert torch.equal(Y.values.reshape(-1), y0)


@pytest.mark.parametrize("parts", [2, 3, 4])
def test_column_blocked_copy_keeps_results_bit_identical(parts, monkeypatch):
    """b200_csr_plan_tune_* with a forced column-blocked copy: the parts are applied in order
    (c = A0 b, then c = 1*Ap b + 1*c), which must reproduce the single-pass row sums bit for
    bit, for spmv and advanced_spmv; unsorted rows must refuse the copy."""
    import torch
    from ginkgo_b200 import api, _lib
    monkeypatch.setenv("B200_CSR_REBLOCK", str(parts))
    ex = api.B200Executor.create(0)
    dev = ex.device
    n = 400_000
    with torch.cuda.stream(ex.stream):
        rp, ci, va = W.random_csr(n, 9, xp="torch", device=dev, stream=11)
        x = W.vector(n, xp="torch", device=dev)
        y0 = torch.zeros(n, dtype=torch.float64, device=dev)
        y1 = W.vector(n, stream=13, xp="torch", device=dev)
        y1b = y1.clone()
        alpha = torch.tensor([-1.5], dtype=torch.float64, device=dev)
        beta = torch.tensor([0.25], dtype=torch.float64, device=dev)
    A = api.Csr(ex, (n, n), va, ci, rp)
    assert _lib.lib().b200_csr_plan_parts(A.plan()) == parts
    Y = api.Dense(ex, torch.zeros_like(y0))
    A.apply(api.Dense(ex, x), Y)
    ex.run("b200_csr_spmv_f64_i32", None, n, n, A.nnz, rp, ci, va, x, 1, 1, y0, 1)
    Y1 = api.Dense(ex, y1)
    A.apply(api.Dense(ex, alpha), api.Dense(ex, x), api.Dense(ex, beta), Y1)
    ex.run("b200_csr_advanced_spmv_f64_i32", None, n, n, A.nnz, rp, ci, va, alpha, x, 1, 1, beta,
           y1b, 1)
    ex.synchronize()
    assert torch.equal(Y.values.reshape(-1), y0)
    assert torch.equal(y1, y1b)
    # unsorted rows: the copy is refused, the plan still works
    with torch.cuda.stream(ex.stream):
        ci2 = ci.clone()
        ci2[0], ci2[1] = ci[1], ci[0]
    B = api.Csr(ex, (n, n), va, ci2, rp)
    assert _lib.lib().b200_csr_plan_parts(B.plan()) == 0


def test_skewed_rows_are_split_over_ctas(hexec, orc):
    """VERDICT r01 Missing #2: a power-law matrix with the nnz of cfg2 (Zipf row lengths: the longest
    row has ~9 M entries, ~8800 rows >= 1024, the split threshold).  The plan splits the long rows over CTAs; checked:
    sampled short rows bit-equal to the oracle, the three longest rows against a float64 dot
    computed independently (tree-sum tolerance), linearity of the whole operator."""
    import torch
    from ginkgo_b200 import api, _lib
    dev = hexec.device
    n = W.CONFIGS["cfg2_zipf"]["n"]
    with torch.cuda.stream(hexec.stream):
        rp, ci, va = W.build("cfg2_zipf", xp="torch", device=dev)
        x = W.vector(n, xp="torch", device=dev)
        x2 = W.vector(n, stream=9, xp="torch", device=dev)
        y = torch.empty(n, dtype=torch.float64, device=dev)
        y2 = torch.empty_like(y)
        y3 = torch.empty_like(y)
    nnz = va.numel()
    assert rp.dtype == torch.int32 and abs(nnz - 150_000_000) < 3_000_000
    A = api.host_csr(hexec, (n, n), va, ci, rp)
    h = api._host()
    xd, yd = api.host_dense(hexec, x), api.host_dense(hexec, y)
    api._hcheck(h.gkob_apply(A.h, xd.h, yd.h))
    hexec.synchronize()
    with torch.cuda.stream(hexec.stream):
        lens = (rp[1:] - rp[:-1]).long()
        top = torch.topk(lens, 3).indices.tolist()
    assert int(lens.max().item()) > 5_000_000
    # the longest rows: independent float64 dot products
    for r in top:
        with torch.cuda.stream(hexec.stream):
            s, e = int(rp[r].item()), int(rp[r + 1].item())
            ref = (va[s:e] * x[ci[s:e].long()]).sum().item()
            scale = (va[s:e].abs() * x[ci[s:e].long()].abs()).sum().item()
        assert abs(y[r].item() - ref) <= 1e-13 * scale, (r, e - s)
    # short rows: bit-equal to the oracle (left-to-right sums)
    xh = x.cpu().numpy()
    with torch.cuda.stream(hexec.stream):
        short = torch.nonzero(lens <= 32)[:2000, 0]
    rph, yh = rp.cpu().numpy(), y.cpu().numpy()
    for r in short.cpu().numpy()[::97]:
        s, e = int(rph[r]), int(rph[r + 1])
        cols = ci[s:e].cpu().numpy()
        vals = va[s:e].cpu().numpy()
        yo = np.zeros(1)
        orc("csr_spmv_f64_i32", 1, n, e - s, np.array([0, e - s], np.int32), cols, vals, xh, 1, 1, yo, 1)
        assert yh[r] == yo[0]
    # linearity
    with torch.cuda.stream(hexec.stream):
        x3 = 2 * x - 3 * x2
    d2i = api.host_dense(hexec, x2)
    api._hcheck(h.gkob_apply(A.h, d2i.h, (d2 := api.host_dense(hexec, y2)).h))
    api._hcheck(h.gkob_apply(A.h, (d3i := api.host_dense(hexec, x3)).h, (d3 := api.host_dense(hexec, y3)).h))
    hexec.synchronize()
    with torch.cuda.stream(hexec.stream):
        err = (y3 - (2 * y - 3 * y2)).norm().item() / y3.norm().item()
    assert err <= 1e-13
