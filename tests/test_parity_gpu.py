"""Parity of the CUDA library (through the C ABI) against the oracle on seeded
random inputs -- the shape of the reference's `test/**` common tests (run the op
on ReferenceExecutor and on the device executor, compare with
GKO_ASSERT_MTX_NEAR(…, r<T>)): test/matrix/csr_kernels2.cpp:218-445,
test/solver/cg_kernels.cpp:42-158, test/preconditioner/jacobi_kernels.cpp:426-660.

Bit-exact (np.array_equal) wherever the CUDA kernel keeps the reference's
operation order (element-wise kernels, SpMV rows summed by one thread);
otherwise the reference tolerance r<T> = 10 eps, scaled by sqrt(row length) for
tree-ordered sums."""
import numpy as np
import pytest

from tests import helpers as H
from tests.helpers import IT, R, VT, rel_err

pytestmark = pytest.mark.gpu
VTS = ["f64", "f32"]
ITS = ["i32", "i64"]


def both(orc, cuda, fname, args_fn, plan=None):
    """run `fname` on both backends on identical inputs; returns the two arg lists"""
    a = args_fn()
    b = [x.copy() if isinstance(x, np.ndarray) else x for x in a]
    orc(fname, *a)
    cuda(fname, *b, plan=plan)
    return a, b


# --------------------------------------------------------------------------- CSR
def csr_case(rng, kind, vt, it):
    if kind == "ref_common":  # test/matrix/csr_kernels2.cpp: 532 x 231, 1..50 per row
        n, m = 532, 231
        lens = rng.integers(1, 51, size=n)
    elif kind == "empty_rows":
        n, m = 3000, 3000
        lens = rng.integers(0, 4, size=n)
        lens[rng.integers(0, n, size=n // 2)] = 0
        lens[:700] = 0
        lens[-900:] = 0
    elif kind == "laplace_like":
        n, m = 20000, 20000
        lens = np.full(n, 5)
    elif kind == "wide_rows":  # avg 100 -> LANES 4
        n, m = 900, 5000
        lens = rng.integers(60, 140, size=n)
    elif kind == "long_rows":  # rows longer than a tile and than the products buffer
        n, m = 40, 20000
        lens = rng.integers(0, 30, size=n)
        lens[3], lens[17], lens[18], lens[39] = 2500, 5000, 9000, 4200
    elif kind == "split_rows":  # rows >= 1024 entries (16384 until r02j): split over CTAs by the plan (long_rows_kernel)
        n, m = 300, 70000
        lens = rng.integers(0, 12, size=n)
        lens[0], lens[5], lens[6], lens[150], lens[299] = 16384, 16383, 50000, 24577, 33000
        lens[10], lens[11], lens[12] = 4096, 4095, 8193  # chunk size, one below, one entry into a third chunk
        lens[13], lens[14], lens[15], lens[16] = 1024, 1023, 64, 63  # split threshold; warp-summed rows (>= 64)
    elif kind == "one_row":
        n, m = 1, 7000
        lens = np.array([6500])
    elif kind == "all_empty":
        n, m = 5000, 10
        lens = np.zeros(n, dtype=np.int64)
    else:
        raise ValueError(kind)
    rp, ci, va = H.random_csr(rng, n, m, lens, vt, it, sort=(kind != "ref_common"))
    return n, m, rp, ci, va


CSR_KINDS = ["ref_common", "empty_rows", "laplace_like", "wide_rows", "long_rows", "split_rows", "one_row",
             "all_empty"]
EXACT_KINDS = {"ref_common", "empty_rows", "laplace_like", "all_empty"}  # LANES == 1, rows fit


@pytest.mark.parametrize("kind", CSR_KINDS)
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("use_plan", [False, True, "pipe", "ring"])
def test_csr_spmv_vector(orc, cuda, kind, vt, it, use_plan):
    rng = np.random.default_rng(42)
    n, m, rp, ci, va = csr_case(rng, kind, vt, it)
    nnz = len(va)
    plan = cuda.make_csr_plan(vt, it, n, nnz, rp) if use_plan else None
    if use_plan in ("pipe", "ring") and plan is not None:
        # every kernel variant on every case: the bulk-copy ring (csr_ring.cuh) incl. its tails (nnz and
        # num_rows + 1 not multiples of 4), rows longer than a stage, empty tiles inside long rows
        cuda.l.b200_csr_plan_set_variant(plan, {"pipe": 4, "ring": 5}[use_plan])
    x = H.dense(rng, m, 1, vt=vt)
    tol = R[vt] * max(1.0, np.sqrt(np.diff(rp.astype(np.int64)).max(initial=1)))

    def simple():
        return [n, m, nnz, rp, ci, va, x, 1, 1, np.full((n, 1), np.nan, VT[vt]), 1]

    a, b = both(orc, cuda, "csr_spmv_%s_%s" % (vt, it), simple, plan)
    if kind in EXACT_KINDS:
        assert np.array_equal(a[-2], b[-2])
    assert rel_err(a[-2], b[-2]) <= tol

    y0 = H.dense(rng, n, 1, vt=vt)
    for alpha, beta in [(-1.0, 2.0), (0.5, 0.0), (2.0, 1.0)]:
        def adv():
            y = y0.copy()
            if beta == 0.0:
                y[:] = np.nan  # beta == 0 must overwrite (csr_kernels.cpp:106)
            return [n, m, nnz, rp, ci, va, np.array([alpha], VT[vt]), x, 1, 1,
                    np.array([beta], VT[vt]), y, 1]
        a, b = both(orc, cuda, "csr_advanced_spmv_%s_%s" % (vt, it), adv, plan)
        if kind in EXACT_KINDS:
            assert np.array_equal(a[-2], b[-2])
        assert rel_err(a[-2], b[-2]) <= tol


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
def test_csr_spmv_multi_rhs_strided(orc, cuda, vt, it):
    rng = np.random.default_rng(7)
    n, m, rp, ci, va = csr_case(rng, "ref_common", vt, it)
    nnz = len(va)
    # aligned and odd strides / counts: P lanes per row, grid.y tiles of P right-hand sides
    for nrhs, bs, cs in [(3, 3, 3), (3, 5, 4), (43, 45, 46), (8, 8, 8), (8, 12, 10), (12, 12, 16), (64, 64, 64),
                         (4, 4, 4), (136, 136, 140)]:
        x = H.dense(rng, m, nrhs, bs, vt)
        y0 = H.dense(rng, n, nrhs, cs, vt)
        a, b = both(orc, cuda, "csr_spmv_%s_%s" % (vt, it),
                    lambda: [n, m, nnz, rp, ci, va, x, bs, nrhs, y0.copy(), cs])
        assert np.array_equal(a[-2], b[-2])  # padding columns untouched, values bit-equal
        a, b = both(orc, cuda, "csr_advanced_spmv_%s_%s" % (vt, it),
                    lambda: [n, m, nnz, rp, ci, va, np.array([-0.75], VT[vt]), x, bs, nrhs,
                             np.array([1.5], VT[vt]), y0.copy(), cs])
        assert np.array_equal(a[-2], b[-2])


def test_csr_unaligned_views(orc, cuda):
    """values/col_idxs not 32-byte aligned -> scalar-load variant of the slab kernel"""
    import ctypes
    import torch
    rng = np.random.default_rng(3)
    n, m, rp, ci, va = csr_case(rng, "laplace_like", "f64", "i32")
    nnz = len(va)
    x = H.dense(rng, m, 1)
    y_ref = np.zeros((n, 1))
    orc("csr_spmv_f64_i32", n, m, nnz, rp, ci, va, x, 1, 1, y_ref, 1)
    with torch.cuda.stream(cuda.stream):
        t_ci = torch.zeros(nnz + 1, dtype=torch.int32, device="cuda")
        t_va = torch.zeros(nnz + 1, dtype=torch.float64, device="cuda")
        t_ci[1:] = torch.from_numpy(ci).cuda()
        t_va[1:] = torch.from_numpy(va).cuda()
        t_rp, t_x = torch.from_numpy(rp).cuda(), torch.from_numpy(x).cuda()
        t_y = torch.zeros(n, dtype=torch.float64, device="cuda")
        from ginkgo_b200 import _lib
        _lib.check(cuda.l.b200_csr_spmv_f64_i32(cuda.ctx, None, n, m, nnz, t_rp.data_ptr(),
                                                t_ci[1:].data_ptr(), t_va[1:].data_ptr(),
                                                t_x.data_ptr(), 1, 1, t_y.data_ptr(), 1))
        cuda.stream.synchronize()
    assert np.array_equal(t_y.cpu().numpy(), y_ref[:, 0])


# ----------------------------------------------------------------- ELL / SELL-P / COO
@pytest.mark.parametrize("kind", ["ref_common", "empty_rows", "laplace_like", "wide_small"])
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
def test_ell_spmv(orc, cuda, kind, vt, it):
    rng = np.random.default_rng(11)
    if kind == "wide_small":  # few rows, wide -> LANES > 1 (tree order)
        n, m = 300, 4000
        rp, ci, va = H.random_csr(rng, n, m, rng.integers(50, 200, size=n), vt, it)
    else:
        n, m, rp, ci, va = csr_case(rng, kind, vt, it)
    width, stride, cols, vals = H.csr_to_ell(rp, ci, va, n, pad_extra=2, stride_extra=3)
    for nrhs in (1, 3):
        x = H.dense(rng, m, nrhs, nrhs + 1, vt)
        y0 = H.dense(rng, n, nrhs, nrhs + 2, vt)
        a, b = both(orc, cuda, "ell_spmv_%s_%s" % (vt, it),
                    lambda: [n, m, width, stride, cols, vals, x, nrhs + 1, nrhs, y0.copy(), nrhs + 2])
        exact = width < 16  # wider ELL with few rows splits rows over lanes (tree order)
        if exact:
            assert np.array_equal(a[-2], b[-2])
        assert rel_err(a[-2], b[-2]) <= R[vt] * 15
        a, b = both(orc, cuda, "ell_advanced_spmv_%s_%s" % (vt, it),
                    lambda: [n, m, width, stride, cols, vals, np.array([-1.5], VT[vt]), x, nrhs + 1,
                             nrhs, np.array([0.5], VT[vt]), y0.copy(), nrhs + 2])
        if exact:
            assert np.array_equal(a[-2], b[-2])
        assert rel_err(a[-2], b[-2]) <= R[vt] * 15


@pytest.mark.parametrize("kind", ["ref_common", "empty_rows", "laplace_like"])
@pytest.mark.parametrize("slice_size,stride_factor", [(64, 1), (32, 4), (8, 2)])
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
def test_sellp_spmv(orc, cuda, kind, slice_size, stride_factor, vt, it):
    rng = np.random.default_rng(13)
    n, m, rp, ci, va = csr_case(rng, kind, vt, it)
    sets, lens, cols, vals = H.csr_to_sellp(rp, ci, va, n, slice_size, stride_factor)
    for nrhs in (1, 3):
        x = H.dense(rng, m, nrhs, vt=vt)
        y0 = H.dense(rng, n, nrhs, nrhs + 1, vt)
        a, b = both(orc, cuda, "sellp_spmv_%s_%s" % (vt, it),
                    lambda: [n, m, slice_size, sets, lens, cols, vals, x, nrhs, nrhs, y0.copy(),
                             nrhs + 1])
        assert np.array_equal(a[-2], b[-2])
        a, b = both(orc, cuda, "sellp_advanced_spmv_%s_%s" % (vt, it),
                    lambda: [n, m, slice_size, sets, lens, cols, vals, np.array([2.5], VT[vt]), x,
                             nrhs, nrhs, np.array([-0.5], VT[vt]), y0.copy(), nrhs + 1])
        assert np.array_equal(a[-2], b[-2])


@pytest.mark.parametrize("kind", ["ref_common", "empty_rows", "laplace_like", "long_rows",
                                  "all_empty"])
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("use_plan", [False, True])
def test_coo_spmv(orc, cuda, kind, vt, it, use_plan):
    rng = np.random.default_rng(17)
    n, m, rp, ci, va = csr_case(rng, kind, vt, it)
    nnz = len(va)
    rows = H.csr_to_coo_rows(rp, n, it)
    plan = cuda.make_coo_plan(vt, it, n, nnz, rows) if use_plan else None
    exact = kind != "long_rows"
    tol = R[vt] * max(1.0, np.sqrt(np.diff(rp.astype(np.int64)).max(initial=1)))
    for nrhs in (1, 3):
        x = H.dense(rng, m, nrhs, vt=vt)
        y0 = H.dense(rng, n, nrhs, vt=vt)
        al, be_ = np.array([-0.5], VT[vt]), np.array([2.0], VT[vt])
        for fname, mk in [
            ("coo_spmv", lambda: [n, m, nnz, rows, ci, va, x, nrhs, nrhs, y0.copy(), nrhs]),
            ("coo_advanced_spmv", lambda: [n, m, nnz, rows, ci, va, al, x, nrhs, nrhs, be_,
                                           y0.copy(), nrhs]),
            ("coo_spmv2", lambda: [n, m, nnz, rows, ci, va, x, nrhs, nrhs, y0.copy(), nrhs]),
            ("coo_advanced_spmv2", lambda: [n, m, nnz, rows, ci, va, al, x, nrhs, nrhs, y0.copy(),
                                            nrhs]),
        ]:
            a, b = both(orc, cuda, "%s_%s_%s" % (fname, vt, it), mk, plan)
            if exact:
                assert np.array_equal(a[-2], b[-2]), fname
            assert rel_err(a[-2], b[-2]) <= tol, fname


# ------------------------------------------------------------------------ dense BLAS-1
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols,stride", [(597, 43, 45), (1, 1, 1), (100003, 1, 1),
                                              (2_000_003, 1, 1), (4099, 1, 3), (5000, 70, 70),
                                              (0, 3, 3)])
def test_dense_reductions(orc, cuda, vt, rows, cols, stride):
    rng = np.random.default_rng(5)
    x = H.dense(rng, rows, cols, stride, vt)
    y = H.dense(rng, rows, cols, stride + 1, vt)
    # the ORACLE sums sequentially (error grows ~ n eps); the device tree sum is the more
    # accurate of the two, so the bound is the sequential one
    tol = R[vt] * max(1.0, rows / 64.0)
    for fname, args in [
        ("dense_compute_dot", lambda: [rows, cols, x, stride, y, stride + 1, np.zeros(cols, VT[vt])]),
        ("dense_compute_conj_dot", lambda: [rows, cols, x, stride, y, stride + 1,
                                            np.zeros(cols, VT[vt])]),
        ("dense_compute_norm2", lambda: [rows, cols, x, stride, np.zeros(cols, VT[vt])]),
        ("dense_compute_squared_norm2", lambda: [rows, cols, x, stride, np.zeros(cols, VT[vt])]),
    ]:
        a, b = both(orc, cuda, fname + "_" + vt, args)
        ref, got = a[-1].astype(np.float64), b[-1].astype(np.float64)
        scale = np.maximum(np.abs(ref), 1e-30)
        if "dot" in fname:  # cancellation: compare against sum |x||y|
            scale = (np.abs(x[:, :cols].astype(np.float64)) *
                     np.abs(y[:, :cols].astype(np.float64))).sum(0) + 1e-30
        assert (np.abs(ref - got) <= tol * scale).all(), fname
    # run-to-run determinism (fixed reduction tree, no atomics)
    if rows:
        r1, r2 = np.zeros(cols, VT[vt]), np.zeros(cols, VT[vt])
        cuda("dense_compute_dot_" + vt, rows, cols, x, stride, y, stride + 1, r1)
        cuda("dense_compute_dot_" + vt, rows, cols, x, stride, y, stride + 1, r2)
        assert np.array_equal(r1, r2)


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols,xs,ys", [(597, 43, 45, 46), (100001, 1, 1, 1), (33, 1, 2, 3)])
def test_dense_elementwise(orc, cuda, vt, rows, cols, xs, ys):
    rng = np.random.default_rng(6)
    x = H.dense(rng, rows, cols, xs, vt)
    y0 = H.dense(rng, rows, cols, ys, vt)
    a1 = np.array([0.37], VT[vt])
    a0 = np.array([0.0], VT[vt])
    ac = rng.uniform(-1, 1, cols).astype(VT[vt])
    ac[0] = 0.0
    for alpha, ncol in [(a1, 1), (a0, 1), (ac, cols)]:
        if ncol != 1 and cols == 1:
            continue
        for f in ("dense_add_scaled", "dense_sub_scaled"):
            a, b = both(orc, cuda, f + "_" + vt, lambda: [rows, cols, alpha, ncol, x, xs, y0.copy(), ys])
            assert np.array_equal(a[-2], b[-2]), f
        ynan = y0.copy()
        ynan[0, 0] = np.nan
        a, b = both(orc, cuda, "dense_scale_" + vt, lambda: [rows, cols, alpha, ncol, ynan.copy(), ys])
        assert np.array_equal(a[-2], b[-2], equal_nan=True)
        if (alpha != 0).all():
            a, b = both(orc, cuda, "dense_inv_scale_" + vt,
                        lambda: [rows, cols, alpha, ncol, y0.copy(), ys])
            assert np.array_equal(a[-2], b[-2])
    out = y0.copy()
    cuda("dense_copy_" + vt, rows, cols, x, xs, out, ys)
    assert np.array_equal(out[:, :cols], x[:, :cols]) and np.array_equal(out[:, cols:], y0[:, cols:])
    cuda("dense_fill_" + vt, rows, cols, out, ys, 2.5)
    assert (out[:, :cols] == 2.5).all() and np.array_equal(out[:, cols:], y0[:, cols:])


# ------------------------------------------------------------- CG / BiCGStab step kernels
def solver_vectors(rng, vt, rows, cols, names, strides):
    return {nm: H.dense(rng, rows, cols, st, vt) for nm, st in zip(names, strides)}


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols", [(597, 43), (100001, 1)])
def test_cg_steps(orc, cuda, vt, rows, cols):
    # test/solver/cg_kernels.cpp:42-76: padded strides, a zero prev_rho column, a stopped column
    rng = np.random.default_rng(8)
    st = dict(b=cols + 1, r=cols + 2, z=cols + 3, p=cols + 2, q=cols + 1, x=cols)
    v = {k: H.dense(rng, rows, cols, s, vt) for k, s in st.items()}
    rho = rng.uniform(0.5, 1, cols).astype(VT[vt])
    prev_rho = rng.uniform(0.5, 1, cols).astype(VT[vt])
    beta = rng.uniform(0.5, 1, cols).astype(VT[vt])
    stop = np.zeros(cols, dtype=np.uint8)
    if cols > 3:
        prev_rho[2] = 0
        beta[3] = 0
        stop[1] = 1 | 0x40
    a, b = both(orc, cuda, "cg_initialize_" + vt,
                lambda: [rows, cols, v["b"], st["b"], v["r"].copy(), st["r"], v["z"].copy(), st["z"],
                         v["p"].copy(), st["p"], v["q"].copy(), st["q"], prev_rho.copy(), rho.copy(),
                         np.full(cols, 0x81, np.uint8)])
    for i in range(len(a)):
        if isinstance(a[i], np.ndarray):
            assert np.array_equal(a[i], b[i])
    a, b = both(orc, cuda, "cg_step_1_" + vt,
                lambda: [rows, cols, v["p"].copy(), st["p"], v["z"], st["z"], rho, prev_rho, stop])
    assert np.array_equal(a[2], b[2])
    a, b = both(orc, cuda, "cg_step_2_" + vt,
                lambda: [rows, cols, v["x"].copy(), st["x"], v["r"].copy(), st["r"], v["p"], st["p"],
                         v["q"], st["q"], beta, rho, stop])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[4], b[4])


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols", [(597, 43), (50001, 1), (0, 2)])
def test_bicgstab_steps(orc, cuda, vt, rows, cols):
    rng = np.random.default_rng(9)
    names = ["b", "r", "rr", "y", "s", "t", "z", "v", "p", "x"]
    st = {nm: cols + (i % 3) for i, nm in enumerate(names)}
    v = {nm: H.dense(rng, rows, cols, st[nm], vt) for nm in names}
    sc = {nm: rng.uniform(0.5, 1, cols).astype(VT[vt])
          for nm in ["prev_rho", "rho", "alpha", "beta", "gamma", "omega"]}
    stop = np.zeros(cols, dtype=np.uint8)
    if cols > 4:
        sc["prev_rho"][2] = 0
        sc["beta"][3] = 0
        stop[1] = 1 | 0x40
        stop[4] = 2  # stopped, not finalized
    a, b = both(orc, cuda, "bicgstab_initialize_" + vt, lambda: [
        rows, cols, v["b"], st["b"], v["r"].copy(), st["r"], v["rr"].copy(), st["rr"], v["y"].copy(),
        st["y"], v["s"].copy(), st["s"], v["t"].copy(), st["t"], v["z"].copy(), st["z"],
        v["v"].copy(), st["v"], v["p"].copy(), st["p"], sc["prev_rho"].copy(), sc["rho"].copy(),
        sc["alpha"].copy(), sc["beta"].copy(), sc["gamma"].copy(), sc["omega"].copy(),
        np.full(cols, 0x81, np.uint8)])
    for i in range(len(a)):
        if isinstance(a[i], np.ndarray):
            assert np.array_equal(a[i], b[i]), i
    a, b = both(orc, cuda, "bicgstab_step_1_" + vt, lambda: [
        rows, cols, v["r"], st["r"], v["p"].copy(), st["p"], v["v"], st["v"], sc["rho"],
        sc["prev_rho"], sc["alpha"], sc["omega"], stop])
    assert np.array_equal(a[4], b[4])
    a, b = both(orc, cuda, "bicgstab_step_2_" + vt, lambda: [
        rows, cols, v["r"], st["r"], v["s"].copy(), st["s"], v["v"], st["v"], sc["rho"],
        sc["alpha"].copy(), sc["beta"], stop])
    assert np.array_equal(a[4], b[4]) and np.array_equal(a[9], b[9])
    a, b = both(orc, cuda, "bicgstab_step_3_" + vt, lambda: [
        rows, cols, v["x"].copy(), st["x"], v["r"].copy(), st["r"], v["s"], st["s"], v["t"], st["t"],
        v["y"], st["y"], v["z"], st["z"], sc["alpha"], sc["beta"], sc["gamma"], sc["omega"].copy(),
        stop])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[4], b[4]) and np.array_equal(a[17], b[17])
    a, b = both(orc, cuda, "bicgstab_finalize_" + vt, lambda: [
        rows, cols, v["x"].copy(), st["x"], v["y"], st["y"], sc["alpha"], stop.copy()])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[7], b[7])


# ------------------------------------------------------------------------------- GMRES
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols,kd", [(597, 3, 5), (30011, 1, 30)])
def test_gmres_kernels(orc, cuda, vt, rows, cols, kd):
    rng = np.random.default_rng(10)
    b = H.dense(rng, rows, cols, cols + 1, vt)
    stop = np.zeros(cols, dtype=np.uint8)
    a, bb = both(orc, cuda, "common_gmres_initialize_" + vt, lambda: [
        rows, cols, kd, b, cols + 1, H.dense(rng, rows, cols, cols, vt), cols,
        np.full((kd, cols), 7.0, VT[vt]), cols, np.full((kd, cols), 7.0, VT[vt]), cols,
        np.full(cols, 0x81, np.uint8)])
    for i in range(len(a)):
        if isinstance(a[i], np.ndarray) and i != 5:
            assert np.array_equal(a[i], bb[i])
    assert np.array_equal(a[5], bb[5])
    res = a[5]
    rn = np.sqrt((res.astype(np.float64) ** 2).sum(0)).astype(VT[vt])
    kb0 = np.full(((kd + 1) * rows, cols), 3.0, VT[vt])
    a, bb = both(orc, cuda, "gmres_restart_" + vt, lambda: [
        rows, cols, res, cols, rn, np.zeros((kd + 1, cols), VT[vt]), kb0.copy(), cols,
        np.full(cols, 99, np.uint64)])
    assert np.array_equal(a[5], bb[5]) and np.array_equal(a[6], bb[6]) and np.array_equal(a[8], bb[8])
    # a few random "basis" vectors for multi_dot / multi_axpy
    nb = min(kd, 7)
    kb = rng.uniform(-1, 1, ((kd + 1) * rows, cols)).astype(VT[vt])
    w = H.dense(rng, rows, cols, cols, vt)
    a, bb = both(orc, cuda, "gmres_multi_dot_" + vt, lambda: [
        rows, cols, nb, kb, cols, w, cols, np.zeros((nb + 1, cols), VT[vt]), cols])
    assert rel_err(a[7], bb[7]) <= R[vt] * np.sqrt(rows)
    y = rng.uniform(-1, 1, (kd, cols)).astype(VT[vt])
    fin = np.full(cols, nb, np.uint64)
    st2 = stop.copy()
    st2[0] = 5  # stopped, not finalized -> gets finalized
    a, bb = both(orc, cuda, "gmres_multi_axpy_" + vt, lambda: [
        rows, cols, kb, cols, y, cols, np.zeros((rows, cols), VT[vt]), cols, fin, st2.copy()])
    assert np.array_equal(a[6], bb[6]) and np.array_equal(a[9], bb[9])
    # hessenberg_qr over several iterations + solve_krylov, chained on both backends
    hs = (kd + 1) * cols
    st_o = {"gsin": np.zeros((kd, cols), VT[vt]), "gcos": np.zeros((kd, cols), VT[vt]),
            "rn": rn.copy(), "rnc": np.zeros((kd + 1, cols), VT[vt]),
            "hess": np.zeros((kd, hs), VT[vt]), "fin": np.zeros(cols, np.uint64)}
    st_o["rnc"][0] = rn
    st_c = {k: v.copy() for k, v in st_o.items()}
    for it_ in range(min(kd, 6)):
        col = rng.uniform(-1, 1, (it_ + 2, cols)).astype(VT[vt])
        for be_, S in ((orc, st_o), (cuda, st_c)):
            hit = S["hess"][it_].reshape(-1)[: (it_ + 2) * cols].reshape(it_ + 2, cols)
            tmp = col.copy()
            be_("common_gmres_hessenberg_qr_" + vt, cols, S["gsin"], cols, S["gcos"], cols, S["rn"],
                S["rnc"], cols, tmp, cols, it_, S["fin"], stop)
            hit[...] = tmp
    for k in st_o:
        assert rel_err(st_o[k], st_c[k]) <= R[vt] * 4, k
    yo, yc = np.zeros((kd, cols), VT[vt]), np.zeros((kd, cols), VT[vt])
    orc("common_gmres_solve_krylov_" + vt, cols, st_o["rnc"], cols, st_o["hess"], hs, yo, cols,
        st_o["fin"], stop)
    cuda("common_gmres_solve_krylov_" + vt, cols, st_o["rnc"], cols, st_o["hess"], hs, yc, cols,
         st_o["fin"], stop)
    assert rel_err(yo, yc) <= R[vt] * 50


# ------------------------------------------------------------------------- stop, Jacobi
@pytest.mark.parametrize("vt", VTS)
def test_set_all_statuses_and_residual_norm(orc, cuda, vt):
    rng = np.random.default_rng(12)
    cols = 300
    stop = rng.integers(0, 3, cols).astype(np.uint8)
    a, b = both(orc, cuda, "set_all_statuses", lambda: [cols, 5, 1, stop.copy()])
    assert np.array_equal(a[-1], b[-1])
    tau = rng.uniform(0, 2, cols).astype(VT[vt])
    orig = np.ones(cols, VT[vt])
    for goal in (0.0, 1.0, 3.0):
        oa, ob = [H.OutInt(), H.OutInt()], [H.OutInt(), H.OutInt()]
        s1, s2 = stop.copy(), stop.copy()
        orc("residual_norm_" + vt, cols, tau, orig, goal, 3, 0, s1, np.zeros(2, np.uint8), *oa)
        cuda("residual_norm_" + vt, cols, tau, orig, goal, 3, 0, s2, np.zeros(2, np.uint8), *ob)
        assert np.array_equal(s1, s2)
        assert [o.value for o in oa] == [o.value for o in ob]


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("max_bs", [1, 3, 16, 32, 13])
def test_block_jacobi_apply(orc, cuda, vt, it, max_bs):
    # test/preconditioner/jacobi_kernels.cpp:426-660: random block sizes up to max_bs, multi rhs
    rng = np.random.default_rng(14)
    nblocks = 257
    sizes = rng.integers(1, max_bs + 1, nblocks)
    sizes[0] = max_bs
    ptrs = np.zeros(nblocks + 1, dtype=IT[it])
    ptrs[1:] = np.cumsum(sizes)
    n = int(ptrs[-1])
    # storage scheme: include/ginkgo/core/preconditioner/jacobi.hpp:589-625
    pow2 = 1
    while pow2 < max_bs:
        pow2 *= 2
    group_size = 32 // pow2
    gp = int(np.log2(group_size))
    bo = max_bs
    stride = bo << gp
    go = max_bs * stride
    blocks = rng.uniform(-1, 1, go * ((nblocks + group_size - 1) // group_size)).astype(VT[vt])
    for nrhs, bs, xs in [(1, 1, 1), (3, 4, 5)]:
        b = H.dense(rng, n, nrhs, bs, vt)
        x0 = H.dense(rng, n, nrhs, xs, vt)
        a, c = both(orc, cuda, "jacobi_simple_apply_%s_%s" % (vt, it), lambda: [
            nblocks, max_bs, bo, go, gp, ptrs, blocks, b, bs, nrhs, x0.copy(), xs])
        assert np.array_equal(a[-2], c[-2])
        a, c = both(orc, cuda, "jacobi_apply_%s_%s" % (vt, it), lambda: [
            nblocks, max_bs, bo, go, gp, ptrs, blocks, np.array([1.5], VT[vt]), b, bs, nrhs,
            np.array([-0.5], VT[vt]), x0.copy(), xs])
        assert np.array_equal(a[-2], c[-2])


# ----------------------------------------------------------- integer-exact conversions
@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("kind", ["ref_common", "empty_rows", "all_empty", "long_rows"])
def test_index_conversions_bit_exact(orc, cuda, it, kind):
    rng = np.random.default_rng(19)
    n, m, rp, ci, va = csr_case(rng, kind, "f64", it)
    nnz = len(va)
    a, b = both(orc, cuda, "convert_ptrs_to_idxs_" + it, lambda: [rp, n, np.full(nnz, -7, IT[it])])
    assert np.array_equal(a[-1], b[-1])
    rows = a[-1]
    a, b = both(orc, cuda, "convert_idxs_to_ptrs_" + it,
                lambda: [rows, nnz, n, np.full(n + 1, -7, IT[it])])
    assert np.array_equal(a[-1], b[-1]) and np.array_equal(b[-1], rp)


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
def test_extract_diagonal(orc, cuda, vt, it):
    rng = np.random.default_rng(20)
    n, m, rp, ci, va = csr_case(rng, "empty_rows", vt, it)
    a, b = both(orc, cuda, "csr_extract_diagonal_%s_%s" % (vt, it),
                lambda: [n, rp, ci, va, np.full(n, np.nan, VT[vt])])
    assert np.array_equal(a[-1], b[-1])


@pytest.mark.parametrize("vt", VTS)
def test_hybrid_is_ell_plus_coo(orc, cuda, vt):
    """Hybrid::apply = ell->apply then coo->apply2 (core/matrix/hybrid.cpp:175-201): split every
    row into its first k entries (ELL part) and the rest (COO part)"""
    rng = np.random.default_rng(23)
    n, m, rp, ci, va = csr_case(rng, "ref_common", vt, "i32")
    k = 6
    lens = np.diff(rp)
    ell_mask = np.concatenate([np.arange(l) < k for l in lens])
    erp = np.zeros(n + 1, np.int32)
    erp[1:] = np.cumsum(np.minimum(lens, k))
    width, stride, ecols, evals = H.csr_to_ell(erp, ci[ell_mask], va[ell_mask], n)
    rows = H.csr_to_coo_rows(rp, n, "i32")[~ell_mask]
    ccols, cvals = ci[~ell_mask], va[~ell_mask]
    x = H.dense(rng, m, 1, vt=vt)
    y = np.zeros((n, 1), VT[vt])
    cuda("ell_spmv_%s_i32" % vt, n, m, width, stride, ecols, evals, x, 1, 1, y, 1)
    cuda("coo_spmv2_%s_i32" % vt, n, m, len(cvals), rows, ccols, cvals, x, 1, 1, y, 1)
    yo = np.zeros((n, 1), VT[vt])
    orc("csr_spmv_%s_i32" % vt, n, m, len(va), rp, ci, va, x, 1, 1, yo, 1)
    # same left-to-right order: ELL part first, COO part appended
    assert np.array_equal(y, yo)


# ------------------------------------------- CSR -> ELL / SELL-P / Hybrid, sort (8f-1)
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("kind", ["ref_common", "empty_rows", "all_empty", "long_rows"])
def test_convert_to_ell_bit_exact(orc, cuda, vt, it, kind):
    rng = np.random.default_rng(31)
    n, m, rp, ci, va = csr_case(rng, kind, vt, it)
    mo, mc = H.OutI64(), H.OutI64()
    orc("ell_compute_max_row_nnz_" + it, rp, n, mo)
    cuda("ell_compute_max_row_nnz_" + it, rp, n, mc)
    assert mo.value == mc.value == (int(np.diff(rp).max()) if n else 0)
    width, stride = mo.value, n + 3  # padded stride: rows >= n must stay untouched
    a, b = both(orc, cuda, "csr_convert_to_ell_%s_%s" % (vt, it),
                lambda: [n, rp, ci, va, width, stride, np.full(width * stride, 77, IT[it]),
                         np.full(width * stride, 7, VT[vt])])
    assert np.array_equal(a[-1], b[-1]) and np.array_equal(a[-2], b[-2])


@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("kind", ["ref_common", "empty_rows", "all_empty", "long_rows"])
@pytest.mark.parametrize("slice_size,stride_factor", [(64, 1), (32, 4), (8, 3)])
def test_convert_to_sellp_bit_exact(orc, cuda, it, kind, slice_size, stride_factor):
    vt = "f64"
    rng = np.random.default_rng(32)
    n, m, rp, ci, va = csr_case(rng, kind, vt, it)
    ns = (n + slice_size - 1) // slice_size
    a, b = both(orc, cuda, "sellp_compute_slice_sets_" + it,
                lambda: [rp, n, slice_size, stride_factor, np.full(ns + 1, 9, np.uint64),
                         np.full(max(ns, 1), 9, np.uint64)])
    assert np.array_equal(a[-2], b[-2]) and np.array_equal(a[-1][:ns], b[-1][:ns])
    ss, sl = a[-2], a[-1]
    tot = int(ss[-1]) * slice_size
    a, b = both(orc, cuda, "csr_convert_to_sellp_%s_%s" % (vt, it),
                lambda: [n, slice_size, ss, sl, rp, ci, va, np.full(max(tot, 1), 77, IT[it]),
                         np.full(max(tot, 1), 7, VT[vt])])
    assert np.array_equal(a[-1], b[-1]) and np.array_equal(a[-2], b[-2])


@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("kind", ["ref_common", "empty_rows", "long_rows"])
@pytest.mark.parametrize("strategy", [(0, 0, 0, 0), (1, 3, 0, 0), (2, 0, 0.8, 0), (3, 0, 0.5, 0.01),
                                      (4, 0, 0, 0)])
def test_convert_to_hybrid_bit_exact(orc, cuda, it, kind, strategy):
    vt = "f64"
    rng = np.random.default_rng(33)
    n, m, rp, ci, va = csr_case(rng, kind, vt, it)
    if it == "i64":
        pytest.skip("order statistic helper is exercised with i32 row pointers")
    lim_o = H.hybrid_ell_lim(orc, rp, n, m, *strategy, vbytes=8, ibytes=4)
    lim_c = H.hybrid_ell_lim(cuda, rp, n, m, *strategy, vbytes=8, ibytes=4)
    assert lim_o == lim_c
    a, b = both(orc, cuda, "csr_compute_hybrid_coo_row_ptrs_" + it,
                lambda: [rp, n, lim_o, np.full(n + 1, -5, np.int64)])
    assert np.array_equal(a[-1], b[-1])
    crp = a[-1]
    cn, stride = int(crp[-1]), n + 2
    a, b = both(orc, cuda, "csr_convert_to_hybrid_%s_%s" % (vt, it),
                lambda: [n, rp, ci, va, lim_o, stride, np.full(max(lim_o * stride, 1), 77, IT[it]),
                         np.full(max(lim_o * stride, 1), 7, VT[vt]), crp,
                         np.full(max(cn, 1), -3, IT[it]), np.full(max(cn, 1), -3, IT[it]),
                         np.full(max(cn, 1), 5, VT[vt])])
    for i in (6, 7, 9, 10, 11):  # ell cols / vals, coo rows / cols / vals
        assert np.array_equal(a[i], b[i])


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("kind", ["ref_common", "empty_rows", "long_rows", "very_long"])
def test_sort_by_column_index_bit_exact(orc, cuda, vt, it, kind):
    rng = np.random.default_rng(34)
    if kind == "very_long":  # rows beyond the shared-memory path (> 2048 entries)
        n, m = 40, 9000
        lens = rng.integers(0, 50, n)
        lens[[3, 17]] = [2500, 5000]
        rp, ci, va = H.random_csr(rng, n, m, lens, vt, it)
    else:
        n, m, rp, ci, va = csr_case(rng, kind, vt, it)
    ci2, va2 = ci.copy(), va.copy()
    for r in range(n):
        s, e = int(rp[r]), int(rp[r + 1])
        perm = rng.permutation(e - s)
        ci2[s:e], va2[s:e] = ci[s:e][perm], va[s:e][perm]
    a, b = both(orc, cuda, "csr_sort_by_column_index_%s_%s" % (vt, it),
                lambda: [n, rp, ci2.copy(), va2.copy()])
    assert np.array_equal(a[-2], b[-2]) and np.array_equal(a[-1], b[-1])
    for r in range(n):  # independent truth: columns are distinct, so the order is unique
        s, e = int(rp[r]), int(rp[r + 1])
        o = np.argsort(ci2[s:e], kind="stable")
        assert np.array_equal(b[-2][s:e], ci2[s:e][o]) and np.array_equal(b[-1][s:e], va2[s:e][o])


# ------------------------------------------------- block-Jacobi generate on the device (8f-2)
@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("max_bs,singular", [(1, False), (2, False), (7, False), (16, False),
                                            (32, False), (13, True)])
def test_jacobi_generate_bit_exact(orc, cuda, vt, it, max_bs, singular):
    rng = np.random.default_rng(60 + max_bs)
    n = 700
    sizes = []
    while sum(sizes) < n:
        sizes.append(int(rng.integers(1, max_bs + 1)))
    sizes[-1] -= sum(sizes) - n
    if sizes[-1] == 0:
        sizes.pop()
    bp = np.concatenate([[0], np.cumsum(sizes)]).astype(IT[it])
    rp, ci, va = H.random_csr(rng, n, n, rng.integers(3, 40, n), vt, it)
    if singular:
        b0, b1 = int(bp[3]), int(bp[4])
        for r in range(b0, b1):
            s, e = int(rp[r]), int(rp[r + 1])
            va[s:e][(ci[s:e] >= b0) & (ci[s:e] < b1)] = 0
    pow2 = 1
    while pow2 < max_bs:
        pow2 *= 2
    group_size = 32 // pow2
    gp = group_size.bit_length() - 1
    block_offset, group_offset = max_bs, max_bs * group_size * max_bs
    nb = len(bp) - 1
    space = (nb + group_size - 1) // group_size * group_offset
    a, b = both(orc, cuda, "jacobi_generate_%s_%s" % (vt, it),
                lambda: [n, rp, ci, va, nb, max_bs, block_offset, group_offset, gp, bp,
                         np.zeros(space, VT[vt])])
    assert np.array_equal(a[-1], b[-1], equal_nan=True)
    if not singular:  # and it is an inverse: apply to A_block * e
        k = nb // 2
        bs, st = int(bp[k + 1] - bp[k]), int(bp[k])
        off = group_offset * (k >> gp) + block_offset * (k & ((1 << gp) - 1))
        stride = block_offset << gp
        inv = b[-1][off + np.arange(bs)[:, None] + np.arange(bs)[None, :] * stride].astype(np.float64)
        blk = np.zeros((bs, bs))
        for r in range(bs):
            for p in range(int(rp[st + r]), int(rp[st + r + 1])):
                c = int(ci[p]) - st
                if 0 <= c < bs:
                    blk[r, c] = va[p]
        err = np.abs(inv @ blk - np.eye(bs)).max()
        assert err < (1e-9 if vt == "f64" else 1e-2) * max(1.0, np.linalg.cond(blk))


@pytest.mark.parametrize("it", ITS)
@pytest.mark.parametrize("max_bs", [1, 2, 5, 16, 32])
@pytest.mark.parametrize("n,max_run", [(1, 1), (5000, 1), (5000, 3), (5000, 40), (70000, 7)])
def test_find_blocks_bit_exact(orc, cuda, it, max_bs, n, max_run):
    rng = np.random.default_rng(80 + max_bs + max_run)
    cols_of = []
    r = 0
    while r < n:
        run = int(rng.integers(1, max_run + 1))
        k = int(rng.integers(1, 9))
        cols = np.sort(rng.choice(n, size=min(k, n), replace=False))
        for _ in range(min(run, n - r)):
            cols_of.append(cols)
            r += 1
    rp = np.zeros(n + 1, IT[it])
    rp[1:] = np.cumsum([len(c) for c in cols_of])
    ci = np.concatenate(cols_of).astype(IT[it])
    no, nc = H.OutI64(), H.OutI64()
    bo, bc = np.full(n + 1, -1, IT[it]), np.full(n + 1, -1, IT[it])
    orc("jacobi_find_blocks_" + it, n, rp, ci, max_bs, bo, no)
    cuda("jacobi_find_blocks_" + it, n, rp, ci, max_bs, bc, nc)
    assert no.value == nc.value
    assert np.array_equal(bo[:no.value + 1], bc[:nc.value + 1])
    assert bc[nc.value] == n and np.all(np.diff(bc[:nc.value + 1]) <= max_bs)


# ------------------------------------------------------- sibling Krylov kernels (8f-3)
def _all_equal(a, b):
    for i in range(len(a)):
        if isinstance(a[i], np.ndarray):
            assert np.array_equal(a[i], b[i]), i


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols", [(597, 43), (100001, 1), (0, 2)])
def test_fcg_steps(orc, cuda, vt, rows, cols):
    rng = np.random.default_rng(90)
    st = dict(b=cols + 1, r=cols + 2, z=cols + 3, p=cols + 2, q=cols + 1, t=cols + 4, x=cols)
    v = {k: H.dense(rng, rows, cols, s, vt) for k, s in st.items()}
    sc = {k: rng.uniform(0.5, 1, cols).astype(VT[vt]) for k in ("rho", "prev_rho", "rho_t", "beta")}
    stop = np.zeros(cols, dtype=np.uint8)
    if cols > 3:
        sc["prev_rho"][2] = 0
        sc["beta"][3] = 0
        stop[1] = 1 | 0x40
    a, b = both(orc, cuda, "fcg_initialize_" + vt,
                lambda: [rows, cols, v["b"], st["b"], v["r"].copy(), st["r"], v["z"].copy(), st["z"],
                         v["p"].copy(), st["p"], v["q"].copy(), st["q"], v["t"].copy(), st["t"],
                         sc["prev_rho"].copy(), sc["rho"].copy(), sc["rho_t"].copy(),
                         np.full(cols, 0x81, np.uint8)])
    _all_equal(a, b)
    a, b = both(orc, cuda, "fcg_step_1_" + vt,
                lambda: [rows, cols, v["p"].copy(), st["p"], v["z"], st["z"], sc["rho_t"],
                         sc["prev_rho"], stop])
    _all_equal(a, b)
    a, b = both(orc, cuda, "fcg_step_2_" + vt,
                lambda: [rows, cols, v["x"].copy(), st["x"], v["r"].copy(), st["r"], v["t"].copy(),
                         st["t"], v["p"], st["p"], v["q"], st["q"], sc["beta"], sc["rho"], stop])
    _all_equal(a, b)


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols", [(597, 43), (100001, 1), (0, 2)])
def test_cgs_steps(orc, cuda, vt, rows, cols):
    rng = np.random.default_rng(91)
    names = ("b", "r", "r_tld", "p", "q", "u", "u_hat", "v_hat", "t", "x")
    st = {k: cols + (i % 4) for i, k in enumerate(names)}
    v = {k: H.dense(rng, rows, cols, s, vt) for k, s in st.items()}
    sc = {k: rng.uniform(0.5, 1, cols).astype(VT[vt])
          for k in ("alpha", "beta", "gamma", "prev_rho", "rho")}
    stop = np.zeros(cols, dtype=np.uint8)
    if cols > 3:
        sc["prev_rho"][2] = 0
        sc["gamma"][3] = 0
        stop[1] = 1 | 0x40
    a, b = both(orc, cuda, "cgs_initialize_" + vt,
                lambda: [rows, cols, v["b"], st["b"]] + sum(
                    [[v[k].copy(), st[k]] for k in ("r", "r_tld", "p", "q", "u", "u_hat", "v_hat", "t")],
                    []) + [sc[k].copy() for k in ("alpha", "beta", "gamma", "prev_rho", "rho")] +
                [np.full(cols, 0x81, np.uint8)])
    _all_equal(a, b)
    a, b = both(orc, cuda, "cgs_step_1_" + vt,
                lambda: [rows, cols, v["r"], st["r"], v["u"].copy(), st["u"], v["p"].copy(), st["p"],
                         v["q"], st["q"], sc["beta"].copy(), sc["rho"], sc["prev_rho"], stop])
    _all_equal(a, b)
    a, b = both(orc, cuda, "cgs_step_2_" + vt,
                lambda: [rows, cols, v["u"], st["u"], v["v_hat"], st["v_hat"], v["q"].copy(), st["q"],
                         v["t"].copy(), st["t"], sc["alpha"].copy(), sc["rho"], sc["gamma"], stop])
    _all_equal(a, b)
    a, b = both(orc, cuda, "cgs_step_3_" + vt,
                lambda: [rows, cols, v["t"], st["t"], v["u_hat"], st["u_hat"], v["r"].copy(), st["r"],
                         v["x"].copy(), st["x"], sc["alpha"], stop])
    _all_equal(a, b)


def test_snapshot_reads_the_state_at_its_point_of_the_stream(cuda):
    """b200_snapshot_begin / end (the fused solvers' control-block poll): the bytes are taken IN STREAM
    ORDER at the begin, work enqueued afterwards neither delays nor changes what end hands out"""
    if not hasattr(cuda, "ctx"):
        pytest.skip("harness self-check: no device")
    import torch
    with torch.cuda.stream(cuda.stream):
        t = torch.arange(8, dtype=torch.int32, device="cuda")
        cuda._libmod.check(cuda.l.b200_snapshot_begin(cuda.ctx, 0, t.data_ptr(), 32))
        t.add_(100)
        big = torch.zeros(1 << 26, dtype=torch.float32, device="cuda")
        for _ in range(8):  # keep the stream busy behind the first snapshot
            big.add_(1.0)
        cuda._libmod.check(cuda.l.b200_snapshot_begin(cuda.ctx, 1, t.data_ptr(), 32))
        t.add_(100)
    a, b = np.zeros(8, np.int32), np.zeros(8, np.int32)
    cuda._libmod.check(cuda.l.b200_snapshot_end(cuda.ctx, 0, a.ctypes.data, 32))
    cuda._libmod.check(cuda.l.b200_snapshot_end(cuda.ctx, 1, b.ctypes.data, 32))
    assert np.array_equal(a, np.arange(8)) and np.array_equal(b, np.arange(8) + 100)
    cuda.stream.synchronize()
    assert int(t[0].item()) == 200
