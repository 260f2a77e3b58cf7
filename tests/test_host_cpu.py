"""The C++ HOST layer (ginkgo_b200/host/*.hpp: solver loops, stopping criteria, Jacobi set-up,
format conversions, file I/O, staging) exercised WITHOUT a GPU: capi.cpp is compiled together
with tests/mock (a host-memory stand-in for the C ABI whose entry points forward to the
oracle's restatement of the same reference kernels -- test infrastructure only).  Because the
kernels underneath are the oracle's, every result must be BIT-IDENTICAL to the oracle's own
restatement of the reference host loops (which tests/test_oracle_vs_ref.py pins to the real
reference): any difference is a bug in the C++ host logic."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import workloads as W
from tests import helpers as H
from tests.helpers import VT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _CpuExec:
    """stands where api.HostExecutor stands: a gkob executor handle on the mock"""

    def __init__(self, lib):
        self.h = lib.gkob_exec_create(0, None)
        assert self.h
        self.device = torch.device("cpu")
        self.stream = None
        self._lib = lib

    def synchronize(self):
        pass


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    from ginkgo_b200 import api
    d = str(tmp_path_factory.mktemp("mock"))
    inc = os.path.join(ROOT, "include")
    gen = os.path.join(d, "mock_gen.c")
    subprocess.run(["python", os.path.join(ROOT, "tests", "mock", "gen_mock.py"),
                    os.path.join(inc, "ginkgo_b200.h"), os.path.join(ROOT, "oracle", "liboracle.so"),
                    os.path.join(ROOT, "tests", "mock", "mock_base.c"), gen], check=True,
                   capture_output=True)
    objs = []
    for src in (os.path.join(ROOT, "tests", "mock", "mock_base.c"), gen):
        o = os.path.join(d, os.path.basename(src) + ".o")
        subprocess.run(["gcc", "-O1", "-fPIC", "-I" + inc, "-c", src, "-o", o], check=True)
        objs.append(o)
    so = os.path.join(d, "libgko_b200_host_mock.so")
    # one DSO, -Bsymbolic: the b200_* references of the host layer bind to the mock inside it,
    # whatever else the process has loaded
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-o", so,
                    os.path.join(ROOT, "ginkgo_b200", "host", "capi.cpp")] + objs +
                   ["-L" + os.path.join(ROOT, "oracle"), "-loracle",
                    "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    lib = api._configure_host_lib(ctypes.CDLL(so))
    saved = api._HOST_LIB
    api._HOST_LIB = lib
    yield _CpuExec(lib)
    api._HOST_LIB = saved


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def host_solve(host, kind, vt, rp, ci, va, b, x0, precond_max_bs=0, block_ptrs=None, **kw):
    from ginkgo_b200 import api
    t = [_t(va), _t(ci), _t(rp)]
    tb, tx = _t(b), _t(x0).clone()
    n = len(rp) - 1
    A = api.host_csr(host, (n, n), *t)
    s = api.HostSolver(host, kind, A, precond_max_bs=precond_max_bs, block_ptrs=block_ptrs,
                       fused=False, **kw)
    s.apply(api.host_dense(host, tb), api.host_dense(host, tx))
    return tx.numpy(), s.num_iterations, s.stop_status


@pytest.mark.parametrize("kind", ["cg", "fcg", "cgs", "pipe_cg", "minres", "bicgstab", "gmres", "gcr"])
@pytest.mark.parametrize("precond", [0, 1, 2, 3])
@pytest.mark.parametrize("vt", ["f64", "f32"])
def test_host_solver_loops_are_the_oracle_loops(host, kind, precond, vt):
    """precond 3 = block Jacobi with blocks detected by find_blocks (no block pointers)"""
    from oracle import ref
    rp, ci, va = W.laplace(20, 2, vdtype=VT[vt])
    n = len(rp) - 1
    rng = np.random.default_rng(5)
    b = rng.uniform(-1, 1, (n, 2)).astype(VT[vt])
    x0 = np.zeros((n, 2), VT[vt])
    red = 1e-9 if vt == "f64" else 1e-4
    max_bs = {0: 0, 1: 1, 2: 8, 3: 4}[precond]
    bp = np.arange(0, n + 1, 8, dtype=np.int32) if precond == 2 else None
    jac = None
    if precond:
        if not ref.available():
            pytest.skip("needs oracle/_ref for the inverted blocks")
        jac = ref.jacobi_generate(rp, ci, va, max_bs, bp)
    for iter_first in (1, 0):
        xo, ito, stop_o = H.orc_solve(kind, vt, rp, ci, va, b, x0, min(precond, 2), jac, max_iters=300,
                                      reduction=red, iter_first=iter_first, krylov_dim=15)
        xh, ith, stop_h = host_solve(host, kind, vt, rp, ci, va, b, x0, max_bs, bp, max_iters=300,
                                     reduction=red, iter_first=bool(iter_first), krylov_dim=15)
        assert ith == ito
        assert stop_h == stop_o[0]
        assert np.array_equal(xh, xo)


@pytest.mark.parametrize("res_kind,baseline", [(1, 0), (1, 1), (1, 2), (2, 0), (2, 1)])
def test_host_criteria(host, res_kind, baseline):
    rp, ci, va = W.laplace(16, 2)
    n = len(rp) - 1
    b = np.ones((n, 1))
    x0 = np.full((n, 1), 0.5)
    kw = dict(max_iters=200, res_kind=res_kind, baseline=baseline, reduction=1e-6 if baseline != 2 else 1e-5,
              iter_first=1, krylov_dim=10)
    xo, ito, stop_o = H.orc_solve("cg", "f64", rp, ci, va, b, x0, 0, None, **kw)
    kw["iter_first"] = True
    xh, ith, stop_h = host_solve(host, "cg", "f64", rp, ci, va, b, x0, **kw)
    assert (ith, stop_h) == (ito, stop_o[0]) and np.array_equal(xh, xo)


@pytest.mark.parametrize("vt", ["f64", "f32"])
@pytest.mark.parametrize("fmt,kw", [("ell", {}), ("sellp", {}), ("sellp", dict(slice_size=32, stride_factor=4)),
                                    ("coo", {}), ("hybrid", {}),
                                    ("hybrid", dict(strategy="column_limit", columns=3)),
                                    ("hybrid", dict(strategy="imbalance_limit", percent=0.5)),
                                    ("hybrid", dict(strategy="minimal_storage_limit"))])
def test_host_convert_then_apply(host, orc, vt, fmt, kw):
    from ginkgo_b200 import api
    rng = np.random.default_rng(41)
    n, m = 700, 600
    rp, ci, va = H.random_csr(rng, n, m, rng.integers(0, 13, n), vt, "i32")
    x = rng.uniform(-1, 1, m).astype(VT[vt])
    A = api.host_csr(host, (n, m), _t(va), _t(ci), _t(rp))
    B = api.host_convert(A, fmt, **kw)
    tx, ty = _t(x), torch.zeros(n, dtype=_t(x).dtype)
    xd, yd = api.host_dense(host, tx), api.host_dense(host, ty)
    api._hcheck(api._host().gkob_apply(B.h, xd.h, yd.h))
    yo = np.zeros(n, VT[vt])
    orc("csr_spmv_%s_i32" % vt, n, m, len(va), rp, ci, va, x, 1, 1, yo, 1)
    assert np.array_equal(ty.numpy(), yo)
    # advanced apply through the converted operator
    al, be = torch.tensor([-1.5], dtype=tx.dtype), torch.tensor([0.25], dtype=tx.dtype)
    y0 = rng.uniform(-1, 1, n).astype(VT[vt])
    ty2 = _t(y0).clone()
    ald, bed, yd2 = api.host_dense(host, al), api.host_dense(host, be), api.host_dense(host, ty2)
    api._hcheck(api._host().gkob_apply4(B.h, ald.h, xd.h, bed.h, yd2.h))  # handles kept alive
    yo2 = y0.copy()
    orc("csr_advanced_spmv_%s_i32" % vt, n, m, len(va), rp, ci, va, np.array([-1.5], VT[vt]), x, 1, 1,
        np.array([0.25], VT[vt]), yo2, 1)
    if fmt in ("ell", "sellp"):
        assert np.array_equal(ty2.numpy(), yo2)
    else:  # coo / hybrid scale y first, then accumulate: one rounding apart at most
        assert H.rel_err(ty2.numpy(), yo2) <= H.R[vt]


def test_host_sort_and_files(host, orc, tmp_path):
    from ginkgo_b200 import api
    rng = np.random.default_rng(42)
    n, m = 300, 250
    lens = rng.integers(0, 9, n)
    lens[7] = 120
    rp, ci, va = H.random_csr(rng, n, m, lens, "f64", "i32")
    ci2, va2 = ci.copy(), va.copy()
    for r in range(n):
        s, e = rp[r], rp[r + 1]
        perm = rng.permutation(e - s)
        ci2[s:e], va2[s:e] = ci[s:e][perm], va[s:e][perm]
    t = [_t(va2), _t(ci2), _t(rp)]
    A = api.host_csr(host, (n, m), *t)
    api.host_sort_by_column_index(A)
    assert np.array_equal(t[1].numpy(), ci) and np.array_equal(t[0].numpy(), va)
    # write / read round trips of the sorted matrix, every layout that keeps the pattern
    x = rng.uniform(-1, 1, m)
    yo = np.zeros(n)
    orc("csr_spmv_f64_i32", n, m, len(va), rp, ci, va, x, 1, 1, yo, 1)
    for layout in ("coordinate", "binary"):
        out = tmp_path / ("m." + layout)
        api.host_write_csr(A, out, layout)
        B = api.host_read_csr(host, out)
        assert B.size == (n, m)
        ty = torch.zeros(n, dtype=torch.float64)
        xd, yd = api.host_dense(host, _t(x)), api.host_dense(host, ty)
        api._hcheck(api._host().gkob_apply(B.h, xd.h, yd.h))
        assert np.array_equal(ty.numpy(), yo)


def test_host_staged_apply(host, orc):
    from ginkgo_b200 import api
    rng = np.random.default_rng(43)
    n, m = 400, 300
    rp, ci, va = H.random_csr(rng, n, m, rng.integers(0, 12, n), "f64", "i32")
    A = api.host_csr(host, (n, m), _t(va), _t(ci), _t(rp))
    st = api.StagedApply(A)
    xs = [_t(rng.uniform(-1, 1, m)) for _ in range(5)]
    ys = [torch.full((n,), float("nan"), dtype=torch.float64) for _ in range(5)]
    for x, y in zip(xs, ys):
        st.apply(x, y)
    st.wait()
    for x, y in zip(xs, ys):
        yo = np.zeros(n)
        orc("csr_spmv_f64_i32", n, m, len(va), rp, ci, va, x.numpy(), 1, 1, yo, 1)
        assert np.array_equal(y.numpy(), yo)


def test_host_errors(host):
    from ginkgo_b200 import api
    rp, ci, va = W.laplace(6, 2)
    n = len(rp) - 1
    A = api.host_csr(host, (n, n), _t(va), _t(ci), _t(rp))
    bad = api.host_dense(host, torch.zeros(n + 1, dtype=torch.float64))
    y = api.host_dense(host, torch.zeros(n, dtype=torch.float64))
    with pytest.raises(api.DimensionMismatch):
        api._hcheck(api._host().gkob_apply(A.h, bad.h, y.h))
    with pytest.raises(Exception):  # block larger than max_block_size
        api.HostSolver(host, "cg", A, precond_max_bs=2, block_ptrs=np.array([0, 5, n], np.int32),
                       max_iters=5, fused=False)


@pytest.mark.parametrize("kind,extra", [("ir", dict(relaxation_factor=1.0)), ("ir", dict(relaxation_factor=0.3)),
                                        ("chebyshev", dict(foci=(0.3, 7.9))),
                                        ("chebyshev", dict(foci=(0.4, 1.7)))])
@pytest.mark.parametrize("precond", [0, 1, 2])
@pytest.mark.parametrize("vt", ["f64", "f32"])
def test_host_ir_and_chebyshev(host, kind, extra, precond, vt):
    """update_residual.hpp semantics (ignore_residual_check) + the Ir / Chebyshev host loops"""
    from oracle import ref
    rp, ci, va = W.laplace(14, 2, vdtype=VT[vt])
    n = len(rp) - 1
    rng = np.random.default_rng(11)
    b = rng.uniform(-1, 1, (n, 2)).astype(VT[vt])
    x0 = rng.uniform(-1, 1, (n, 2)).astype(VT[vt])
    if kind == "ir" and precond == 0:
        extra = dict(relaxation_factor=extra["relaxation_factor"] * 0.2)
    max_bs = {0: 0, 1: 1, 2: 8}[precond]
    bp = np.arange(0, n + 1, 8, dtype=np.int32)[: n // 8 + 1] if precond == 2 else None
    if precond == 2 and bp[-1] != n:
        bp = np.append(bp, n).astype(np.int32)
    jac = None
    if precond:
        if not ref.available():
            pytest.skip("needs oracle/_ref for the inverted blocks")
        jac = ref.jacobi_generate(rp, ci, va, max_bs, bp)
    for iter_first in (1, 0):
        for res_kind in (1, 0):
            kw = dict(max_iters=40, reduction=1e-3, res_kind=res_kind, krylov_dim=10, **extra)
            xo, ito, stop_o = H.orc_solve(kind, vt, rp, ci, va, b, x0, precond, jac, iter_first=iter_first, **kw)
            xh, ith, stop_h = host_solve(host, kind, vt, rp, ci, va, b, x0, max_bs, bp,
                                         iter_first=bool(iter_first), **kw)
            assert (ith, stop_h) == (ito, stop_o[0])
            assert np.array_equal(xh, xo, equal_nan=True)


@pytest.mark.parametrize("ortho", [0, 1, 2])
@pytest.mark.parametrize("krylov_dim", [5, 30])
def test_host_gmres_ortho_and_restart(host, ortho, krylov_dim):
    """GMRES orthogonalisation variants (mgs / cgs / cgs2) and restarts through the C++ host loop"""
    rp, ci, va = W.laplace(12, 2)
    n = len(rp) - 1
    rng = np.random.default_rng(4)
    b = rng.uniform(-1, 1, (n, 3))
    x0 = rng.uniform(-1, 1, (n, 3))
    kw = dict(max_iters=80, reduction=1e-10, krylov_dim=krylov_dim, ortho=ortho)
    xo, ito, stop_o = H.orc_solve("gmres", "f64", rp, ci, va, b, x0, 0, None, iter_first=1, **kw)
    xh, ith, stop_h = host_solve(host, "gmres", "f64", rp, ci, va, b, x0, iter_first=True, **kw)
    assert (ith, stop_h) == (ito, stop_o[0]) and np.array_equal(xh, xo)


def test_host_advanced_solver_apply(host):
    """x = alpha * solve(b) + beta * x  (core/solver/cg.cpp:184-200: clone, solve, scale, add_scaled)"""
    from ginkgo_b200 import api
    rp, ci, va = W.laplace(10, 2)
    n = len(rp) - 1
    rng = np.random.default_rng(8)
    b = rng.uniform(-1, 1, (n, 1))
    x0 = rng.uniform(-1, 1, (n, 1))
    A = api.host_csr(host, (n, n), _t(va), _t(ci), _t(rp))
    s = api.HostSolver(host, "cg", A, max_iters=200, reduction=1e-10, fused=False)
    tx = _t(x0).clone()
    al, be = torch.tensor([2.0], dtype=torch.float64), torch.tensor([-0.5], dtype=torch.float64)
    ald, bed, bd, xd = (api.host_dense(host, t) for t in (al, be, _t(b), tx))
    api._hcheck(api._host().gkob_apply4(s.obj.h, ald.h, bd.h, bed.h, xd.h))
    xs, _, _ = host_solve(host, "cg", "f64", rp, ci, va, b, x0, max_iters=200, reduction=1e-10)
    expect = x0 * -0.5
    expect = expect + 2.0 * xs
    assert np.array_equal(tx.numpy(), expect)
