"""GPU tests of the last additions of round 1 (IR / Chebyshev / PipeCG / GCR kernels and
solvers, Csr file I/O).  Kept in the file that sorts last: these were written after the round's
GPU budget was spent (their oracle side, the C++ host loops and the test bodies themselves were
verified on the CPU -- tests/test_oracle_vs_ref.py, tests/test_host_cpu.py and the
B200_TEST_SELFCHECK harness), so a surprise here cannot hide any other test."""
import numpy as np
import pytest

import workloads as W
from tests import helpers as H
from tests.helpers import VT
from tests.test_parity_gpu import VTS, _all_equal, both
from tests.test_solvers_gpu import device_solve, hexec, ref_jacobi, true_rel_res  # noqa: F401

pytestmark = H.first_gpu_run_marks()


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols", [(597, 43), (100001, 1), (0, 2)])
def test_ir_and_chebyshev_kernels(orc, cuda, vt, rows, cols):
    rng = np.random.default_rng(92)
    st = dict(inner=cols + 1, update=cols + 2, out=cols)
    v = {k: H.dense(rng, rows, cols, s, vt) for k, s in st.items()}
    a, b = both(orc, cuda, "chebyshev_init_update_" + vt,
                lambda: [rows, cols, 0.37, v["inner"], st["inner"], v["update"].copy(), st["update"],
                         v["out"].copy(), st["out"]])
    _all_equal(a, b)
    a, b = both(orc, cuda, "chebyshev_update_" + vt,
                lambda: [rows, cols, 0.41, 0.0625, v["inner"].copy(), st["inner"], v["update"].copy(),
                         st["update"], v["out"].copy(), st["out"]])
    _all_equal(a, b)
    so, sc = np.full(max(cols, 1), 0xC1, np.uint8), np.full(max(cols, 1), 0xC1, np.uint8)
    orc("ir_initialize", cols, so)
    cuda("ir_initialize", cols, sc)
    assert np.array_equal(so, sc) and not so[:cols].any()


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols", [(597, 43), (100001, 1), (0, 2)])
def test_pipe_cg_steps(orc, cuda, vt, rows, cols):
    rng = np.random.default_rng(93)
    names = ("b", "r", "z1", "z2", "w", "p", "q", "f", "g", "m", "n", "x")
    st = {k: cols + (i % 3) for i, k in enumerate(names)}
    v = {k: H.dense(rng, rows, cols, s, vt) for k, s in st.items()}
    sc = {k: rng.uniform(0.5, 1, cols).astype(VT[vt]) for k in ("rho", "prev_rho", "beta", "delta")}
    stop = np.zeros(cols, dtype=np.uint8)
    if cols > 4:
        sc["prev_rho"][2] = 0
        sc["beta"][3] = 0
        stop[1] = 1 | 0x40
        # a column whose updated beta is exactly zero: beta = delta - |rho/prev_rho|^2 * beta
        sc["rho"][4], sc["prev_rho"][4], sc["beta"][4], sc["delta"][4] = 2.0, 1.0, 0.25, 1.0
    a, b = both(orc, cuda, "pipe_cg_initialize_1_" + vt,
                lambda: [rows, cols, v["b"], st["b"], v["r"].copy(), st["r"], sc["prev_rho"].copy(),
                         np.full(cols, 0x81, np.uint8)])
    _all_equal(a, b)
    a, b = both(orc, cuda, "pipe_cg_initialize_2_" + vt,
                lambda: [rows, cols] + sum([[v[k].copy(), st[k]] for k in ("p", "q", "f", "g")], []) +
                [sc["beta"].copy()] + sum([[v[k], st[k]] for k in ("z1", "w", "m", "n")], []) + [sc["delta"]])
    _all_equal(a, b)
    a, b = both(orc, cuda, "pipe_cg_step_1_" + vt,
                lambda: [rows, cols] + sum([[v[k].copy(), st[k]] for k in ("x", "r", "z1", "z2", "w")], []) +
                sum([[v[k], st[k]] for k in ("p", "q", "f", "g")], []) + [sc["rho"], sc["beta"], stop])
    _all_equal(a, b)
    a, b = both(orc, cuda, "pipe_cg_step_2_" + vt,
                lambda: [rows, cols, sc["beta"].copy()] + sum([[v[k].copy(), st[k]] for k in ("p", "q", "f", "g")], []) +
                sum([[v[k], st[k]] for k in ("z1", "w", "m", "n")], []) +
                [sc["prev_rho"], sc["rho"], sc["delta"], stop])
    _all_equal(a, b)


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols", [(597, 43), (100001, 1), (0, 2)])
def test_gcr_kernels(orc, cuda, vt, rows, cols):
    rng = np.random.default_rng(94)
    names = ("b", "res", "ares", "p", "ap", "x")
    st = {k: cols + (i % 3) for i, k in enumerate(names)}
    v = {k: H.dense(rng, rows, cols, s, vt) for k, s in st.items()}
    ap_norm = rng.uniform(0.5, 1, cols).astype(VT[vt])
    rap = rng.uniform(-1, 1, cols).astype(VT[vt])
    stop = np.zeros(cols, dtype=np.uint8)
    if cols > 3:
        ap_norm[2] = 0
        stop[1] = 1 | 0x40
    a, b = both(orc, cuda, "gcr_initialize_" + vt,
                lambda: [rows, cols, v["b"], st["b"], v["res"].copy(), st["res"], np.full(cols, 0x81, np.uint8)])
    _all_equal(a, b)
    a, b = both(orc, cuda, "gcr_restart_" + vt,
                lambda: [rows, cols, v["res"], st["res"], v["ares"], st["ares"], v["p"].copy(), st["p"],
                         v["ap"].copy(), st["ap"], np.full(cols, 7, np.uint64)])
    _all_equal(a, b)
    a, b = both(orc, cuda, "gcr_step_1_" + vt,
                lambda: [rows, cols, v["x"].copy(), st["x"], v["res"].copy(), st["res"], v["p"], st["p"],
                         v["ap"], st["ap"], ap_norm, rap, stop])
    _all_equal(a, b)


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("rows,cols", [(597, 43), (100001, 1), (0, 2)])
def test_minres_kernels(orc, cuda, vt, rows, cols):
    rng = np.random.default_rng(95)
    names = ("r", "z", "p", "p_prev", "q", "q_prev", "v", "z_tilde", "x")
    st = {k: cols + (i % 3) for i, k in enumerate(names)}
    v = {k: H.dense(rng, rows, cols, s, vt) for k, s in st.items()}
    sc = {k: rng.uniform(0.3, 1, cols).astype(VT[vt])
          for k in ("alpha", "beta", "gamma", "delta", "cos_prev", "cos", "sin_prev", "sin", "eta_next",
                    "eta", "tau")}
    stop = np.zeros(cols, dtype=np.uint8)
    if cols > 4:
        stop[1] = 1 | 0x40
        sc["beta"][2] = 0    # safe_divide by zero
        sc["alpha"][3] = 0
    a, b = both(orc, cuda, "minres_initialize_" + vt,
                lambda: [rows, cols, v["r"], st["r"]] +
                sum([[v[k].copy(), st[k]] for k in ("z", "p", "p_prev", "q", "q_prev", "v")], []) +
                [sc[k].copy() for k in ("beta", "gamma", "delta", "cos_prev", "cos", "sin_prev", "sin",
                                        "eta_next", "eta")] + [np.full(cols, 0x81, np.uint8)])
    _all_equal(a, b)
    a, b = both(orc, cuda, "minres_step_1_" + vt,
                lambda: [cols] + [sc[k].copy() for k in ("alpha", "beta", "gamma", "delta", "cos_prev",
                                                        "cos", "sin_prev", "sin", "eta", "eta_next",
                                                        "tau")] + [stop])
    _all_equal(a, b)
    # a column whose rotated alpha is exactly zero: gamma = 0, alpha = 0 -> cos = 0, sin = 1
    if cols > 4:
        z = {k: sc[k].copy() for k in sc}
        z["alpha"][4] = 0
        z["gamma"][4] = 0
        a, b = both(orc, cuda, "minres_step_1_" + vt,
                    lambda: [cols] + [z[k].copy() for k in ("alpha", "beta", "gamma", "delta", "cos_prev",
                                                           "cos", "sin_prev", "sin", "eta", "eta_next",
                                                           "tau")] + [stop])
        _all_equal(a, b)
    a, b = both(orc, cuda, "minres_step_2_" + vt,
                lambda: [rows, cols, v["x"].copy(), st["x"], v["p"].copy(), st["p"], v["p_prev"], st["p_prev"],
                         v["z"].copy(), st["z"], v["z_tilde"], st["z_tilde"], v["q"].copy(), st["q"],
                         v["q_prev"].copy(), st["q_prev"], v["v"].copy(), st["v"], sc["alpha"], sc["beta"],
                         sc["gamma"], sc["delta"], sc["cos"], sc["eta"], stop])
    _all_equal(a, b)


@pytest.mark.parametrize("kind,extra", [("ir", dict(relaxation_factor=0.9)), ("chebyshev", dict(foci=(0.4, 1.7))),
                                        ("pipe_cg", {}), ("gcr", dict(krylov_dim=20)), ("minres", {})])
@pytest.mark.parametrize("vt", ["f64", "f32"])
def test_zy_ir_and_chebyshev_match_oracle(hexec, kind, extra, vt):
    """Jacobi-preconditioned Richardson / Chebyshev iteration: no inner products, so the device
    run follows the oracle (= reference) loop to rounding of the residual norms only"""
    rp, ci, va = W.laplace(30, 2, vdtype=VT[vt])
    n = len(rp) - 1
    rng = np.random.default_rng(6)
    b = rng.uniform(-1, 1, (n, 1)).astype(VT[vt])
    x0 = np.zeros((n, 1), VT[vt])
    jac = ref_jacobi(vt, rp, ci, va, 1, None)
    red = 1e-4 if vt == "f64" else 1e-3
    xo, ito, stop_o = H.orc_solve(kind, vt, rp, ci, va, b, x0, 1, jac, max_iters=3000, reduction=red,
                                  iter_first=1, **extra)
    xd, itd, stop_d, _ = device_solve(hexec, kind, vt, rp, ci, va, b, x0, 1, None, max_iters=3000,
                                      reduction=red, iter_first=True, fused=False, **extra)
    if kind in ("pipe_cg", "gcr", "minres"):  # dot products: tree vs sequential order (PipeCG amplifies it)
        assert abs(itd - ito) <= max(3, 0.2 * ito) and stop_d == stop_o[0]
        rd = true_rel_res(rp, ci, va, b, xd)
        assert rd[0] <= 20 * red, rd
        return
    assert abs(itd - ito) <= 2 and stop_d == stop_o[0]
    if itd == ito:  # same number of steps: the iterates agree to rounding
        assert H.rel_err(xo, xd) <= (1e-10 if vt == "f64" else 1e-4)
    else:  # the residual norm crossed the threshold one step apart
        assert true_rel_res(rp, ci, va, b, xd)[0] <= 2 * red


def test_zz_read_write_csr_files(hexec, orc, tmp_path):
    """gko::read_generic<Csr> / gko::write on the device executor: a file written by the host
    layer is read back into a Csr whose apply matches the oracle (kept last in the suite)."""
    import torch
    from ginkgo_b200 import api
    rng = np.random.default_rng(44)
    n, m = 300, 250
    rp, ci, va = H.random_csr(rng, n, m, rng.integers(0, 9, n), "f64", "i32")
    rows = np.repeat(np.arange(n), np.diff(rp))
    path = tmp_path / "a.mtx"
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")  # no %-formatting here
        f.write("%d %d %d\n" % (n, m, len(va)))
        for r, c, v in zip(rows, ci, va):
            f.write("%d %d %.17g\n" % (r + 1, c + 1, v))
    A = api.host_read_csr(hexec, path)
    assert A.size == (n, m)
    x = rng.uniform(-1, 1, m)
    with torch.cuda.stream(hexec.stream):
        tx = torch.from_numpy(x).to(hexec.device)
        ty = torch.zeros(n, dtype=torch.float64, device=hexec.device)
    xd, yd = api.host_dense(hexec, tx), api.host_dense(hexec, ty)
    api._hcheck(api._host().gkob_apply(A.h, xd.h, yd.h))
    hexec.synchronize()
    yo = np.zeros(n)
    orc("csr_spmv_f64_i32", n, m, len(va), rp, ci, va, x, 1, 1, yo, 1)
    assert np.array_equal(ty.cpu().numpy(), yo)
    for layout in ("coordinate", "binary"):
        out = tmp_path / ("b." + layout)
        api.host_write_csr(A, out, layout)
        B = api.host_read_csr(hexec, out)
        with torch.cuda.stream(hexec.stream):
            ty.zero_()
        api._hcheck(api._host().gkob_apply(B.h, xd.h, yd.h))
        hexec.synchronize()
        assert np.array_equal(ty.cpu().numpy(), yo)
