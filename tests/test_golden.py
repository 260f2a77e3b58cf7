"""Known-answer tests transcribed from the reference's own reference-executor
suites (the literal vectors, not the code): every case runs on the oracle
(CPU, pins the oracle) and -- marked gpu -- on the CUDA library through the
C ABI.  Citations are relative to /root/reference."""
import numpy as np
import pytest

from tests.helpers import IT, R, VT, OutI64, OutInt, rel_err

BACKENDS = ["oracle", pytest.param("cuda", marks=pytest.mark.gpu)]
VTS = ["f64", "f32"]
ITS = ["i32", "i64"]


@pytest.fixture(params=BACKENDS)
def be(request):
    return request.getfixturevalue("orc" if request.param == "oracle" else "cuda")


def arr(x, vt):
    return np.array(x, dtype=VT[vt])


def csr_fix(vt, it):
    # reference/test/matrix/csr_kernels.cpp:84-106:  [[1,3,2],[0,5,0]]
    return (np.array([0, 3, 4], dtype=IT[it]), np.array([0, 1, 2, 1], dtype=IT[it]),
            arr([1, 3, 2, 5], vt))


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
class TestSpmvFormats:
    def test_csr_applies_to_dense_vector(self, be, vt, it):
        # reference/test/matrix/csr_kernels.cpp:353-364
        rp, ci, va = csr_fix(vt, it)
        x, y = arr([2, 1, 4], vt), arr([0, 0], vt)
        be("csr_spmv_%s_%s" % (vt, it), 2, 3, 4, rp, ci, va, x, 1, 1, y, 1)
        assert y.tolist() == [13.0, 5.0]

    def test_csr_applies_to_dense_matrix(self, be, vt, it):
        # reference/test/matrix/csr_kernels.cpp:414-430
        rp, ci, va = csr_fix(vt, it)
        x = arr([[2, 3], [1, -1.5], [4, 2.5]], vt)
        y = np.zeros((2, 2), dtype=VT[vt])
        be("csr_spmv_%s_%s" % (vt, it), 2, 3, 4, rp, ci, va, x, 2, 2, y, 2)
        assert y.tolist() == [[13.0, 3.5], [5.0, -7.5]]

    def test_csr_linear_combination(self, be, vt, it):
        # reference/test/matrix/csr_kernels.cpp:505-518
        rp, ci, va = csr_fix(vt, it)
        x, y = arr([2, 1, 4], vt), arr([1, 2], vt)
        be("csr_advanced_spmv_%s_%s" % (vt, it), 2, 3, 4, rp, ci, va, arr([-1], vt), x, 1, 1,
           arr([2], vt), y, 1)
        assert y.tolist() == [-11.0, -1.0]

    def test_csr_zero_beta_overwrites_nan(self, be, vt, it):
        # reference/test/matrix/csr_kernels.cpp:521-534
        rp, ci, va = csr_fix(vt, it)
        x, y = arr([2, 1, 4], vt), arr([np.nan, np.nan], vt)
        be("csr_advanced_spmv_%s_%s" % (vt, it), 2, 3, 4, rp, ci, va, arr([-1], vt), x, 1, 1,
           arr([0], vt), y, 1)
        assert y.tolist() == [-13.0, -5.0]

    def test_csr_linear_combination_dense_matrix(self, be, vt, it):
        # reference/test/matrix/csr_kernels.cpp:590-608
        rp, ci, va = csr_fix(vt, it)
        x = arr([[2, 3], [1, -1.5], [4, 2.5]], vt)
        y = arr([[1, 0.5], [2, -1.5]], vt)
        be("csr_advanced_spmv_%s_%s" % (vt, it), 2, 3, 4, rp, ci, va, arr([-1], vt), x, 2, 2,
           arr([2], vt), y, 2)
        assert y.tolist() == [[-11.0, -2.5], [-1.0, 4.5]]

    def _ell(self, vt, it):
        # same 2x3 matrix in ELL (reference/test/matrix/ell_kernels.cpp fixture):
        # 3 stored entries per row, stride 2, padding col -1
        cols = np.array([0, 1, 1, -1, 2, -1], dtype=IT[it])
        vals = arr([1, 5, 3, 0, 2, 0], vt)
        return cols, vals

    def test_ell_apply(self, be, vt, it):
        cols, vals = self._ell(vt, it)
        x, y = arr([2, 1, 4], vt), arr([0, 0], vt)
        be("ell_spmv_%s_%s" % (vt, it), 2, 3, 3, 2, cols, vals, x, 1, 1, y, 1)
        assert y.tolist() == [13.0, 5.0]
        y = arr([1, 2], vt)
        be("ell_advanced_spmv_%s_%s" % (vt, it), 2, 3, 3, 2, cols, vals, arr([-1], vt), x, 1, 1,
           arr([2], vt), y, 1)
        assert y.tolist() == [-11.0, -1.0]

    def test_sellp_apply(self, be, vt, it):
        # one slice of size 64 holding the 2 rows, slice length 3
        ss = 64
        cols = np.full(3 * ss, -1, dtype=IT[it])
        vals = np.zeros(3 * ss, dtype=VT[vt])
        for i, (c0, v0) in enumerate([(0, 1), (1, 3), (2, 2)]):
            cols[i * ss + 0], vals[i * ss + 0] = c0, v0
        cols[0 * ss + 1], vals[0 * ss + 1] = 1, 5
        sets = np.array([0, 3], dtype=np.uint64)
        lens = np.array([3], dtype=np.uint64)
        x, y = arr([2, 1, 4], vt), arr([0, 0], vt)
        be("sellp_spmv_%s_%s" % (vt, it), 2, 3, ss, sets, lens, cols, vals, x, 1, 1, y, 1)
        assert y.tolist() == [13.0, 5.0]
        y = arr([1, 2], vt)
        be("sellp_advanced_spmv_%s_%s" % (vt, it), 2, 3, ss, sets, lens, cols, vals,
           arr([-1], vt), x, 1, 1, arr([2], vt), y, 1)
        assert y.tolist() == [-11.0, -1.0]

    def test_coo_apply(self, be, vt, it):
        # reference/test/matrix/coo_kernels.cpp: same matrix as sorted triples
        rows = np.array([0, 0, 0, 1], dtype=IT[it])
        _, ci, va = csr_fix(vt, it)
        x, y = arr([2, 1, 4], vt), arr([0, 0], vt)
        be("coo_spmv_%s_%s" % (vt, it), 2, 3, 4, rows, ci, va, x, 1, 1, y, 1)
        assert y.tolist() == [13.0, 5.0]
        y = arr([1, 2], vt)
        be("coo_advanced_spmv_%s_%s" % (vt, it), 2, 3, 4, rows, ci, va, arr([-1], vt), x, 1, 1,
           arr([2], vt), y, 1)
        assert y.tolist() == [-11.0, -1.0]
        y = arr([2, 1], vt)  # apply2: y += A x
        be("coo_spmv2_%s_%s" % (vt, it), 2, 3, 4, rows, ci, va, x, 1, 1, y, 1)
        assert y.tolist() == [15.0, 6.0]
        y = arr([2, 1], vt)  # y += alpha A x
        be("coo_advanced_spmv2_%s_%s" % (vt, it), 2, 3, 4, rows, ci, va, arr([-1], vt), x, 1, 1,
           y, 1)
        assert y.tolist() == [-11.0, -4.0]


STOPPED = 1 | 0x40  # stopping_status.stop(1): id 1 + finalized


@pytest.mark.parametrize("vt", VTS)
class TestKrylovKernels:
    def test_cg_initialize(self, be, vt):
        # reference/test/solver/cg_kernels.cpp:113-138 (b stride 3: cols + 1)
        b = np.full((2, 3), 2, dtype=VT[vt])
        r = np.zeros((2, 2), VT[vt])
        z, p, q = (np.ones((2, 2), VT[vt]) for _ in range(3))
        prev_rho, rho = arr([0, 0], vt), arr([1, 1], vt)
        stop = np.array([STOPPED, STOPPED], dtype=np.uint8)
        be("cg_initialize_" + vt, 2, 2, b, 3, r, 2, z, 2, p, 2, q, 2, prev_rho, rho, stop)
        assert (r == 2).all() and (z == 0).all() and (p == 0).all() and (q == 0).all()
        assert rho.tolist() == [0, 0] and prev_rho.tolist() == [1, 1]
        assert stop.tolist() == [0, 0]

    def test_cg_step_1(self, be, vt):
        # reference/test/solver/cg_kernels.cpp:141-156
        p, z = np.full((2, 2), 3, VT[vt]), np.full((2, 2), -2, VT[vt])
        stop = np.array([0, STOPPED], dtype=np.uint8)
        be("cg_step_1_" + vt, 2, 2, p, 2, z, 2, arr([2, 3], vt), arr([8, 3], vt), stop)
        assert p.tolist() == [[-1.25, 3.0], [-1.25, 3.0]]

    def test_cg_step_1_div_by_zero(self, be, vt):
        # reference/test/solver/cg_kernels.cpp:159-171
        p, z = np.full((2, 2), 3, VT[vt]), np.full((2, 2), -2, VT[vt])
        stop = np.zeros(2, dtype=np.uint8)
        be("cg_step_1_" + vt, 2, 2, p, 2, z, 2, arr([1, 1], vt), arr([0, 0], vt), stop)
        assert (p == -2).all()

    def test_cg_step_2(self, be, vt):
        # reference/test/solver/cg_kernels.cpp:174-194 (x stride 4: cols + 2)
        x = np.full((2, 4), -2, VT[vt])
        r, p, q = (np.full((2, 2), v, VT[vt]) for v in (4, 3, -5))
        stop = np.array([0, STOPPED], dtype=np.uint8)
        be("cg_step_2_" + vt, 2, 2, x, 4, r, 2, p, 2, q, 2, arr([8, 3], vt), arr([2, 3], vt), stop)
        assert x[:, :2].tolist() == [[-1.25, -2.0], [-1.25, -2.0]]
        assert r.tolist() == [[5.25, 4.0], [5.25, 4.0]]

    def test_cg_step_2_div_by_zero(self, be, vt):
        # reference/test/solver/cg_kernels.cpp:197-212
        x = np.full((2, 2), -2, VT[vt])
        r, p, q = (np.full((2, 2), v, VT[vt]) for v in (4, 3, -5))
        stop = np.zeros(2, dtype=np.uint8)
        be("cg_step_2_" + vt, 2, 2, x, 2, r, 2, p, 2, q, 2, arr([0, 0], vt), arr([1, 1], vt), stop)
        assert (x == -2).all() and (r == 4).all()

    def test_gmres_hessenberg_qr_iter0(self, be, vt):
        # reference/test/solver/gmres_kernels.cpp:213-263
        nan = np.nan
        gcos, gsin = arr([[-0.5, 1.], [70., -71]], vt), arr([[1., 0.], [-72., 73.]], vt)
        rn = arr([nan, nan], vt)
        rnc = arr([[1.25, 1.5], [nan, nan], [95., 94.]], vt)
        hess = arr([0.5, -0.75, -0.5, 1, 97., 96.], vt)
        fin = np.array([0, 0], dtype=np.uint64)
        stop = np.zeros(2, dtype=np.uint8)
        be("common_gmres_hessenberg_qr_" + vt, 2, gsin, 2, gcos, 2, rn, rnc, 2, hess, 2, 0, fin, stop)
        s2 = np.sqrt(2.)
        assert fin.tolist() == [1, 1]
        assert rel_err(gcos, [[0.5 * s2, -0.6], [70., -71.]]) <= R[vt]
        assert rel_err(gsin, [[-0.5 * s2, 0.8], [-72., 73.]]) <= R[vt]
        assert rel_err(hess, [0.5 * s2, 1.25, 0., 0., 97., 96.]) <= R[vt]
        assert rel_err(rnc, [[0.625 * s2, -0.9], [0.625 * s2, -1.2], [95., 94.]]) <= R[vt]
        assert rel_err(rn, [0.625 * s2, 1.2]) <= R[vt]

    def test_gmres_hessenberg_qr_iter1(self, be, vt):
        # reference/test/solver/gmres_kernels.cpp:266-316
        nan = np.nan
        gcos, gsin = arr([[1., 0.5], [-0.5, 1.]], vt), arr([[0.5, 0.25], [1., 0.]], vt)
        rn = arr([nan, nan], vt)
        rnc = arr([[95., 94.], [1.25, 1.5], [nan, nan]], vt)
        hess = arr([-0.5, 4, 0.25, 0.5, -0.5, 1], vt)
        fin = np.array([1, 1], dtype=np.uint64)
        stop = np.zeros(2, dtype=np.uint8)
        be("common_gmres_hessenberg_qr_" + vt, 2, gsin, 2, gcos, 2, rn, rnc, 2, hess, 2, 1, fin, stop)
        s2 = np.sqrt(2.)
        assert fin.tolist() == [2, 2]
        assert rel_err(gcos, [[1., 0.5], [0.5 * s2, -0.6]]) <= R[vt]
        assert rel_err(gsin, [[0.5, 0.25], [-0.5 * s2, 0.8]]) <= R[vt]
        assert rel_err(hess, [-0.375, 2.125, 0.5 * s2, 1.25, 0., 0.]) <= R[vt]
        assert rel_err(rnc, [[95., 94.], [0.625 * s2, -0.9], [0.625 * s2, -1.2]]) <= R[vt]
        assert rel_err(rn, [0.625 * s2, 1.2]) <= R[vt]

    def test_gmres_solve_krylov(self, be, vt):
        # reference/test/solver/gmres_kernels.cpp:319-340: hessenberg 2 x 6 (stride 6)
        nan = np.nan
        hess = arr([[-1, 3, 0, 0, nan, nan], [2, -4, 1, 5, nan, nan]], vt)
        rnc = arr([[12, 3], [-3, 15]], vt)
        y = arr([[nan, nan], [nan, nan]], vt)
        fin = np.array([2, 2], dtype=np.uint64)
        stop = np.zeros(2, dtype=np.uint8)
        be("common_gmres_solve_krylov_" + vt, 2, rnc, 2, hess, 6, y, 2, fin, stop)
        assert rel_err(y, [[-18., 5.], [-3., 3.]]) <= R[vt]

    def test_gmres_multi_axpy(self, be, vt):
        # reference/test/solver/gmres_kernels.cpp:343-382
        nan = np.nan
        y = arr([[1., 2.], [3., -1.]], vt)
        kb = arr([[1, 10], [2, 11], [3, 12], [4, 13], [5, 14], [6, 15], [nan, nan], [nan, nan],
                  [nan, nan]], vt)
        x = arr([[nan, nan]] * 3, vt)
        fin = np.array([2, 2], dtype=np.uint64)
        stop = np.array([7, 0], dtype=np.uint8)  # stop(7, false)
        be("gmres_multi_axpy_" + vt, 3, 2, kb, 2, y, 2, x, 2, fin, stop)
        assert stop.tolist() == [7 | 0x40, 0]
        assert rel_err(x, [[13., 7.], [17., 8.], [21., 9.]]) <= R[vt]

    def test_gmres_multi_dot(self, be, vt):
        # reference/test/solver/gmres_kernels.cpp:384-420 (3 rows in the column, 2 bases dotted)
        nk = arr([[-1.0, 2.3], [-14.0, -22.0], [8.4, 14.2]], vt)
        kb = arr([[1, 10], [2, 11], [3, 12], [4, 13], [5, 14], [6, 15], [7, 16], [8, 17],
                  [9, 18]], vt)
        h = np.zeros((3, 2), VT[vt])
        be("gmres_multi_dot_" + vt, 3, 2, 2, kb, 2, nk, 2, h, 2)
        assert rel_err(h, [[-3.8, -48.6], [-23.6, -65.1], [0.0, 0.0]]) <= R[vt]

    def test_gmres_restart(self, be, vt):
        # reference/test/solver/gmres_kernels.cpp:170-210
        b = arr([[1, 2], [3, 4], [-5, 6]], vt)
        rn = np.sqrt((b.astype(np.float64) ** 2).sum(0)).astype(VT[vt])
        rnc = np.full((3, 2), np.nan, VT[vt])
        kb = np.full((9, 2), 9999, VT[vt])
        fin = np.array([999, 999], dtype=np.uint64)
        be("gmres_restart_" + vt, 3, 2, b, 2, rn, rnc, kb, 2, fin)
        assert fin.tolist() == [0, 0]
        assert rnc[0].tolist() == rn.tolist()
        assert rel_err(kb[:3], b / rn) <= R[vt]
        assert (kb[3:] == 9999).all()

    def test_residual_norm(self, be, vt):
        # reference/test/stop/residual_norm_kernels.cpp (WaitsTillResidualGoal pattern)
        goal = 1e-3 if vt == "f32" else 1e-9
        orig = arr([100.0, 100.0], vt)
        stop = np.zeros(2, dtype=np.uint8)
        store = np.zeros(2, dtype=np.uint8)
        ac, oc = OutInt(), OutInt()
        tau = arr([100.0 * goal * 10, 100.0 * goal * 0.5], vt)
        be("residual_norm_" + vt, 2, tau, orig, VT[vt](goal).item(), 1, 1, stop, store, ac, oc)
        assert (ac.value, oc.value) == (0, 1)
        assert stop.tolist() == [0, 0x80 | 0x40 | 1]
        tau = arr([100.0 * goal * 0.5, 100.0 * goal * 0.5], vt)
        be("residual_norm_" + vt, 2, tau, orig, VT[vt](goal).item(), 2, 0, stop, store, ac, oc)
        assert (ac.value, oc.value) == (1, 1)
        assert stop.tolist() == [0x80 | 2, 0x80 | 0x40 | 1]

    def test_implicit_residual_norm(self, be, vt):
        goal = 1e-3
        orig = arr([4.0], vt)
        stop = np.zeros(1, dtype=np.uint8)
        store = np.zeros(2, dtype=np.uint8)
        ac, oc = OutInt(), OutInt()
        be("implicit_residual_norm_" + vt, 1, arr([-1.0], vt), orig, goal, 1, 1, stop, store, ac, oc)
        assert (ac.value, oc.value, stop.tolist()) == (0, 0, [0])
        be("implicit_residual_norm_" + vt, 1, arr([-1e-6 * 15.9], vt), orig, goal, 1, 1, stop, store,
           ac, oc)  # sqrt(|tau|) = 3.987e-3 <= 4e-3
        assert (ac.value, oc.value, stop.tolist()) == (1, 1, [0x80 | 0x40 | 1])


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
class TestJacobi:
    def _blocks(self, vt):
        # reference/test/preconditioner/jacobi_kernels.cpp:244-268: inverted diagonal blocks of the
        # 5x5 fixture (:35-72), max_block_size 3 -> block_offset 3, group_size 32/4 = 8,
        # stride 24, group_offset 72 (include/ginkgo/core/preconditioner/jacobi.hpp:589-625)
        bo, gp = 3, 3
        stride = bo << gp
        go = 3 * stride
        blocks = np.zeros(go, dtype=np.float64)
        b1 = np.array([[4, 2], [1, 4]]) / 14.0
        b2 = np.array([[14, 8, 4], [4, 16, 8], [1, 4, 14]]) / 48.0
        for r in range(2):
            for c in range(2):
                blocks[0 * bo + r + c * stride] = b1[r, c]
        for r in range(3):
            for c in range(3):
                blocks[1 * bo + r + c * stride] = b2[r, c]
        return bo, go, gp, blocks.astype(VT[vt])

    def test_block_apply(self, be, vt, it):
        # reference/test/preconditioner/jacobi_kernels.cpp:627-640
        bo, go, gp, blocks = self._blocks(vt)
        ptrs = np.array([0, 2, 5], dtype=IT[it])
        x, b = arr([1, -1, 2, -2, 3], vt), arr([4, -1, -2, 4, -1], vt)
        be("jacobi_simple_apply_%s_%s" % (vt, it), 2, 3, bo, go, gp, ptrs, blocks, b, 1, 1, x, 1)
        assert rel_err(x, [1, 0, 0, 1, 0]) <= R[vt]

    def test_block_apply_linear_combination(self, be, vt, it):
        # reference/test/preconditioner/jacobi_kernels.cpp:868-885
        bo, go, gp, blocks = self._blocks(vt)
        ptrs = np.array([0, 2, 5], dtype=IT[it])
        x, b = arr([1, -1, 2, -2, 3], vt), arr([4, -1, -2, 4, -1], vt)
        be("jacobi_apply_%s_%s" % (vt, it), 2, 3, bo, go, gp, ptrs, blocks, arr([2], vt), b, 1, 1,
           arr([-1], vt), x, 1)
        assert rel_err(x, [1, 1, -2, 4, -3]) <= R[vt]

    def test_scalar_jacobi(self, be, vt, it):
        # diag of the fixture is all 4: reference/test/preconditioner/jacobi_kernels.cpp:643-700
        diag = arr([4, 4, 4, 0, 4], vt)
        inv = np.zeros(5, VT[vt])
        be("jacobi_invert_diagonal_" + vt, 5, diag, inv)
        assert inv.tolist() == [0.25, 0.25, 0.25, 1.0, 0.25]  # zero diagonal -> 1
        inv[3] = 0.25
        x, b = arr([1, -1, 2, -2, 3], vt), arr([4, -1, -2, 4, -1], vt)
        be("jacobi_simple_scalar_apply_" + vt, 5, 1, inv, b, 1, x, 1)
        assert x.tolist() == [1.0, -0.25, -0.5, 1.0, -0.25]
        x = arr([1, -1, 2, -2, 3], vt)
        be("jacobi_scalar_apply_" + vt, 5, 1, inv, arr([2], vt), b, 1, arr([-1], vt), x, 1)
        assert x.tolist() == [1.0, 0.5, -3.0, 4.0, -3.5]


    # ---- set-up kernels -------------------------------------------------------
    def _fixture(self, vt, it):
        # reference/test/preconditioner/jacobi_kernels.cpp:58-71: the 5x5 test matrix
        rp = np.array([0, 3, 5, 7, 10, 13], dtype=IT[it])
        ci = np.array([0, 1, 4, 0, 1, 2, 3, 2, 3, 4, 0, 3, 4], dtype=IT[it])
        va = arr([4, -2, -2, -1, 4, 4, -2, -1, 4, -2, -1, -1, 4], vt)
        return rp, ci, va

    def test_generate_inverts_diagonal_blocks(self, be, vt, it):
        # reference/test/preconditioner/jacobi_kernels.cpp:244-268 (InvertsDiagonalBlocks)
        bo, go, gp, expect = self._blocks(vt)
        rp, ci, va = self._fixture(vt, it)
        ptrs = np.array([0, 2, 5], dtype=IT[it])
        blocks = np.zeros_like(expect)
        be("jacobi_generate_%s_%s" % (vt, it), 5, rp, ci, va, 2, 3, bo, go, gp, ptrs, blocks)
        assert rel_err(blocks, expect) <= R[vt]

    def test_generate_pivots(self, be, vt, it):
        # reference/test/preconditioner/jacobi_kernels.cpp:452-490 (PivotsWhenInvertingBlocks)
        rp = np.array([0, 3, 6, 9], dtype=IT[it])
        ci = np.array([0, 1, 2, 0, 1, 2, 0, 1, 2], dtype=IT[it])
        va = arr([0, 2, 0, 0, 0, 4, 1, 0, 0], vt)
        bo, gp = 3, 3
        stride = bo << gp
        blocks = np.zeros(3 * stride, VT[vt])
        be("jacobi_generate_%s_%s" % (vt, it), 3, rp, ci, va, 1, 3, bo, 3 * stride, gp,
           np.array([0, 3], dtype=IT[it]), blocks)
        inv = np.array([[blocks[r + c * stride] for c in range(3)] for r in range(3)])
        assert rel_err(inv, np.array([[0, 0, 4], [2, 0, 0], [0, 1, 0]]) / 4.0) <= R[vt]

    @pytest.mark.parametrize("case", ["natural", "agglomeration", "size_bound", "fixture"])
    def test_find_blocks(self, be, vt, it, case):
        # reference/test/preconditioner/jacobi_kernels.cpp:129-241 (FindsNaturalBlocks,
        # ExecutesSupervariableAgglomeration, AdheresToBlockSizeBound,
        # CanBeGeneratedWithUnknownBlockSizes), all with max_block_size 3
        if vt == "f32":
            pytest.skip("index-only kernel")
        if case == "natural":
            rp, ci, expect = [0, 2, 4, 6, 8], [0, 1, 0, 1, 0, 2, 0, 2], [0, 2, 4]
        elif case == "agglomeration":
            rp, ci, expect = [0, 2, 4, 6, 8, 9], [0, 1, 0, 1, 2, 3, 2, 3, 4], [0, 2, 5]
        elif case == "size_bound":
            rp, ci, expect = list(range(8)), list(range(7)), [0, 3, 6, 7]
        else:
            rp, ci, _ = self._fixture(vt, it)
            rp, ci, expect = rp.tolist(), ci.tolist(), [0, 3, 5]
        n = len(rp) - 1
        ptrs = np.full(n + 1, -1, dtype=IT[it])
        nb = OutI64()
        be("jacobi_find_blocks_" + it, n, np.array(rp, dtype=IT[it]), np.array(ci, dtype=IT[it]), 3,
           ptrs, nb)
        assert nb.value == len(expect) - 1
        assert ptrs[:len(expect)].tolist() == expect


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
class TestConversions:
    """reference/test/matrix/csr_kernels.cpp: ConvertsToEll (:1585), ConvertsToSellp (:1251),
    ConvertsToHybridAutomatically (:1301), ConvertsToHybridByColumn2 (:1326), SortUnsortedMatrix
    (:2335) with the expectations of assert_equal_to_mtx (:190-332)."""

    def test_converts_to_ell(self, be, vt, it):
        rp, ci, va = csr_fix(vt, it)
        mx = OutI64()
        be("ell_compute_max_row_nnz_" + it, rp, 2, mx)
        assert mx.value == 3
        cols, vals = np.full(6, 9, IT[it]), np.full(6, 9, VT[vt])
        be("csr_convert_to_ell_%s_%s" % (vt, it), 2, rp, ci, va, 3, 2, cols, vals)
        assert cols.tolist() == [0, 1, 1, -1, 2, -1]
        assert vals.tolist() == [1.0, 5.0, 3.0, 0.0, 2.0, 0.0]

    def test_converts_to_sellp(self, be, vt, it):
        rp, ci, va = csr_fix(vt, it)
        ss, sl = np.full(2, 9, np.uint64), np.full(1, 9, np.uint64)
        be("sellp_compute_slice_sets_" + it, rp, 2, 64, 1, ss, sl)
        assert ss.tolist() == [0, 3] and sl.tolist() == [3]
        cols, vals = np.full(192, 9, IT[it]), np.full(192, 9, VT[vt])
        be("csr_convert_to_sellp_%s_%s" % (vt, it), 2, 64, ss, sl, rp, ci, va, cols, vals)
        assert [cols[i] for i in (0, 1, 64, 65, 128, 129)] == [0, 1, 1, -1, 2, -1]
        assert [vals[i] for i in (0, 1, 64, 65, 128, 129)] == [1.0, 5.0, 3.0, 0.0, 2.0, 0.0]

    def test_converts_to_hybrid_automatically(self, be, vt, it):
        # automatic = imbalance_bounded_limit(1/3, 0.001): min(sorted_nnz[0], 2 * 0.001) = 0
        rp, ci, va = csr_fix(vt, it)
        k = OutI64()
        be("csr_row_nnz_order_statistic_" + it, rp, 2, int(2 * (1.0 / 3.0)), k)
        ell_lim = min(k.value, int(2 * 0.001))
        assert ell_lim == 0
        crp = np.full(3, -1, np.int64)
        be("csr_compute_hybrid_coo_row_ptrs_" + it, rp, 2, ell_lim, crp)
        assert crp.tolist() == [0, 3, 4]
        ec, ev = np.zeros(1, IT[it]), np.zeros(1, VT[vt])
        cr, cc, cv = np.full(4, 9, IT[it]), np.full(4, 9, IT[it]), np.full(4, 9, VT[vt])
        be("csr_convert_to_hybrid_%s_%s" % (vt, it), 2, rp, ci, va, 0, 2, ec, ev, crp, cr, cc, cv)
        assert (cr.tolist(), cc.tolist(), cv.tolist()) == ([0, 0, 0, 1], [0, 1, 2, 1],
                                                           [1.0, 3.0, 2.0, 5.0])

    def test_converts_to_hybrid_by_column_2(self, be, vt, it):
        # mtx2 keeps an explicit zero: [[1,3,2],[{0},5,0]]
        rp = np.array([0, 3, 5], dtype=IT[it])
        ci = np.array([0, 1, 2, 0, 1], dtype=IT[it])
        va = arr([1, 3, 2, 0, 5], vt)
        crp = np.full(3, -1, np.int64)
        be("csr_compute_hybrid_coo_row_ptrs_" + it, rp, 2, 2, crp)
        assert crp.tolist() == [0, 1, 1]
        ec, ev = np.full(4, 9, IT[it]), np.full(4, 9, VT[vt])
        cr, cc, cv = np.full(1, 9, IT[it]), np.full(1, 9, IT[it]), np.full(1, 9, VT[vt])
        be("csr_convert_to_hybrid_%s_%s" % (vt, it), 2, rp, ci, va, 2, 2, ec, ev, crp, cr, cc, cv)
        assert (cr.tolist(), cc.tolist(), cv.tolist()) == ([0], [2], [2.0])
        assert ev.tolist() == [1.0, 0.0, 3.0, 5.0] and ec.tolist() == [0, 0, 1, 1]

    def test_sort_unsorted_matrix(self, be, vt, it):
        rp = np.array([0, 2, 5, 7], dtype=IT[it])
        cols = np.array([2, 1, 1, 2, 0, 2, 0], dtype=IT[it])
        vals = arr([1, 2, 1, 8, 3, 3, 2], vt)
        be("csr_sort_by_column_index_%s_%s" % (vt, it), 3, rp, cols, vals)
        assert cols.tolist() == [1, 2, 0, 1, 2, 0, 2]
        assert vals.tolist() == [2.0, 1.0, 3.0, 1.0, 8.0, 2.0, 3.0]
