"""Known-answer tests transcribed from the reference's own suites for the paths added late in
round 1 (BiCG, Csr::transpose): the literal vectors run on the oracle (CPU) and -- marked gpu --
on the CUDA library through the C ABI.  Kept apart from tests/test_golden.py, in a file that
sorts late, because the gpu variants have not run on a B200 yet.  The distributed set-up literals
(partition / separate_local_nonlocal / index_map) are in tests/test_dist_assembly_cpu.py.
Citations are relative to /root/reference."""
import numpy as np
import pytest

from tests import helpers as H
from tests.helpers import IT, R, VT, rel_err

BACKENDS = ["oracle", pytest.param("cuda", marks=H.first_gpu_run_marks())]
VTS = ["f64", "f32"]
ITS = ["i32", "i64"]


@pytest.fixture(params=BACKENDS)
def be(request):
    return request.getfixturevalue("orc" if request.param == "oracle" else "cuda")


def arr(x, vt):
    return np.array(x, dtype=VT[vt])


def dense_to_csr(rows, vt, it="i32"):
    a = np.array(rows, dtype=VT[vt])
    rp, ci, va = [0], [], []
    for r in a:
        nz = np.nonzero(r)[0]
        ci += list(nz)
        va += list(r[nz])
        rp.append(len(ci))
    return np.array(rp, IT[it]), np.array(ci, IT[it]), np.array(va, VT[vt])


def csr_to_dense(n, m, rp, ci, va):
    out = np.zeros((n, m), va.dtype)
    for r in range(n):
        out[r, ci[rp[r]:rp[r + 1]]] = va[rp[r]:rp[r + 1]]
    return out


@pytest.mark.parametrize("vt", VTS)
@pytest.mark.parametrize("it", ITS)
class TestTranspose:
    def _transpose(self, be, vt, it, rows):
        rp, ci, va = dense_to_csr(rows, vt, it)
        n, m = len(rows), len(rows[0])
        trp, tci, tva = np.zeros(m + 1, IT[it]), np.zeros(len(va), IT[it]), np.zeros(len(va), VT[vt])
        be("csr_transpose_%s_%s" % (vt, it), n, m, len(va), rp, ci, va, trp, tci, tva)
        return csr_to_dense(m, n, trp, tci, tva), trp

    def test_square_mtx_is_transposable(self, be, vt, it):
        # reference/test/matrix/csr_kernels.cpp:1613-1631
        t, trp = self._transpose(be, vt, it, [[1.0, 3.0, 2.0], [0.0, 5.0, 0.0], [0.0, 1.5, 2.0]])
        assert t.tolist() == [[1.0, 0.0, 0.0], [3.0, 5.0, 1.5], [2.0, 0.0, 2.0]]
        assert trp.tolist() == [0, 1, 4, 6]

    def test_non_square_mtx_is_transposable(self, be, vt, it):
        # reference/test/matrix/csr_kernels.cpp:1634-1645 (the 2 x 3 fixture of :84-106)
        t, trp = self._transpose(be, vt, it, [[1.0, 3.0, 2.0], [0.0, 5.0, 0.0]])
        assert t.tolist() == [[1.0, 0.0], [3.0, 5.0], [2.0, 0.0]]
        assert trp.tolist() == [0, 1, 3, 4]


@pytest.mark.parametrize("vt", VTS)
class TestBicgKernels:
    # the fixture of reference/test/solver/bicg_kernels.cpp:70-110: 2 x 2 vectors with padded
    # strides, stop status of column 1 optionally "stopped"
    def _v(self, vt, fill, stride=2):
        a = np.full((2, stride), 99.0, VT[vt])
        a[:, :2] = fill
        return a

    def test_step_1(self, be, vt):
        # :163-182
        p, z, p2, z2 = self._v(vt, 3), self._v(vt, -2), self._v(vt, 3, 3), self._v(vt, -2, 3)
        rho, prev_rho = arr([2, 3], vt), arr([8, 3], vt)
        stop = np.array([0, 1], np.uint8)  # stopping_status::stop(1)
        be("bicg_step_1_" + vt, 2, 2, p, 2, z, 2, p2, 3, z2, 3, rho, prev_rho, stop)
        assert p[:, :2].tolist() == [[-1.25, 3.0], [-1.25, 3.0]]
        assert p2[:, :2].tolist() == [[-1.25, 3.0], [-1.25, 3.0]]
        assert (p2[:, 2] == 99.0).all()

    def test_step_1_div_by_zero(self, be, vt):
        # :185-201
        p, z, p2, z2 = self._v(vt, 3), self._v(vt, -2), self._v(vt, 3), self._v(vt, -2)
        stop = np.zeros(2, np.uint8)
        be("bicg_step_1_" + vt, 2, 2, p, 2, z, 2, p2, 2, z2, 2, arr([1, 1], vt), arr([0, 0], vt), stop)
        assert p.tolist() == [[-2.0, -2.0], [-2.0, -2.0]] and p2.tolist() == p.tolist()

    def test_step_2(self, be, vt):
        # :204-227
        x, p, r, q = self._v(vt, -2, 4), self._v(vt, 3), self._v(vt, 4), self._v(vt, -5)
        r2, q2 = self._v(vt, 4), self._v(vt, -5)
        stop = np.array([0, 1], np.uint8)
        be("bicg_step_2_" + vt, 2, 2, x, 4, r, 2, r2, 2, p, 2, q, 2, q2, 2, arr([8, 3], vt), arr([2, 3], vt), stop)
        assert x[:, :2].tolist() == [[-1.25, -2.0], [-1.25, -2.0]]
        assert r.tolist() == [[5.25, 4.0], [5.25, 4.0]] and r2.tolist() == r.tolist()

    def test_step_2_div_by_zero(self, be, vt):
        # :230-250
        x, p, r, q = self._v(vt, -2), self._v(vt, 3), self._v(vt, 4), self._v(vt, -5)
        r2, q2 = self._v(vt, 4), self._v(vt, -5)
        stop = np.zeros(2, np.uint8)
        be("bicg_step_2_" + vt, 2, 2, x, 2, r, 2, r2, 2, p, 2, q, 2, q2, 2, arr([0, 0], vt), arr([1, 1], vt), stop)
        assert x.tolist() == [[-2.0, -2.0], [-2.0, -2.0]]
        assert r.tolist() == [[4.0, 4.0], [4.0, 4.0]] and r2.tolist() == r.tolist()

    def test_initialize(self, be, vt):
        # :125-160
        b = arr([[1, 2], [3, 4]], vt)
        v = {k: self._v(vt, 7) for k in ("r", "z", "p", "q", "r2", "z2", "p2", "q2")}
        rho, prev_rho = arr([5, 5], vt), arr([5, 5], vt)
        stop = np.full(2, 0x81, np.uint8)
        be("bicg_initialize_" + vt, 2, 2, b, 2, v["r"], 2, v["z"], 2, v["p"], 2, v["q"], 2, prev_rho, rho,
           v["r2"], 2, v["z2"], 2, v["p2"], 2, v["q2"], 2, stop)
        assert v["r"].tolist() == b.tolist() and v["r2"].tolist() == b.tolist()
        for k in ("z", "p", "q", "z2", "p2", "q2"):
            assert not v[k].any()
        assert rho.tolist() == [0, 0] and prev_rho.tolist() == [1, 1] and not stop.any()


# the solver literals run through the oracle's BiCG loop (host loop, CPU only: the C++ host loop is
# checked against it in tests/test_host_cpu.py and on the GPU in tests/test_zzz_dist_assembly_gpu.py)
@pytest.mark.parametrize("vt", VTS)
class TestBicgSolves:
    def test_solves_stencil_system(self, vt):
        # reference/test/solver/bicg_kernels.cpp:253-264; criteria Iteration(4) + ResidualNorm(r<T>)
        rp, ci, va = dense_to_csr([[2, -1.0, 0.0], [-1.0, 2, -1.0], [0.0, -1.0, 2]], vt)
        x, it, _ = H.orc_solve("bicg", vt, rp, ci, va, arr([[-1], [3], [1]], vt), np.zeros((3, 1), VT[vt]),
                               max_iters=4, reduction=R[vt])
        assert rel_err(x[:, 0], arr([1, 3, 2], vt)) <= R[vt]

    def test_solves_multiple_stencil_systems(self, vt):
        # :325-340
        rp, ci, va = dense_to_csr([[2, -1.0, 0.0], [-1.0, 2, -1.0], [0.0, -1.0, 2]], vt)
        b = arr([[-1, 1], [3, 0], [1, 1]], vt)
        x, _, _ = H.orc_solve("bicg", vt, rp, ci, va, b, np.zeros((3, 2), VT[vt]), max_iters=4, reduction=R[vt])
        assert rel_err(x, arr([[1, 1], [3, 1], [2, 1]], vt)) <= R[vt]

    def test_solves_non_symmetric_stencil_system(self, vt):
        # :502-513
        rp, ci, va = dense_to_csr([[1.0, 2.0, 3.0], [3.0, 2.0, -1.0], [0.0, -1.0, 2]], vt)
        x, _, _ = H.orc_solve("bicg", vt, rp, ci, va, arr([[13], [7], [1]], vt), np.zeros((3, 1), VT[vt]),
                               max_iters=4, reduction=R[vt])
        assert rel_err(x[:, 0], arr([1, 3, 2], vt)) <= R[vt] * 1e2

    @pytest.mark.parametrize("res_kind", [1, 2])
    def test_solves_big_dense_system(self, vt, res_kind):
        # :445-461 (ResidualNorm) and :483-499 (ImplicitResidualNorm), Iteration(100)
        m = [[8828.0, 2673.0, 4150.0, -3139.5, 3829.5, 5856.0], [2673.0, 10765.5, 1805.0, 73.0, 1966.0, 3919.5],
             [4150.0, 1805.0, 6472.5, 2656.0, 2409.5, 3836.5], [-3139.5, 73.0, 2656.0, 6048.0, 665.0, -132.0],
             [3829.5, 1966.0, 2409.5, 665.0, 4240.5, 4373.5], [5856.0, 3919.5, 3836.5, -132.0, 4373.5, 5678.0]]
        rp, ci, va = dense_to_csr(m, vt)
        b = arr([[1300083.0], [1018120.5], [906410.0], [-42679.5], [846779.5], [1176858.5]], vt)
        x, _, _ = H.orc_solve("bicg", vt, rp, ci, va, b, np.zeros((6, 1), VT[vt]), max_iters=100,
                               reduction=R[vt], res_kind=res_kind)
        assert rel_err(x[:, 0], arr([81.0, 55.0, 45.0, 5.0, 85.0, -10.0], vt)) <= R[vt] * 1e2
