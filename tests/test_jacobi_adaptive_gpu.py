"""Adaptive-precision block-Jacobi on the B200: generate with conditioning and
precision detection, apply / advanced apply in the stored precision and transpose, through the C ABI,
against the oracle (which tests/test_jacobi_adaptive_cpu.py pins to the real reference): chosen
precisions, condition numbers and stored bits identical; apply bit-identical (same operation order,
-fmad=false).  Then the same through the C++ host layer (Jacobi::with_storage_optimization)."""
import numpy as np
import pytest

from tests import helpers as H
from tests import jacobi_cases as JC
from tests.helpers import IT, VT

pytestmark = pytest.mark.gpu

STORAGES = [JC.AUTODETECT, 0x01, 0x02, 0x10, 0x11, 0x20, "mixed", None]


def generate(backend, vt, it, rp, ci, va, ptrs, max_bs, storage, accuracy, with_cond=True):
    nb = len(ptrs) - 1
    bo, go, gp, space = JC.scheme(max_bs, nb)
    prec = JC.storage_request(storage, nb)
    cond = np.zeros(nb, VT[vt]) if (prec is not None and with_cond) else None
    blocks = np.zeros(space, VT[vt])
    backend("jacobi_generate_adaptive_%s_%s" % (vt, it), len(rp) - 1, rp, ci, va, nb, max_bs, float(accuracy), bo,
            go, gp, cond, prec, ptrs, blocks)
    return dict(block_offset=bo, group_offset=go, group_power=gp, blocks=blocks, precisions=prec,
                conditioning=cond, space=space)


@pytest.mark.parametrize("vt", ["f64", "f32"])
@pytest.mark.parametrize("it", ["i32", "i64"])
@pytest.mark.parametrize("max_bs", [1, 2, 4, 7, 8, 13, 16, 23, 32])
@pytest.mark.parametrize("storage", STORAGES)
def test_generate_apply_transpose_match_the_oracle(orc, cuda, vt, it, max_bs, storage):
    if it == "i64" and max_bs not in (7, 16, 32):
        pytest.skip("int64 indices: three block sizes are enough")
    n = 777
    for accuracy in (0.1, 1e-3):
        rp, ci, va, ptrs = JC.make(n, max_bs, max_bs * 7 + int(accuracy * 1000), VT[vt], IT[it],
                                   singular_block=5 if max_bs == 13 else None)
        nb = len(ptrs) - 1
        O = generate(orc, vt, it, rp, ci, va, ptrs, max_bs, storage, accuracy)
        C = generate(cuda, vt, it, rp, ci, va, ptrs, max_bs, storage, accuracy)
        if storage is not None:
            assert np.array_equal(C["precisions"], O["precisions"])
            assert np.array_equal(C["conditioning"], O["conditioning"], equal_nan=True)
        assert np.array_equal(C["blocks"].view(np.uint8), O["blocks"].view(np.uint8))
        args = lambda G: (nb, max_bs, G["block_offset"], G["group_offset"], G["group_power"], G["precisions"], ptrs,
                          G["blocks"])
        for nrhs in (1, 3):
            b = np.random.default_rng(9).uniform(-1, 1, (n, nrhs)).astype(VT[vt])
            xo, xc = np.zeros((n, nrhs), VT[vt]), np.zeros((n, nrhs), VT[vt])
            orc("jacobi_simple_apply_adaptive_%s_%s" % (vt, it), *args(O), b, nrhs, nrhs, xo, nrhs)
            cuda("jacobi_simple_apply_adaptive_%s_%s" % (vt, it), *args(C), b, nrhs, nrhs, xc, nrhs)
            assert np.array_equal(xc, xo, equal_nan=True)
            x0 = np.random.default_rng(10).uniform(-1, 1, (n, nrhs)).astype(VT[vt])
            al, be = np.array([-0.75], VT[vt]), np.array([1.5], VT[vt])
            xo, xc = x0.copy(), x0.copy()
            orc("jacobi_apply_adaptive_%s_%s" % (vt, it), *args(O), al, b, nrhs, nrhs, be, xo, nrhs)
            cuda("jacobi_apply_adaptive_%s_%s" % (vt, it), *args(C), al, b, nrhs, nrhs, be, xc, nrhs)
            assert np.array_equal(xc, xo, equal_nan=True)
        to, tc = np.zeros(O["space"], VT[vt]), np.zeros(O["space"], VT[vt])
        orc("jacobi_transpose_adaptive_%s_%s" % (vt, it), *args(O), to)
        cuda("jacobi_transpose_adaptive_%s_%s" % (vt, it), *args(C), tc)
        assert np.array_equal(tc.view(np.uint8), to.view(np.uint8))


@pytest.mark.parametrize("case", range(8))
def test_device_matches_the_committed_reference_outputs(cuda, case):
    """the device against tests/golden/jacobi_adaptive_reference.json (outputs of the REAL reference for
    deterministic inputs): precisions, condition numbers, stored bytes, apply and advanced apply bit for bit"""
    JC.check_against_golden(cuda, JC.load_golden()[case])


@pytest.mark.parametrize("vt", ["f64", "f32"])
def test_no_conditioning_means_no_detection(orc, cuda, vt):
    """reference/preconditioner/jacobi_kernels.cpp:357: autodetect() without a conditioning array
    falls back to the singleton of the byte, i.e. full precision"""
    rp, ci, va, ptrs = JC.make(300, 8, 1, VT[vt])
    O = generate(orc, vt, "i32", rp, ci, va, ptrs, 8, JC.AUTODETECT, 0.1, with_cond=False)
    C = generate(cuda, vt, "i32", rp, ci, va, ptrs, 8, JC.AUTODETECT, 0.1, with_cond=False)
    assert np.all(C["precisions"] == 0) and np.array_equal(C["precisions"], O["precisions"])
    assert np.array_equal(C["blocks"].view(np.uint8), O["blocks"].view(np.uint8))


def test_initialize_precisions(orc, cuda):
    src = np.array([0xFF, 0x01, 0x20], np.uint8)
    a, b = np.zeros(1000, np.uint8), np.zeros(1000, np.uint8)
    orc("jacobi_initialize_precisions", src, 3, a, 1000)
    cuda("jacobi_initialize_precisions", src, 3, b, 1000)
    assert np.array_equal(a, b) and np.array_equal(b[:6], [0xFF, 1, 0x20, 0xFF, 1, 0x20])


@pytest.mark.parametrize("vt", ["f64", "f32"])
def test_cfg4_sized_generate(orc, cuda, vt):
    """cfg4's shape scaled down: uniform 16 x 16 blocks, every group full, autodetect"""
    n = 16 * 3000
    rng = np.random.default_rng(77)
    rp, ci, va = H.random_csr(rng, n, n, np.full(n, 20), vt, "i32")
    for r in range(n):  # diagonally dominant as in cfg4
        s, e = int(rp[r]), int(rp[r + 1])
        d = np.where(ci[s:e] == r)[0]
        if len(d):
            va[s + d[0]] = np.abs(va[s:e]).sum() + 1
    ptrs = np.arange(0, n + 1, 16, dtype=np.int32)
    O = generate(orc, vt, "i32", rp, ci, va, ptrs, 16, JC.AUTODETECT, 0.1)
    C = generate(cuda, vt, "i32", rp, ci, va, ptrs, 16, JC.AUTODETECT, 0.1)
    assert np.array_equal(C["precisions"], O["precisions"])
    assert np.array_equal(C["conditioning"], O["conditioning"])
    assert np.array_equal(C["blocks"].view(np.uint8), O["blocks"].view(np.uint8))


# ------------------------------------------------------------------ through the C++ host layer
@pytest.fixture(scope="module")
def hexec():
    from ginkgo_b200 import api
    import os
    if os.environ.get("B200_TEST_SELFCHECK") == "1":
        pytest.skip("host-layer body is covered by tests/test_jacobi_adaptive_cpu.py on the mock")
    return api.HostExecutor(0)


@pytest.mark.parametrize("vt", ["f64", "f32"])
@pytest.mark.parametrize("max_bs,storage", [(16, JC.AUTODETECT), (13, "mixed"), (32, 0x11), (8, None)])
def test_host_layer_jacobi_on_the_device(orc, hexec, vt, max_bs, storage):
    import torch
    from ginkgo_b200 import api
    n = 500
    rp, ci, va, ptrs = JC.make(n, max_bs, 200 + max_bs, VT[vt])
    nb = len(ptrs) - 1
    so = storage if storage != "mixed" else JC.storage_request("mixed", nb)
    O = generate(orc, vt, "i32", rp, ci, va, ptrs, max_bs, storage, 0.05)
    dev = hexec.device
    with torch.cuda.stream(hexec.stream):
        t = [torch.from_numpy(a).to(dev) for a in (va, ci, rp)]
        A = api.host_csr(hexec, (n, n), *t)
        J = api.host_jacobi(A, max_bs, ptrs, None if so is None else (so if np.isscalar(so) else list(so)), 0.05)
        G = api.host_jacobi_get(J)
        if storage is not None:
            assert np.array_equal(G["precisions"], O["precisions"])
            assert np.array_equal(G["conditioning"], O["conditioning"])
        assert np.array_equal(G["blocks"].view(np.uint8), O["blocks"].view(np.uint8))
        b = np.random.default_rng(3).uniform(-1, 1, (n, 2)).astype(VT[vt])
        x = np.zeros((n, 2), VT[vt])
        orc("jacobi_simple_apply_adaptive_%s_i32" % vt, nb, max_bs, O["block_offset"], O["group_offset"],
            O["group_power"], O["precisions"], ptrs, O["blocks"], b, 2, 2, x, 2)
        tb = torch.from_numpy(b).to(dev)
        tx = torch.zeros(n, 2, dtype=tb.dtype, device=dev)
        api.host_apply(J, api.host_dense(hexec, tb), api.host_dense(hexec, tx))
        hexec.synchronize()
        assert np.array_equal(tx.cpu().numpy(), x)
        JT = api.host_jacobi_transpose(J)
        GT = api.host_jacobi_get(JT)
        bt = np.zeros(O["space"], VT[vt])
        orc("jacobi_transpose_adaptive_%s_i32" % vt, nb, max_bs, O["block_offset"], O["group_offset"],
            O["group_power"], O["precisions"], ptrs, O["blocks"], bt)
        assert np.array_equal(GT["blocks"].view(np.uint8), bt.view(np.uint8))


# ------------------------------------------------------------------ many right-hand sides: tensor cores
@pytest.mark.parametrize("vt", ["f64", "f32"])
@pytest.mark.parametrize("max_bs", [5, 8, 13, 16, 32])
@pytest.mark.parametrize("nrhs", [8, 13, 32, 64])
@pytest.mark.parametrize("storage", [None, JC.AUTODETECT])
def test_many_rhs_apply_on_the_fp64_tensor_cores(orc, cuda, vt, max_bs, nrhs, storage):
    """num_rhs >= 8: block_apply_mma_kernel (mma.sync.m8n8k4.f64, fp64 accumulation).  The inner
    products are summed in another order than the reference's sequential loop: r<T>-scaled tolerance."""
    n = 999
    rp, ci, va, ptrs = JC.make(n, max_bs, 300 + max_bs, VT[vt])
    nb = len(ptrs) - 1
    O = generate(orc, vt, "i32", rp, ci, va, ptrs, max_bs, storage, 0.1)
    args = (nb, max_bs, O["block_offset"], O["group_offset"], O["group_power"], O["precisions"], ptrs, O["blocks"])
    b = np.random.default_rng(19).uniform(-1, 1, (n, nrhs)).astype(VT[vt])
    ld = nrhs + 3  # strided operands
    bpad = np.zeros((n, ld), VT[vt])
    bpad[:, :nrhs] = b
    xo = np.zeros((n, nrhs), VT[vt])
    orc("jacobi_simple_apply_adaptive_%s_i32" % vt, *args, b, nrhs, nrhs, xo, nrhs)
    xc = np.full((n, ld), 7.0, VT[vt])
    cuda("jacobi_simple_apply_adaptive_%s_i32" % vt, *args, bpad, ld, nrhs, xc, ld)
    assert H.rel_err(xc[:, :nrhs], xo) <= 4 * H.R[vt]
    assert np.all(xc[:, nrhs:] == 7.0)  # nothing written past the last right-hand side
    x0 = np.random.default_rng(20).uniform(-1, 1, (n, nrhs)).astype(VT[vt])
    al, be = np.array([-0.75], VT[vt]), np.array([1.5], VT[vt])
    xo2, xc2 = x0.copy(), x0.copy()
    orc("jacobi_apply_adaptive_%s_i32" % vt, *args, al, b, nrhs, nrhs, be, xo2, nrhs)
    cuda("jacobi_apply_adaptive_%s_i32" % vt, *args, al, b, nrhs, nrhs, be, xc2, nrhs)
    assert H.rel_err(xc2, xo2) <= 4 * H.R[vt]
    if storage is None:  # the plain entry points take the same kernel
        xc3 = np.zeros((n, nrhs), VT[vt])
        cuda("jacobi_simple_apply_%s_i32" % vt, nb, max_bs, O["block_offset"], O["group_offset"], O["group_power"],
             ptrs, O["blocks"], b, nrhs, nrhs, xc3, nrhs)
        assert np.array_equal(xc3, xc[:, :nrhs])
        # beta == 0 never reads x (NaN in, clean out)
        xn = np.full((n, nrhs), np.nan, VT[vt])
        cuda("jacobi_apply_%s_i32" % vt, nb, max_bs, O["block_offset"], O["group_offset"], O["group_power"], ptrs,
             O["blocks"], al, b, nrhs, nrhs, np.array([0.0], VT[vt]), xn, nrhs)
        assert np.all(np.isfinite(xn))
