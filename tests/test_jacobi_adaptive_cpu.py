"""Adaptive-precision block-Jacobi without a GPU:
 1. the oracle's restatement (oracle/oracle_jacobi_adaptive.h, oracle_precision.h) against the REAL
    reference (gko::preconditioner::Jacobi with storage_optimization through oracle/_ref): chosen
    precisions, condition numbers, stored bits, apply, advanced apply and transpose are identical;
 2. the storage-type conversions against hand-checked values of gko::half / gko::truncated;
 3. the C++ host layer (preconditioner::Jacobi::with_storage_optimization / with_accuracy, transpose)
    on the mock: bit-identical to the oracle (= to the reference)."""
import numpy as np
import pytest
import torch

from tests import helpers as H
from tests import jacobi_cases as JC
from tests.helpers import VT

STORAGES = [JC.AUTODETECT, 0x01, 0x02, 0x10, 0x11, 0x20, "mixed", None]


def oracle_generate(orc, vt, rp, ci, va, ptrs, max_bs, storage, accuracy, fill=0):
    nb = len(ptrs) - 1
    bo, go, gp, space = JC.scheme(max_bs, nb)
    prec = JC.storage_request(storage, nb)
    cond = None if prec is None else np.zeros(nb, VT[vt])
    blocks = np.zeros(space, VT[vt])
    blocks.view(np.uint8)[:] = fill
    orc("jacobi_generate_adaptive_%s_i32" % vt, len(rp) - 1, rp, ci, va, nb, max_bs, float(accuracy), bo, go, gp,
        cond, prec, ptrs, blocks)
    return dict(block_offset=bo, group_offset=go, group_power=gp, blocks=blocks, precisions=prec,
                conditioning=cond, space=space)


def test_gko_half_and_truncated_conversions(orc):
    x = np.array([1.0, 65504.0, 65520.0, 1e6, -2.5, 0.1, 6.1035156e-5, 6.1e-5, 1e-8, -1e-8, 1.0009765625,
                  1.00048828125, 1.00146484375], np.float32)
    h = np.zeros(len(x), np.uint16)
    orc("float_to_gko_half", x, len(x), h)
    # normal range: IEEE round-to-nearest-even; below the smallest normal half: a signed zero
    # (include/ginkgo/core/base/half.hpp:417-419 "TODO: handle denormals"); 65520 rounds up to infinity
    assert list(h) == [0x3C00, 0x7BFF, 0x7C00, 0x7C00, 0xC100, 0x2E66, 0x0400, 0x0000, 0x0000, 0x8000, 0x3C01,
                       0x3C00, 0x3C02]
    back = np.zeros(len(x), np.float32)
    orc("gko_half_to_float", h, len(x), back)
    normal = np.abs(x) >= 6.1035156e-5
    assert np.array_equal(back[normal][:2], x[normal][:2])
    with np.errstate(over="ignore"):
        ieee = x.astype(np.float16)
    keep = normal & np.isfinite(ieee)
    assert np.array_equal(h[keep], ieee.view(np.uint16)[keep])
    sub = np.array([0x0001, 0x83FF], np.uint16)  # half denormals read back as signed zeros
    z = np.zeros(2, np.float32)
    orc("gko_half_to_float", sub, 2, z)
    assert z[0] == 0 and z[1] == 0 and np.signbit(z[1]) and not np.signbit(z[0])


ref = pytest.importorskip("oracle.ref")
import os  # noqa: E402

needs_ref = pytest.mark.skipif(not os.path.exists(ref.LIB_PATH), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("vt", ["f64", "f32"])
@pytest.mark.parametrize("max_bs", [2, 4, 7, 8, 13, 16, 32])
@pytest.mark.parametrize("storage", STORAGES)
def test_oracle_matches_the_real_reference(orc, vt, max_bs, storage):
    n = 200
    for accuracy in (0.1, 1e-3):
        rp, ci, va, ptrs = JC.make(n, max_bs, max_bs * 7 + int(accuracy * 1000), VT[vt])
        nb = len(ptrs) - 1
        so = storage if storage != "mixed" else JC.storage_request("mixed", nb)
        b = np.random.default_rng(9).uniform(-1, 1, (n, 3)).astype(VT[vt])
        x0 = np.random.default_rng(10).uniform(-1, 1, (n, 3)).astype(VT[vt])
        R = ref.jacobi_adaptive(rp, ci, va, max_bs, ptrs, so, accuracy, b=b)
        R2 = ref.jacobi_adaptive(rp, ci, va, max_bs, ptrs, so, accuracy, b=b, x=x0, alpha=-0.75, beta=1.5)
        RT = ref.jacobi_adaptive(rp, ci, va, max_bs, ptrs, so, accuracy, transposed=True)
        O = oracle_generate(orc, vt, rp, ci, va, ptrs, max_bs, storage, accuracy)
        assert (O["block_offset"], O["group_offset"], O["group_power"]) == (
            R["block_offset"], R["group_offset"], R["group_power"])
        if storage is None:
            assert R["precisions"] is None
        else:
            assert np.array_equal(O["precisions"], R["precisions"])
            assert np.array_equal(O["conditioning"], R["conditioning"])
        mask = JC.written_mask(ptrs, O["precisions"], vt == "f64", O["block_offset"], O["group_offset"],
                               O["group_power"], O["space"], VT[vt]().itemsize)
        assert np.array_equal(O["blocks"].view(np.uint8)[mask], R["blocks"].view(np.uint8)[mask])
        # what generate does not write is padding: a second run on a different fill leaves it alone
        O2 = oracle_generate(orc, vt, rp, ci, va, ptrs, max_bs, storage, accuracy, fill=0xAB)
        assert np.all(O2["blocks"].view(np.uint8)[~mask] == 0xAB)
        assert np.array_equal(O2["blocks"].view(np.uint8)[mask], O["blocks"].view(np.uint8)[mask])
        args = (nb, max_bs, O["block_offset"], O["group_offset"], O["group_power"], O["precisions"], ptrs,
                O["blocks"])
        x = np.zeros((n, 3), VT[vt])
        orc("jacobi_simple_apply_adaptive_%s_i32" % vt, *args, b, 3, 3, x, 3)
        assert np.array_equal(x, R["x"])
        x2 = x0.copy()
        orc("jacobi_apply_adaptive_%s_i32" % vt, *args, np.array([-0.75], VT[vt]), b, 3, 3,
            np.array([1.5], VT[vt]), x2, 3)
        assert np.array_equal(x2, R2["x"])
        bt = np.zeros(O["space"], VT[vt])
        orc("jacobi_transpose_adaptive_%s_i32" % vt, *args, bt)
        assert np.array_equal(bt.view(np.uint8)[mask], RT["blocks"].view(np.uint8)[mask])


@needs_ref
def test_autodetect_reaches_every_storage_type(orc):
    seen = {"f64": set(), "f32": set()}
    for vt in seen:
        for max_bs in (2, 7, 13, 32):
            for accuracy in (0.5, 0.1, 1e-3):
                rp, ci, va, ptrs = JC.make(200, max_bs, max_bs * 7 + int(accuracy * 1000), VT[vt])
                O = oracle_generate(orc, vt, rp, ci, va, ptrs, max_bs, JC.AUTODETECT, accuracy)
                seen[vt] |= set(int(p) for p in O["precisions"])
    assert {0x01, 0x02, 0x10, 0x11, 0x20} <= seen["f64"]
    assert {0x00, 0x02, 0x20} <= seen["f32"]


# ------------------------------------------------------------------ the C++ host layer on the mock
from tests.mock_build import CpuExec as _CpuExec  # noqa: E402


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    from ginkgo_b200 import api
    from tests.mock_build import build_mock_host
    lib = build_mock_host(str(tmp_path_factory.mktemp("mock")))
    saved = api._HOST_LIB
    api._HOST_LIB = lib
    yield _CpuExec(lib)
    api._HOST_LIB = saved


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("vt", ["f64", "f32"])
@pytest.mark.parametrize("max_bs", [4, 13, 16, 32])
@pytest.mark.parametrize("storage", [JC.AUTODETECT, 0x11, "mixed", None])
def test_host_jacobi_storage_optimization(orc, host, vt, max_bs, storage):
    from ginkgo_b200 import api
    n = 150
    rp, ci, va, ptrs = JC.make(n, max_bs, 100 + max_bs, VT[vt], singular_block=2 if max_bs == 13 else None)
    nb = len(ptrs) - 1
    so = storage if storage != "mixed" else JC.storage_request("mixed", nb)
    A = api.host_csr(host, (n, n), _t(va), _t(ci), _t(rp))
    J = api.host_jacobi(A, max_bs, ptrs, None if so is None else (so if np.isscalar(so) else list(so)), 0.05)
    G = api.host_jacobi_get(J)
    O = oracle_generate(orc, vt, rp, ci, va, ptrs, max_bs, storage, 0.05)
    assert (G["block_offset"], G["group_offset"], G["group_power"], G["num_blocks"]) == (
        O["block_offset"], O["group_offset"], O["group_power"], nb)
    assert np.array_equal(G["block_ptrs"], ptrs)
    if storage is None:
        assert G["precisions"] is None
    else:
        assert np.array_equal(G["precisions"], O["precisions"])
        assert np.array_equal(G["conditioning"], O["conditioning"], equal_nan=True)
    assert np.array_equal(G["blocks"].view(np.uint8), O["blocks"].view(np.uint8))  # the host layer zero-fills
    b = np.random.default_rng(3).uniform(-1, 1, (n, 2)).astype(VT[vt])
    x = np.zeros((n, 2), VT[vt])
    args = (nb, max_bs, O["block_offset"], O["group_offset"], O["group_power"], O["precisions"], ptrs, O["blocks"])
    orc("jacobi_simple_apply_adaptive_%s_i32" % vt, *args, b, 2, 2, x, 2)
    tb, tx = _t(b), torch.zeros(n, 2, dtype=_t(b).dtype)
    api.host_apply(J, api.host_dense(host, tb), api.host_dense(host, tx))
    assert np.array_equal(tx.numpy(), x, equal_nan=True)
    x2 = np.random.default_rng(4).uniform(-1, 1, (n, 2)).astype(VT[vt])
    tx2 = _t(x2).clone()
    al, be = np.array([[2.5]], VT[vt]), np.array([[-0.5]], VT[vt])
    orc("jacobi_apply_adaptive_%s_i32" % vt, *args, al.reshape(-1), b, 2, 2, be.reshape(-1), x2, 2)
    api.host_apply(J, api.host_dense(host, tb), api.host_dense(host, tx2), api.host_dense(host, _t(al)),
                   api.host_dense(host, _t(be)))
    assert np.array_equal(tx2.numpy(), x2, equal_nan=True)
    JT = api.host_jacobi_transpose(J)
    GT = api.host_jacobi_get(JT)
    bt = np.zeros(O["space"], VT[vt])
    orc("jacobi_transpose_adaptive_%s_i32" % vt, *args, bt)
    assert np.array_equal(GT["blocks"].view(np.uint8), bt.view(np.uint8))
    if storage is not None:
        assert np.array_equal(GT["precisions"], O["precisions"])
