"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/_ref/libgko_ref_shim.so -- the REAL
reference (its core + Reference/OMP executors compiled from /root/reference by
oracle/ref_build/Makefile) behind the small C ABI of oracle/ref_shim.cpp."""
import ctypes
import os

import numpy as np

from ginkgo_b200 import _cdecl

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libgko_ref_shim.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        l = ctypes.CDLL(LIB_PATH)
        src = open(os.path.join(_HERE, "ref_shim.cpp")).read()
        src = src[src.index('extern "C" {'):]
        decls = _cdecl.parse(src, "refshim_")
        _cdecl.bind(l, decls)
        _lib = l
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def _vt(a):
    return 0 if a.dtype == np.float64 else 1


def num_threads():
    return lib().refshim_num_threads()


def use_physical_cores():
    """reference CPU-baseline methodology: one OpenMP thread per
    physical core; returns the thread count in use"""
    try:
        import psutil
        n = psutil.cpu_count(logical=False) or os.cpu_count()
    except Exception:
        n = os.cpu_count()
    lib().refshim_set_num_threads(int(n))
    return num_threads()


def spmv(fmt, rp, ci, va, x, n_cols, alpha=None, beta=None, y=None, exec_kind=0, reps=0,
         strategy="classical"):
    """y = A x (or alpha A x + beta y) through the reference's LinOp::apply; returns (y, s/rep)"""
    n = len(rp) - 1
    x2 = np.ascontiguousarray(x).reshape(n_cols, -1)
    nrhs = x2.shape[1]
    if y is None:
        y = np.zeros((n, nrhs), dtype=va.dtype)
    sec = ctypes.c_double(0)
    a = None if alpha is None else np.array([alpha], va.dtype)
    b = None if beta is None else np.array([beta], va.dtype)
    st = lib().refshim_spmv(exec_kind, fmt.encode(), _vt(va), n, n_cols, len(va), _p(rp), _p(ci),
                            _p(va), _p(x2), nrhs, _p(y), _p(a), _p(b), reps, ctypes.byref(sec),
                            strategy.encode())
    assert st == 0, st
    return y, sec.value


def solve(kind, rp, ci, va, b, x0, precond_max_bs=0, block_ptrs=None, max_iters=-1, res_kind=1,
          baseline=0, reduction=1e-8, iter_first=1, krylov_dim=30, ortho=0, exec_kind=0,
          relaxation_factor=1.0, foci=(0.0, 1.0)):
    n = len(rp) - 1
    b2 = np.ascontiguousarray(b).reshape(n, -1)
    x = np.ascontiguousarray(x0).reshape(n, -1).copy()
    nrhs = b2.shape[1]
    iters = ctypes.c_int64(0)
    sec = ctypes.c_double(0)
    resn = np.zeros(nrhs, va.dtype)
    nb = 0 if block_ptrs is None else len(block_ptrs) - 1
    lib().refshim_solve_params(float(relaxation_factor), float(foci[0]), float(foci[1]))
    st = lib().refshim_solve(exec_kind, {"cg": 0, "bicgstab": 1, "gmres": 2, "fcg": 3, "cgs": 4, "ir": 5,
                                         "chebyshev": 6, "pipe_cg": 7, "gcr": 8, "minres": 9}[kind], _vt(va), n,
                             len(va), _p(rp), _p(ci), _p(va), _p(b2), _p(x), nrhs, precond_max_bs,
                             _p(block_ptrs), nb, max_iters, res_kind, baseline, reduction,
                             iter_first, krylov_dim, ortho, ctypes.byref(iters), _p(resn),
                             ctypes.byref(sec))
    assert st == 0, st
    return x, iters.value, resn, sec.value


def jacobi_generate(rp, ci, va, max_bs, block_ptrs=None):
    n = len(rp) - 1
    nb = 0 if block_ptrs is None else len(block_ptrs) - 1
    cap = 32 * 32 * (n + 64) if max_bs > 1 else n
    blocks = np.zeros(cap, va.dtype)
    meta = np.zeros(5, np.int64)
    ptrs = np.zeros(n + 2, np.int32)
    st = lib().refshim_jacobi_generate(_vt(va), n, len(va), _p(rp), _p(ci), _p(va), max_bs,
                                       _p(block_ptrs), nb, _p(blocks), cap, _p(meta), _p(ptrs))
    assert st == 0, st
    return dict(block_offset=int(meta[0]), group_offset=int(meta[1]), group_power=int(meta[2]),
                num_blocks=int(meta[3]), blocks=blocks[:meta[4]].copy(),
                block_ptrs=ptrs[:meta[3] + 1].copy())


def convert(fmt, rp, ci, va, n_cols, slice_size=64, stride_factor=1, strategy=0, columns=0,
            percent=0.8, ratio=0.0001):
    """Csr::convert_to(Ell|Sellp|Hybrid) / Csr::sort_by_column_index of the real reference.
    Returns a dict of the result arrays (see oracle/ref_shim.cpp `convert_impl`)."""
    n, nnz = len(rp) - 1, len(va)
    lens = np.diff(rp.astype(np.int64)) if n else np.zeros(0, np.int64)
    width = int(lens.max()) if n else 0
    pad = (width + stride_factor) if fmt == "sellp" else max(width, int(columns))
    rows_pad = (n + slice_size) if fmt == "sellp" else n
    cap = max(1, pad * rows_pad, nnz)
    vals = np.zeros(cap, va.dtype)
    cols = np.zeros(cap, np.int32)
    ns = (n + slice_size - 1) // slice_size
    aux0, aux1 = np.zeros(ns + 2, np.uint64), np.zeros(ns + 1, np.uint64)
    crows, ccols, cvals = np.zeros(nnz + 1, np.int32), np.zeros(nnz + 1, np.int32), np.zeros(nnz + 1, va.dtype)
    meta = np.zeros(4, np.int64)
    if fmt == "sellp":
        p0, p1 = slice_size, stride_factor
    else:
        p0, p1 = strategy, columns
    st = lib().refshim_convert(fmt.encode(), _vt(va), n, n_cols, nnz, _p(rp), _p(ci), _p(va), p0, p1,
                               float(percent), float(ratio), _p(meta), _p(vals), _p(cols), cap,
                               _p(aux0), _p(aux1), _p(crows), _p(ccols), _p(cvals), nnz + 1)
    assert st == 0, st
    if fmt == "sort":
        return dict(cols=cols[:nnz], vals=vals[:nnz])
    if fmt == "ell":
        w, stride = int(meta[0]), int(meta[1])
        return dict(width=w, stride=stride, cols=cols[:w * stride], vals=vals[:w * stride])
    if fmt == "sellp":
        tot = int(meta[1])
        return dict(slice_sets=aux0[:ns + 1], slice_lengths=aux1[:ns], cols=cols[:tot], vals=vals[:tot])
    w, stride, cn = int(meta[0]), int(meta[1]), int(meta[2])
    return dict(ell_lim=w, ell_stride=stride, coo_nnz=cn, cols=cols[:w * stride],
                vals=vals[:w * stride], coo_rows=crows[:cn], coo_cols=ccols[:cn], coo_vals=cvals[:cn])


def read_mtx(path, cap=1 << 20):
    """gko::read_generic_raw<double,int32>: (rows, cols, [(r, c, v) ...]) in the reference's order"""
    meta = np.zeros(3, np.int64)
    r, c, v = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float64)
    st = lib().refshim_read_mtx(str(path).encode(), _p(meta), _p(r), _p(c), _p(v), cap)
    if st != 0:
        raise RuntimeError("reference could not read %s (status %d)" % (path, st))
    n = int(meta[2])
    return int(meta[0]), int(meta[1]), r[:n].copy(), c[:n].copy(), v[:n].copy()


def write_mtx(path, layout, nrows, ncols, rows, cols, vals):
    """gko::write_raw (layout 'coordinate' | 'array') / write_binary_raw ('binary')"""
    rows = np.ascontiguousarray(rows, np.int32)
    cols = np.ascontiguousarray(cols, np.int32)
    vals = np.ascontiguousarray(vals, np.float64)
    st = lib().refshim_write_mtx(str(path).encode(), {"coordinate": 0, "array": 1, "binary": 2}[layout],
                                 nrows, ncols, len(vals), _p(rows), _p(cols), _p(vals))
    assert st == 0, st
