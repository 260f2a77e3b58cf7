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
          relaxation_factor=1.0, foci=(0.0, 1.0), initial_guess="provided"):
    n = len(rp) - 1
    b2 = np.ascontiguousarray(b).reshape(n, -1)
    x = np.ascontiguousarray(x0).reshape(n, -1).copy()
    nrhs = b2.shape[1]
    iters = ctypes.c_int64(0)
    sec = ctypes.c_double(0)
    resn = np.zeros(nrhs, va.dtype)
    nb = 0 if block_ptrs is None else len(block_ptrs) - 1
    lib().refshim_solve_params(float(relaxation_factor), float(foci[0]), float(foci[1]))
    if hasattr(lib(), "refshim_solve_guess"):
        lib().refshim_solve_guess({"provided": 0, "zero": 1, "rhs": 2}[initial_guess])
    st = lib().refshim_solve(exec_kind, {"cg": 0, "bicgstab": 1, "gmres": 2, "fcg": 3, "cgs": 4, "ir": 5,
                                         "chebyshev": 6, "pipe_cg": 7, "gcr": 8, "minres": 9, "bicg": 10}[kind], _vt(va), n,
                             len(va), _p(rp), _p(ci), _p(va), _p(b2), _p(x), nrhs, precond_max_bs,
                             _p(block_ptrs), nb, max_iters, res_kind, baseline, reduction,
                             iter_first, krylov_dim, ortho, ctypes.byref(iters), _p(resn),
                             ctypes.byref(sec))
    assert st == 0, st
    return x, iters.value, resn, sec.value


def jacobi_generate(rp, ci, va, max_bs, block_ptrs=None, transposed=False):
    """transposed: the blocks of Jacobi::transpose() of the generated preconditioner"""
    n = len(rp) - 1
    nb = 0 if block_ptrs is None else len(block_ptrs) - 1
    cap = 32 * 32 * (n + 64) if max_bs > 1 else n
    blocks = np.zeros(cap, va.dtype)
    meta = np.zeros(5, np.int64)
    ptrs = np.zeros(n + 2, np.int32)
    fn = lib().refshim_jacobi_generate_transposed if transposed else lib().refshim_jacobi_generate
    st = fn(_vt(va), n, len(va), _p(rp), _p(ci), _p(va), max_bs, _p(block_ptrs), nb, _p(blocks), cap,
            _p(meta), _p(ptrs))
    assert st == 0, st
    return dict(block_offset=int(meta[0]), group_offset=int(meta[1]), group_power=int(meta[2]),
                num_blocks=int(meta[3]), blocks=blocks[:meta[4]].copy(),
                block_ptrs=ptrs[:meta[3] + 1].copy())


def jacobi_adaptive(rp, ci, va, max_bs, block_ptrs=None, storage=None, accuracy=0.1, transposed=False,
                    b=None, x=None, alpha=None, beta=None):
    """gko::preconditioner::Jacobi with storage_optimization of the real reference.
    storage: None | one byte (0xff = autodetect, else preserving << 4 | nonpreserving) | uint8 array
    (block-wise).  b (n x nrhs): also apply; x = J b, or x = alpha J b + beta x when alpha is given.
    Returns the storage scheme, the raw block storage (as the value type), the per-block precisions
    and condition numbers, and x."""
    n = len(rp) - 1
    nb = 0 if block_ptrs is None else len(block_ptrs) - 1
    cap = 32 * 32 * (n + 64)
    blocks = np.zeros(cap, va.dtype)
    meta = np.zeros(6, np.int64)
    ptrs = np.zeros(n + 2, np.int32)
    prec = np.zeros(n + 1, np.uint8)
    cond = np.zeros(n + 1, va.dtype)
    if storage is None:
        so_kind, so = 0, np.zeros(1, np.uint8)
    elif np.isscalar(storage):
        so_kind, so = 1, np.array([storage], np.uint8)
    else:
        so_kind, so = 2, np.ascontiguousarray(storage, dtype=np.uint8)
    nrhs = 0
    xo = None
    if b is not None:
        b2 = np.ascontiguousarray(b.reshape(n, -1))
        nrhs = b2.shape[1]
        xo = np.zeros((n, nrhs), va.dtype) if x is None else np.ascontiguousarray(x.reshape(n, -1)).copy()
    al = None if alpha is None else np.array([alpha], va.dtype)
    be = None if beta is None else np.array([beta], va.dtype)
    st = lib().refshim_jacobi_adaptive(_vt(va), n, len(va), _p(rp), _p(ci), _p(va), max_bs, _p(block_ptrs), nb,
                                       so_kind, _p(so), len(so), float(accuracy), _p(blocks), cap, _p(meta),
                                       _p(ptrs), _p(prec), _p(cond), int(transposed), nrhs,
                                       _p(b2) if b is not None else None, _p(xo) if b is not None else None,
                                       _p(al), _p(be))
    assert st == 0, st
    nblk = int(meta[3])
    return dict(block_offset=int(meta[0]), group_offset=int(meta[1]), group_power=int(meta[2]),
                num_blocks=nblk, blocks=blocks[:meta[4]].copy(), block_ptrs=ptrs[:nblk + 1].copy(),
                precisions=prec[:nblk].copy() if meta[5] else None, conditioning=cond[:nblk].copy(), x=xo)


def csr_transpose(rp, ci, va, n_cols):
    """Csr::transpose() of the real reference -> (row_ptrs, col_idxs, values) of the transpose"""
    n = len(rp) - 1
    trp = np.zeros(n_cols + 1, np.int32)
    tci = np.zeros(max(len(va), 1), np.int32)
    tva = np.zeros(max(len(va), 1), va.dtype)
    st = lib().refshim_csr_transpose(_vt(va), n, n_cols, len(va), _p(rp), _p(ci), _p(va), _p(trp), _p(tci),
                                     _p(tva))
    assert st == 0, st
    return trp, tci[:len(va)], tva[:len(va)]


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


# ---- distributed set-up (reference Partition / separate_local_nonlocal / index_map) ----
def partition(kind, mapping_or_ids=None, ranges=None, num_parts=0, global_size=0):
    """kind 0: build_from_mapping(mapping, num_parts); 1: build_from_contiguous(ranges[, ids]);
    2: build_from_global_size_uniform(num_parts, global_size).  Returns a dict."""
    ids = None if mapping_or_ids is None else np.ascontiguousarray(mapping_or_ids, np.int32)
    rg = None if ranges is None else np.ascontiguousarray(ranges, np.int64)
    if kind == 0:
        n = len(ids)
        cap_r, cap_p = n + 1, num_parts
    elif kind == 1:
        n = len(rg) - 1
        cap_r, cap_p = n + 1, n
    else:
        n = 0
        cap_r, cap_p = num_parts + 1, num_parts
    meta = np.zeros(6, np.int64)
    bounds = np.zeros(cap_r + 1, np.int64)
    pid = np.zeros(max(cap_r, 1), np.int32)
    start = np.zeros(max(cap_r, 1), np.int32)
    sizes = np.zeros(max(cap_p, 1), np.int32)
    st = lib().refshim_partition(kind, n, _p(ids), _p(rg), num_parts, global_size, _p(meta), _p(bounds),
                                 _p(pid), _p(start), _p(sizes))
    assert st == 0, st
    nr, npart = int(meta[1]), int(meta[2])
    return dict(size=int(meta[0]), num_ranges=nr, num_parts=npart, num_empty_parts=int(meta[3]),
                connected=bool(meta[4]), ordered=bool(meta[5]), range_bounds=bounds[:nr + 1].copy(),
                part_ids=pid[:nr].copy(), starting_indices=start[:nr].copy(), part_sizes=sizes[:npart].copy())


def separate_local_nonlocal(shape, rows, cols, vals, row_mapping, col_mapping, num_parts, local_part):
    rows = np.ascontiguousarray(rows, np.int64)
    cols = np.ascontiguousarray(cols, np.int64)
    vals = np.ascontiguousarray(vals, np.float64)
    rm = np.ascontiguousarray(row_mapping, np.int32)
    cm = np.ascontiguousarray(col_mapping, np.int32)
    nnz = len(rows)
    cap = max(nnz, 1)
    counts = np.zeros(2, np.int64)
    lr, lc, nlr = (np.zeros(cap, np.int32) for _ in range(3))
    nlc = np.zeros(cap, np.int64)
    lv, nlv = np.zeros(cap), np.zeros(cap)
    st = lib().refshim_separate_local_nonlocal(shape[0], shape[1], nnz, _p(rows), _p(cols), _p(vals), _p(rm),
                                               _p(cm), num_parts, local_part, _p(counts), _p(lr), _p(lc),
                                               _p(lv), _p(nlr), _p(nlc), _p(nlv))
    assert st == 0, st
    a, b = int(counts[0]), int(counts[1])
    return (lr[:a], lc[:a], lv[:a]), (nlr[:b], nlc[:b], nlv[:b])


def index_map(mapping, num_parts, rank, conns, index_space=0, query=None):
    mp = np.ascontiguousarray(mapping, np.int32)
    cn = np.ascontiguousarray(conns, np.int64)
    m = len(cn)
    meta = np.zeros(2, np.int64)
    rg = np.zeros(max(m, 1), np.int64)
    rl = np.zeros(max(m, 1), np.int32)
    ids = np.zeros(max(num_parts, 1), np.int32)
    sizes = np.zeros(max(num_parts, 1), np.int64)
    q = None if query is None else np.ascontiguousarray(query, np.int64)
    k = 0 if q is None else len(q)
    ql = np.zeros(max(k, 1), np.int32)
    st = lib().refshim_index_map(_p(mp), len(mp), num_parts, rank, m, _p(cn), _p(meta), _p(rg), _p(rl),
                                 _p(ids), _p(sizes), index_space, k, _p(q), _p(ql))
    assert st == 0, st
    a, b = int(meta[0]), int(meta[1])
    return dict(remote_global=rg[:a].copy(), remote_local=rl[:a].copy(), target_ids=ids[:b].copy(),
                remote_sizes=sizes[:b].copy(), query_local=ql[:k].copy())


def vector_build_local(shape, rows, cols, vals, mapping, num_parts, local_part):
    """distributed_vector::build_local into a zeroed n_local x ncols block"""
    rows = np.ascontiguousarray(rows, np.int64)
    cols = np.ascontiguousarray(cols, np.int64)
    vals = np.ascontiguousarray(vals, np.float64)
    mp = np.ascontiguousarray(mapping, np.int32)
    n_local = int((mp == local_part).sum())
    local = np.zeros((max(n_local, 1), shape[1]))
    st = lib().refshim_vector_build_local(shape[0], shape[1], len(rows), _p(rows), _p(cols), _p(vals), _p(mp),
                                          num_parts, local_part, n_local, _p(local))
    assert st == 0, st
    return local[:n_local]
