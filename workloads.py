"""Synthetic workloads of BASELINE.json (cfg1..cfg5), generated identically on the
GPU (torch, for the device arm) and on the host (numpy, for the oracle / the
reference OMP arm) from a counter-based hash -- no RNG state, so any row range
can be produced independently (row-sharded multi-GPU runs, bounded CPU samples).

Harness code (bench.py, tests); not part of the product path.

  cfg1  5-pt 2-D Laplacian 316x316         n=99 856     nnz=498 016   (diag 4, off -1)
  cfg2  random CSR, exactly 15 distinct uniform columns per row, values U(-1,1)
                                           n=10 000 000 nnz=150 000 000
        (+ banded twin: columns row + {-7..7} clipped to the matrix, same nnz/row inside)
  cfg3  7-pt 3-D Laplacian 200^3           n=8 000 000  nnz=55 760 000  (diag 6, off -1)
  cfg4  random nonsymmetric, 20 nnz/row incl. a dominant diagonal, fp32
                                           n=4 000 000  nnz=80 000 000
  cfg5  7-pt 3-D Laplacian 400^3           n=64 000 000 nnz=447 040 000
"""
import numpy as np

SEED = 42
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_G = 0x9E3779B97F4A7C15


# ------------------------------------------------------------------ hash (splitmix64 finaliser)
def _mix_np(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
    return z ^ (z >> np.uint64(31))


def hash_np(stream, a, b):
    """uint64 hash of (SEED, stream, a, b); a, b integer arrays (broadcast)"""
    with np.errstate(over="ignore"):
        a = np.asarray(a).astype(np.uint64)
        b = np.asarray(b).astype(np.uint64)
        z = np.uint64(SEED) * np.uint64(_G) + np.uint64(stream) * np.uint64(_M2)
        z = _mix_np(z + a * np.uint64(_G))
        z = _mix_np(z + b * np.uint64(_M1) + np.uint64(_G))
    return z


def _s64(c):
    """python int -> signed 64-bit two's complement value"""
    c &= (1 << 64) - 1
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr_t(z, s):
    return (z >> s) & ((1 << (64 - s)) - 1)


def _mix_t(z):
    z = (z ^ _lsr_t(z, 30)) * _s64(_M1)
    z = (z ^ _lsr_t(z, 27)) * _s64(_M2)
    return z ^ _lsr_t(z, 31)


def hash_t(stream, a, b):
    """torch int64 flavour of hash_np (bit-identical, two's complement wrap-around)"""
    z0 = _s64(SEED * _G + stream * _M2)
    z = _mix_t(a * _s64(_G) + z0)
    z = _mix_t(z + b * _s64(_M1) + _s64(_G))
    return z


def _unit_np(h, dtype):
    # top 53 bits -> [0,1) -> (-1,1)
    return ((h >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0).astype(dtype)


def _unit_t(h, dtype):
    return (_lsr_t(h, 11).double() * (2.0 / 9007199254740992.0) - 1.0).to(dtype)


def _mod_np(h, m):
    return (h % np.uint64(m)).astype(np.int64)


def _mod_t(h, m):
    # unsigned modulo of the 64-bit pattern, using only signed ops:  h = 2*hi + lo
    hi = _lsr_t(h, 1)
    lo = h & 1
    return ((hi % m) * 2 + lo) % m


# -------------------------------------------------------------------------- vectors
def vector(n, stream=7, r0=0, r1=None, xp="np", dtype=None, device=None):
    """x_i = U(-1,1) from the hash; rows [r0, r1)"""
    r1 = n if r1 is None else r1
    if xp == "np":
        i = np.arange(r0, r1, dtype=np.int64)
        return _unit_np(hash_np(stream, i, 0), dtype or np.float64)
    import torch
    i = torch.arange(r0, r1, dtype=torch.int64, device=device)
    return _unit_t(hash_t(stream, i, torch.zeros((), dtype=torch.int64, device=device)),
                   dtype or torch.float64)


# -------------------------------------------------------------------- random CSR (cfg2/4)
def random_csr(n, per_row, r0=0, r1=None, xp="np", vdtype=None, device=None, diag_dominant=False,
               stream=1):
    """Rows [r0, r1) of an n x n CSR with exactly `per_row` distinct columns per row.

    Columns: `per_row` iid uniform draws from [0, n - per_row], sorted, plus their rank -- a
    bijection from multisets to strictly increasing tuples, so the columns are distinct,
    sorted and (jointly) uniform.  Values U(-1,1).  diag_dominant (cfg4): the entry closest to
    the diagonal is moved onto it and set to sum|a_ij| + 1 (strict row dominance).
    Returns (row_ptrs[int32, local], col_idxs[int32, global], values)."""
    r1 = n if r1 is None else r1
    m = r1 - r0
    span = n - per_row + 1
    if xp == "np":
        rows = np.arange(r0, r1, dtype=np.int64)[:, None]
        ks = np.arange(per_row, dtype=np.int64)[None, :]
        cols = np.sort(_mod_np(hash_np(stream, rows, ks), span), axis=1) + ks
        vals = _unit_np(hash_np(stream + 100, rows, ks), vdtype or np.float64)
        if diag_dominant:
            j = np.abs(cols - rows).argmin(axis=1)
            cols[np.arange(m), j] = rows[:, 0]
            cols = np.sort(cols, axis=1)  # stays distinct only if no clash; fix below
            vals = _fix_dominant_np(cols, vals, rows)
        rp = (np.arange(m + 1, dtype=np.int64) * per_row).astype(np.int32)
        return rp, cols.reshape(-1).astype(np.int32), vals.reshape(-1)
    import torch
    rows = torch.arange(r0, r1, dtype=torch.int64, device=device)[:, None]
    ks = torch.arange(per_row, dtype=torch.int64, device=device)[None, :]
    cols = torch.sort(_mod_t(hash_t(stream, rows, ks), span), dim=1).values + ks
    vals = _unit_t(hash_t(stream + 100, rows, ks), vdtype or torch.float64)
    if diag_dominant:
        j = (cols - rows).abs().argmin(dim=1)
        cols[torch.arange(m, device=device), j] = rows[:, 0]
        cols = torch.sort(cols, dim=1).values
        vals = _fix_dominant_t(cols, vals, rows)
    rp = (torch.arange(m + 1, dtype=torch.int64, device=device) * per_row).to(torch.int32)
    return rp, cols.reshape(-1).to(torch.int32), vals.reshape(-1)


def _fix_dominant_np(cols, vals, rows):
    isd = cols == rows
    off = np.where(isd, 0, np.abs(vals.astype(np.float64)))
    d = off.sum(axis=1, keepdims=True) + 1.0
    return np.where(isd, d, vals).astype(vals.dtype)


def _fix_dominant_t(cols, vals, rows):
    import torch
    isd = cols == rows
    off = torch.where(isd, torch.zeros((), dtype=torch.float64, device=vals.device),
                      vals.double().abs())
    d = off.sum(dim=1, keepdim=True) + 1.0
    return torch.where(isd, d, vals.double()).to(vals.dtype)


def banded_csr(n, half_bw, r0=0, r1=None, xp="np", vdtype=None, device=None, stream=3):
    """banded twin of cfg2: columns row + {-half_bw..half_bw} clipped to [0, n)"""
    r1 = n if r1 is None else r1
    w = 2 * half_bw + 1
    if xp == "np":
        rows = np.arange(r0, r1, dtype=np.int64)[:, None]
        offs = np.arange(-half_bw, half_bw + 1, dtype=np.int64)[None, :]
        cols = rows + offs
        ok = (cols >= 0) & (cols < n)
        vals = _unit_np(hash_np(stream, rows, offs + half_bw), vdtype or np.float64)
        rp = np.zeros(r1 - r0 + 1, dtype=np.int64)
        rp[1:] = np.cumsum(ok.sum(axis=1))
        return rp.astype(np.int32), cols[ok].astype(np.int32), vals[ok]
    import torch
    rows = torch.arange(r0, r1, dtype=torch.int64, device=device)[:, None]
    offs = torch.arange(-half_bw, half_bw + 1, dtype=torch.int64, device=device)[None, :]
    cols = rows + offs
    ok = (cols >= 0) & (cols < n)
    vals = _unit_t(hash_t(stream, rows.expand(-1, w), offs + half_bw), vdtype or torch.float64)
    rp = torch.zeros(r1 - r0 + 1, dtype=torch.int64, device=device)
    rp[1:] = torch.cumsum(ok.sum(dim=1), 0)
    return rp.to(torch.int32), cols[ok].to(torch.int32), vals[ok]


# ---------------------------------------------------------------- skewed rows (power law)
def zipf_lengths(n, nnz_target, xp="np", device=None):
    """row lengths ~ c / rank (Zipf, exponent 1), ranks scattered over the rows by the bijection
    rank(r) = (r * A + B) mod n; c is chosen so that the total is ~ nnz_target.  For n = 10 M and
    nnz = 150 M the longest row has ~9 M entries, ~550 rows have >= 16384, ~1 M rows have one."""
    import math
    c = nnz_target / (math.log(n) + 0.5772156649)
    A = 2_654_435_761 % n
    while math.gcd(A, n) != 1:
        A += 1
    if xp == "np":
        r = np.arange(n, dtype=np.int64)
        rank = (r * A + 12345) % n
        return np.clip(np.rint(c / (rank + 1.0)), 1, n).astype(np.int64)
    import torch
    r = torch.arange(n, dtype=torch.int64, device=device)
    rank = (r * A + 12345) % n
    return torch.clamp(torch.round(c / (rank + 1.0).double()), 1, n).to(torch.int64)


def zipf_csr(n, nnz_target, xp="np", vdtype=None, device=None, stream=5):
    """n x n CSR with Zipf row lengths; the L columns of a row are one uniformly jittered pick out of
    each of L equal bins of [0, n) (sorted, distinct, spread over all of x).  Values U(-1,1).
    Returns (row_ptrs int32/int64, col_idxs int32, values)."""
    lens = zipf_lengths(n, nnz_target, xp, device)
    if xp == "np":
        rp = np.zeros(n + 1, dtype=np.int64)
        rp[1:] = np.cumsum(lens)
        nnz = int(rp[-1])
        row = np.repeat(np.arange(n, dtype=np.int64), lens)
        k = np.arange(nnz, dtype=np.int64) - rp[row]
        L = lens[row]
        u = (hash_np(stream, row, k) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
        lo, hi = k * n // L, (k + 1) * n // L  # integer bins of [0, n): never empty (L <= n), disjoint
        cols = lo + np.minimum((u * (hi - lo)).astype(np.int64), hi - lo - 1)
        vals = _unit_np(hash_np(stream + 100, row, k), vdtype or np.float64)
        return rp.astype(np.int32 if nnz < 2**31 else np.int64), cols.astype(np.int32), vals
    import torch
    rp = torch.zeros(n + 1, dtype=torch.int64, device=device)
    rp[1:] = torch.cumsum(lens, 0)
    nnz = int(rp[-1].item())
    row = torch.repeat_interleave(torch.arange(n, dtype=torch.int64, device=device), lens)
    k = torch.arange(nnz, dtype=torch.int64, device=device) - rp[row]
    L = lens[row]
    u = _lsr_t(hash_t(stream, row, k), 11).double() / 9007199254740992.0
    lo, hi = k * n // L, (k + 1) * n // L
    cols = lo + torch.minimum((u * (hi - lo).double()).to(torch.int64), hi - lo - 1)
    vals = _unit_t(hash_t(stream + 100, row, k), vdtype or torch.float64)
    return rp.to(torch.int32 if nnz < 2**31 else torch.int64), cols.to(torch.int32), vals


# ------------------------------------------------------------------------ stencils
def laplace(grid, dims, r0=0, r1=None, xp="np", vdtype=None, device=None):
    """5-pt (dims=2) / 7-pt (dims=3) Laplacian on a grid^dims box, natural ordering, diag =
    2*dims, off-diag -1 (semantics of benchmark/utils/stencil_matrix.hpp:195-240,408-465).
    Rows [r0, r1) -> (row_ptrs local int32, col_idxs global int32, values)."""
    n = grid ** dims
    r1 = n if r1 is None else r1
    strides = [grid ** (dims - 1 - d) for d in range(dims)]  # x slowest ... last fastest
    offs = sorted([-s for s in strides] + [0] + strides)
    if xp == "np":
        rows = np.arange(r0, r1, dtype=np.int64)[:, None]
        o = np.array(offs, dtype=np.int64)[None, :]
        cols = rows + o
        ok = np.ones(cols.shape, dtype=bool)
        for s in strides:
            c = (rows // s) % grid
            ok &= ~((o == -s) & (c == 0)) & ~((o == s) & (c == grid - 1))
        vals = np.where(o == 0, 2.0 * dims, -1.0) * np.ones(cols.shape)
        rp = np.zeros(r1 - r0 + 1, dtype=np.int64)
        rp[1:] = np.cumsum(ok.sum(axis=1))
        return rp.astype(np.int32), cols[ok].astype(np.int32), vals[ok].astype(vdtype or np.float64)
    import torch
    rows = torch.arange(r0, r1, dtype=torch.int64, device=device)[:, None]
    o = torch.tensor(offs, dtype=torch.int64, device=device)[None, :]
    cols = rows + o
    ok = torch.ones(cols.shape, dtype=torch.bool, device=device)
    for s in strides:
        c = (rows // s) % grid
        ok &= ~((o == -s) & (c == 0)) & ~((o == s) & (c == grid - 1))
    vals = torch.where(o == 0, 2.0 * dims, -1.0).to(vdtype or torch.float64).expand(cols.shape)
    rp = torch.zeros(r1 - r0 + 1, dtype=torch.int64, device=device)
    rp[1:] = torch.cumsum(ok.sum(dim=1), 0)
    return rp.to(torch.int32), cols[ok].to(torch.int32), vals[ok].contiguous()


CONFIGS = {
    "cfg1": dict(kind="laplace", grid=316, dims=2, n=99856, nnz=498016, dtype="f64"),
    "cfg2": dict(kind="random", n=10_000_000, per_row=15, nnz=150_000_000, dtype="f64"),
    "cfg2_banded": dict(kind="banded", n=10_000_000, half_bw=7, dtype="f64"),
    "cfg2_zipf": dict(kind="zipf", n=10_000_000, nnz=150_000_000, dtype="f64"),
    "cfg3": dict(kind="laplace", grid=200, dims=3, n=8_000_000, nnz=55_760_000, dtype="f64"),
    "cfg4": dict(kind="random", n=4_000_000, per_row=20, nnz=80_000_000, dtype="f32",
                 diag_dominant=True),
    "cfg5": dict(kind="laplace", grid=400, dims=3, n=64_000_000, nnz=447_040_000, dtype="f64"),
}


def build(name, r0=0, r1=None, xp="np", device=None, n=None):
    """rows [r0, r1) of a named config (n overrides the size for scaled-down parity cases)"""
    c = dict(CONFIGS[name])
    if n is not None:
        c["n"] = n
    vd = None
    if xp == "np":
        vd = np.float64 if c["dtype"] == "f64" else np.float32
    else:
        import torch
        vd = torch.float64 if c["dtype"] == "f64" else torch.float32
    if c["kind"] == "laplace":
        return laplace(c["grid"], c["dims"], r0, r1, xp, vd, device)
    if c["kind"] == "banded":
        return banded_csr(c["n"], c["half_bw"], r0, r1, xp, vd, device)
    if c["kind"] == "zipf":
        return zipf_csr(c["n"], c["nnz"] * c["n"] // CONFIGS["cfg2_zipf"]["n"], xp, vd, device)
    return random_csr(c["n"], c["per_row"], r0, r1, xp, vd, device, c.get("diag_dominant", False))


def spmv_bytes(n_rows, n_cols, nnz, vbytes=8, ibytes=4, beta=False):
    """algorithmic bytes of one SpMV"""
    return nnz * (vbytes + ibytes) + (n_rows + 1) * ibytes + n_cols * vbytes + \
        n_rows * vbytes * (2 if beta else 1)
