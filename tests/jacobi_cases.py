"""Shared inputs of the adaptive block-Jacobi tests (CPU and GPU): a block-diagonal-dominated CSR
matrix whose diagonal blocks have conditioning spread over six decades, so that autodetect() picks
every storage type, plus the storage scheme of include/ginkgo/core/preconditioner/jacobi.hpp:589-625."""
import numpy as np

AUTODETECT = 0xFF
ALL_REDUCTIONS = (0x00, 0x01, 0x02, 0x10, 0x11, 0x20)


def scheme(max_bs, nb):
    pow2 = 1
    while pow2 < max_bs:
        pow2 *= 2
    group_size = 32 // pow2
    gp = group_size.bit_length() - 1
    block_offset, group_offset = max_bs, max_bs * group_size * max_bs
    space = (nb + group_size - 1) // group_size * group_offset
    return block_offset, group_offset, gp, space


def make(n, bs_max, seed, dtype, itype=np.int32, singular_block=None):
    rng = np.random.default_rng(seed)
    ptrs = [0]
    while ptrs[-1] < n:
        ptrs.append(min(n, ptrs[-1] + int(rng.integers(1, bs_max + 1))))
    ptrs = np.array(ptrs, itype)
    shifts = [3.0, 0.3, 30.0, 0.05, 1000.0, 1e-3]
    rows, cols, vals = [], [], []
    for k in range(len(ptrs) - 1):
        s, e = int(ptrs[k]), int(ptrs[k + 1])
        m = e - s
        B = rng.uniform(-1, 1, (m, m)) + np.eye(m) * shifts[k % len(shifts)]
        if singular_block == k:
            B[:] = 0
        for i in range(m):
            for j in range(m):
                if rng.random() < 0.85 or i == j:
                    rows.append(s + i)
                    cols.append(s + j)
                    vals.append(B[i, j])
            for _ in range(2):
                c = int(rng.integers(0, n))
                if not s <= c < e:
                    rows.append(s + i)
                    cols.append(c)
                    vals.append(rng.uniform(-1, 1))
    rows, cols, vals = np.array(rows), np.array(cols), np.array(vals)
    _, idx = np.unique(rows * n + cols, return_index=True)
    rows, cols, vals = rows[idx], cols[idx], vals[idx]
    rp = np.zeros(n + 1, np.int64)
    np.add.at(rp, rows + 1, 1)
    rp = np.cumsum(rp).astype(itype)
    return rp, cols.astype(itype), vals.astype(dtype), ptrs


def storage_request(kind, nb, seed=5):
    """kind: None | a byte | 'mixed' -> the per-block in/out array handed to generate (or None)"""
    if kind is None:
        return None
    if kind == "mixed":
        rng = np.random.default_rng(seed)
        return rng.choice(np.array((AUTODETECT,) + ALL_REDUCTIONS, np.uint8), size=nb)
    return np.full(nb, kind, np.uint8)


def written_mask(block_ptrs, prec, value_is_double, block_offset, group_offset, gp, space, value_bytes):
    """bytes of the block storage that generate writes (the rest is padding the reference leaves
    uninitialised)"""
    from_kind = {True: {0x01: 4, 0x02: 2, 0x10: 4, 0x11: 2, 0x20: 2},
                 False: {0x01: 2, 0x02: 2, 0x10: 2, 0x11: 2, 0x20: 2}}[value_is_double]
    mask = np.zeros(space * value_bytes, bool)
    stride = block_offset << gp
    for k in range(len(block_ptrs) - 1):
        w = from_kind.get(0 if prec is None else int(prec[k]), value_bytes)
        bs = int(block_ptrs[k + 1] - block_ptrs[k])
        base = group_offset * (k >> gp) * value_bytes
        bo = block_offset * (k & ((1 << gp) - 1))
        i = np.arange(bs)
        idx = (bo + i[:, None] + i[None, :] * stride).reshape(-1)
        for b in range(w):
            mask[base + idx * w + b] = True
    return mask


def load_golden():
    """tests/golden/jacobi_adaptive_reference.json (scripts/gen_jacobi_adaptive_golden.py): outputs of the
    REAL reference for deterministic inputs of make(); returns a list of dicts with decoded arrays"""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jacobi_adaptive_reference.json")
    cases = json.load(open(path))["cases"]
    for c in cases:
        dt = np.float64 if c["vt"] == "f64" else np.float32
        c["dtype"] = dt
        c["precisions"] = np.array(c["precisions"], np.uint8)
        c["conditioning"] = np.frombuffer(bytes.fromhex(c["conditioning_hex"]), dt)
        c["stored_bytes"] = np.frombuffer(bytes.fromhex(c["stored_bytes_hex"]), np.uint8)
        c["x"] = np.frombuffer(bytes.fromhex(c["x_hex"]), dt).reshape(c["n"], 2)
        c["x_advanced"] = np.frombuffer(bytes.fromhex(c["x_advanced_hex"]), dt).reshape(c["n"], 2)
    return cases


def check_against_golden(backend, c):
    """run generate + apply + advanced apply on `backend` (tests.helpers Oracle / Cuda) for golden case c
    and compare with the reference's committed outputs bit for bit"""
    vt, n, max_bs, dt = c["vt"], c["n"], c["max_block_size"], c["dtype"]
    rp, ci, va, ptrs = make(n, max_bs, c["seed"], dt)
    nb = len(ptrs) - 1
    assert nb == c["num_blocks"]
    storage = c["storage"]
    prec = storage_request(storage, nb, c["seed"]) if storage == "mixed" else np.full(nb, storage, np.uint8)
    cond = np.zeros(nb, dt)
    bo, go, gp, space = scheme(max_bs, nb)
    blocks = np.zeros(space, dt)
    backend("jacobi_generate_adaptive_%s_i32" % vt, n, rp, ci, va, nb, max_bs, float(c["accuracy"]), bo, go, gp, cond,
            prec, ptrs, blocks)
    assert np.array_equal(prec, c["precisions"])
    assert np.array_equal(cond.view(np.uint8), c["conditioning"].view(np.uint8))
    mask = written_mask(ptrs, prec, vt == "f64", bo, go, gp, space, dt().itemsize)
    assert np.array_equal(blocks.view(np.uint8)[mask], c["stored_bytes"])
    b = np.random.default_rng(c["seed"] + 100).uniform(-1, 1, (n, 2)).astype(dt)
    x0 = np.random.default_rng(c["seed"] + 200).uniform(-1, 1, (n, 2)).astype(dt)
    x = np.zeros((n, 2), dt)
    backend("jacobi_simple_apply_adaptive_%s_i32" % vt, nb, max_bs, bo, go, gp, prec, ptrs, blocks, b, 2, 2, x, 2)
    assert np.array_equal(x.view(np.uint8), c["x"].view(np.uint8))
    x2 = x0.copy()
    backend("jacobi_apply_adaptive_%s_i32" % vt, nb, max_bs, bo, go, gp, prec, ptrs, blocks, np.array([-0.75], dt), b, 2,
            2, np.array([1.5], dt), x2, 2)
    assert np.array_equal(x2.view(np.uint8), c["x_advanced"].view(np.uint8))
