#!/usr/bin/env python3
"""Generates tests/golden/jacobi_adaptive_reference.json: what the REAL reference
(gko::preconditioner::Jacobi with storage_optimization, through oracle/_ref + oracle/ref_shim.cpp)
produces for the deterministic inputs of tests/jacobi_cases.py -- the storage scheme, the
precision_reduction byte and the condition number of every block, the stored bytes of every block
(hex, only the positions generate writes), and the result of apply / advanced apply on a fixed
right-hand side.  Run here (CPU only, seconds); the oracle (tests/test_jacobi_adaptive_cpu.py) and the
device (tests/test_jacobi_adaptive_gpu.py) are compared with the committed file, so this parity check
does not need /root/reference at test time."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests import jacobi_cases as JC  # noqa: E402

CASES = [  # (value type, n, max block size, seed, requested storage, accuracy)
    ("f64", 96, 4, 11, JC.AUTODETECT, 0.1), ("f64", 96, 13, 12, JC.AUTODETECT, 0.5),
    ("f64", 96, 16, 13, JC.AUTODETECT, 1e-3), ("f64", 96, 32, 14, "mixed", 0.1),
    ("f64", 96, 7, 15, 0x11, 0.1), ("f32", 96, 16, 16, JC.AUTODETECT, 0.1),
    ("f32", 96, 8, 17, "mixed", 0.5), ("f32", 96, 32, 18, 0x10, 0.1),
]


def main():
    out = {"generator": "scripts/gen_jacobi_adaptive_golden.py",
           "reference": "ginkgo v1.12.0 @ 591cd136 (oracle/_ref: ReferenceExecutor)", "cases": []}
    for vt, n, max_bs, seed, storage, acc in CASES:
        dt = np.float64 if vt == "f64" else np.float32
        rp, ci, va, ptrs = JC.make(n, max_bs, seed, dt)
        nb = len(ptrs) - 1
        so = storage if storage != "mixed" else JC.storage_request("mixed", nb, seed)
        b = np.random.default_rng(seed + 100).uniform(-1, 1, (n, 2)).astype(dt)
        x0 = np.random.default_rng(seed + 200).uniform(-1, 1, (n, 2)).astype(dt)
        R = ref.jacobi_adaptive(rp, ci, va, max_bs, ptrs, so, acc, b=b)
        R2 = ref.jacobi_adaptive(rp, ci, va, max_bs, ptrs, so, acc, b=b, x=x0, alpha=-0.75, beta=1.5)
        bo, go, gp, space = JC.scheme(max_bs, nb)
        assert (bo, go, gp) == (R["block_offset"], R["group_offset"], R["group_power"])
        mask = JC.written_mask(ptrs, R["precisions"], vt == "f64", bo, go, gp, space, dt().itemsize)
        stored = R["blocks"].view(np.uint8)[:len(mask)][mask]
        out["cases"].append({
            "vt": vt, "n": n, "max_block_size": max_bs, "seed": seed,
            "storage": storage if isinstance(storage, str) else int(storage), "accuracy": acc,
            "num_blocks": nb, "precisions": [int(p) for p in R["precisions"]],
            "conditioning_hex": R["conditioning"].tobytes().hex(),
            "stored_bytes_hex": stored.tobytes().hex(),
            "x_hex": R["x"].tobytes().hex(), "x_advanced_hex": R2["x"].tobytes().hex()})
    path = os.path.join(ROOT, "tests", "golden", "jacobi_adaptive_reference.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
