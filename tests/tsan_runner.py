"""TEST INFRASTRUCTURE: run by tests/test_dist_assembly_cpu.py in a subprocess with libtsan
preloaded.  Drives host-compiled, ThreadSanitizer-instrumented copies of dist_assembly.cu and
bicg_transpose.cu whose launch_ew runs every "kernel" on 4 real threads (-DB200_SHIM_THREADS=4):
any write/write or read/write conflict inside one launch is reported by TSan."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from tests import dist_driver as D  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_dist_assembly_cpu import random_mapping  # noqa: E402
from tests.test_kernel_sources_cpu import KernelSourceBackend  # noqa: E402
from tests.test_transpose_bicg_cpu import random_csr, transpose  # noqa: E402

da = KernelSourceBackend(ctypes.CDLL(sys.argv[1]))
bt = KernelSourceBackend(ctypes.CDLL(sys.argv[2]))
orc = H.Oracle()
for seed in range(6):
    rng = np.random.default_rng(seed)
    num_parts, n = int(rng.integers(1, 7)), int(rng.integers(1, 300))
    for lt, gt in (("i32", "i32"), ("i32", "i64"), ("i64", "i64")):
        rp = D.partition_from_mapping(da, random_mapping(rng, n, num_parts, 9), num_parts, lt, gt)
        cp = D.partition_uniform(da, num_parts, n, lt, gt)
        nnz = int(rng.integers(0, 2000))
        order = np.unique(rng.integers(0, n * n, nnz)) if nnz else np.zeros(0, np.int64)
        rows, cols = order // n, order % n
        vals = rng.standard_normal(len(order))
        for part in range(num_parts):
            s = D.separate(da, rp, cp, rows, cols, vals, part)
            im = D.IndexMap(da, cp, part, s["kept"][1], skip_part=part)
            for sp in (0, 1, 2):
                im.map_to_local(da, rng.integers(-3, n + 3, 40), sp)
            # unique (row, column) pairs, the documented precondition of build_local
            vo = np.unique(rng.integers(0, n * 3, 200))
            D.vector_build_local(da, rp, vo // 3, vo % 3, rng.standard_normal(len(vo)), 3, part)
for n, m, mr in [(900, 300, 12), (500, 70000, 30), (3000, 256, 5), (40, 1, 3), (6, 6, 0)]:
    for it in ("i32", "i64"):
        rng = np.random.default_rng(n + m)
        rp, ci, va = random_csr(rng, n, m, mr, "f64", it)
        a, b = transpose(orc, rp, ci, va, m, "f64", it), transpose(bt, rp, ci, va, m, "f64", it)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
rng = np.random.default_rng(1)
nb, mbs = 50, 8
ptrs = np.concatenate([[0], np.cumsum(rng.integers(1, mbs + 1, nb))]).astype(np.int32)
space = mbs * 4 * mbs * ((nb + 3) // 4)
bt("jacobi_transpose_f64_i32", nb, mbs, mbs, mbs * 4 * mbs, 2, ptrs, rng.standard_normal(space), np.zeros(space))
v = [H.dense(rng, 300, 5, 6, "f64") for _ in range(9)]
sc = [rng.uniform(0.5, 1, 5) for _ in range(3)]
stop = np.zeros(5, np.uint8)
bt("bicg_step_1_f64", 300, 5, v[0], 6, v[1], 6, v[2], 6, v[3], 6, sc[0], sc[1], stop)
bt("bicg_step_2_f64", 300, 5, v[4], 6, v[5], 6, v[6], 6, v[0], 6, v[7], 6, v[8], 6, sc[2], sc[0], stop)
bt("dense_compute_sqrt_f64", 3, 5, np.abs(v[0][:3]).copy(), 6)
print("TSAN_RUNNER_DONE")
