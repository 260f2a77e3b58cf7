"""The arithmetic of the element-wise CUDA kernels, checked WITHOUT a GPU: a copy of
ginkgo_b200/csrc/krylov_steps.cu is compiled with plain g++ against tests/mock/host_cuda_shim
(the `[=] __device__` lambdas become ordinary lambdas, launch_ew becomes a host loop,
-ffp-contract=off stands for -fmad=false) and driven through the very test bodies the GPU
parity tests use, with the oracle on the other side.  What this cannot see -- launch
configuration, strides on the device -- is what the GPU tests are for."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class KernelSourceBackend:
    """calls b200_<name>(ctx, args...) of the host-compiled kernel source on host memory"""
    name = "kernel-source"

    def __init__(self, lib):
        self._lib = lib
        self._ctx = ctypes.create_string_buffer(64)

    def __call__(self, fname, *args, plan=None):
        conv, boxes = [ctypes.cast(self._ctx, ctypes.c_void_p)], []
        for a in args:
            if isinstance(a, H.OutInt):
                b = a.ctype(0)
                boxes.append((a, b))
                conv.append(ctypes.byref(b))
            elif isinstance(a, np.ndarray):
                assert a.flags.c_contiguous
                conv.append(ctypes.c_void_p(a.ctypes.data))
            elif isinstance(a, float):
                conv.append(ctypes.c_double(a))
            else:
                conv.append(ctypes.c_int64(a))
        fn = getattr(self._lib, "b200_" + fname)
        fn.restype = ctypes.c_int32
        assert fn(*conv) == 0
        for a, b in boxes:
            a.value = b.value


# execution order of the host loop that stands for the grid: ascending, descending, scrambled
@pytest.fixture(scope="module", params=[0, 1, 2], ids=["fwd", "rev", "scrambled"])
def ksrc(tmp_path_factory, request):
    d = str(tmp_path_factory.mktemp("ksrc"))
    src = open(os.path.join(ROOT, "ginkgo_b200", "csrc", "krylov_steps.cu")).read()
    # the one hand-written __global__ kernel of the file (bicgstab's status flip) and its <<<>>>
    # launch cannot be compiled by g++: give the copy the equivalent host loop
    a = src.index("template <typename V>\n__global__ void finalize_status_kernel")
    b = src.index("template <typename V>\nb200_status bicgstab_finalize")
    src = src[:a] + src[b:]
    launch = ("        finalize_status_kernel<V><<<(unsigned)ceildiv(cols, 256), 256, 0, ctx->stream>>>(cols, stop);\n"
              "        B200_LAUNCH_CHECK(ctx);\n")
    assert launch in src
    src = src.replace(launch, "        for (int64_t j = 0; j < cols; ++j)\n"
                              "            if (has_stopped(stop[j])) stop[j] |= kFinalizedMask;\n")
    open(os.path.join(d, "krylov_steps.cpp"), "w").write(src)
    shutil.copy(os.path.join(ROOT, "tests", "mock", "host_cuda_shim", "elementwise.cuh"), d)
    so = os.path.join(d, "libkrylov_steps_host.so")
    # -Bsymbolic: the copy's own template instantiations, not the same-named ones of a CUDA
    # library another test may have loaded into the process
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-ffp-contract=off", "-DB200_SHIM_ORDER=%d" % request.param,
                    "-I" + os.path.join(ROOT, "include"), "-o", so, os.path.join(d, "krylov_steps.cpp")],
                   check=True)
    return KernelSourceBackend(ctypes.CDLL(so))


SHAPES = [(597, 43), (2001, 1), (0, 2)]


@pytest.mark.parametrize("vt", ["f64", "f32"])
@pytest.mark.parametrize("rows,cols", SHAPES)
class TestKernelBodies:
    def test_cg(self, orc, ksrc, vt, rows, cols):
        from tests.test_parity_gpu import test_cg_steps
        test_cg_steps(orc, ksrc, vt, rows, cols)

    def test_bicgstab(self, orc, ksrc, vt, rows, cols):
        from tests.test_parity_gpu import test_bicgstab_steps
        test_bicgstab_steps(orc, ksrc, vt, rows, cols)

    def test_fcg(self, orc, ksrc, vt, rows, cols):
        from tests.test_parity_gpu import test_fcg_steps
        test_fcg_steps(orc, ksrc, vt, rows, cols)

    def test_cgs(self, orc, ksrc, vt, rows, cols):
        from tests.test_parity_gpu import test_cgs_steps
        test_cgs_steps(orc, ksrc, vt, rows, cols)

    def test_ir_and_chebyshev(self, orc, ksrc, vt, rows, cols):
        from tests.test_zz_late_gpu import test_ir_and_chebyshev_kernels
        test_ir_and_chebyshev_kernels(orc, ksrc, vt, rows, cols)

    def test_pipe_cg(self, orc, ksrc, vt, rows, cols):
        from tests.test_zz_late_gpu import test_pipe_cg_steps
        test_pipe_cg_steps(orc, ksrc, vt, rows, cols)

    def test_gcr(self, orc, ksrc, vt, rows, cols):
        from tests.test_zz_late_gpu import test_gcr_kernels
        test_gcr_kernels(orc, ksrc, vt, rows, cols)

    def test_minres(self, orc, ksrc, vt, rows, cols):
        from tests.test_zz_late_gpu import test_minres_kernels
        test_minres_kernels(orc, ksrc, vt, rows, cols)
