"""The drop-in boundary, compiled (VERDICT r01 Missing #1): tests/dropin/Makefile links the REFERENCE's own
core / Reference / OMP objects (oracle/_ref/obj, compiled in place from /root/reference) with
ginkgo_b200/dropin/cuda_backend.cpp in place of the reference's cuda stub
(/root/reference/core/device_hooks/cuda_hooks.cpp) -- a complete Ginkgo whose gko::CudaExecutor runs on
libginkgo_b200.so.  tests/dropin/dropin_check.cpp then runs the reference's own host code
(gko::matrix::{Csr,Ell,Sellp,Coo,Hybrid}::apply, gko::solver::{Cg,Bicgstab,Gmres} + gko::preconditioner::Jacobi +
gko::stop::*) on that executor and compares with gko::ReferenceExecutor.
 * not gpu: the library exists, the hot gko::kernels::cuda::* symbols resolve to the B200 wrappers (they call
   b200_* entry points), everything else still resolves to the reference's NotCompiled stub, and the same
   flow passes on the OmpExecutor (so the check program itself is sound);
 * gpu: the flow on CudaExecutor."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "dropin", "_build")
LIB = os.path.join(BUILD, "libginkgo_b200_dropin.so")
CHECK = os.path.join(BUILD, "dropin_check")

needs_build = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CHECK)),
                                 reason="tests/dropin/_build missing (needs /root/reference at build time)")

HOT = [
    "gko::kernels::cuda::csr::spmv<double, double, double, int>",
    "gko::kernels::cuda::csr::advanced_spmv<double, double, double, int>",
    "gko::kernels::cuda::ell::spmv<double, double, double, int>",
    "gko::kernels::cuda::sellp::spmv<double, int>",
    "gko::kernels::cuda::coo::spmv2<double, int>",
    "gko::kernels::cuda::dense::compute_norm2<double>",
    "gko::kernels::cuda::dense::compute_conj_dot<double>",
    "gko::kernels::cuda::dense::add_scaled<double, double>",
    "gko::kernels::cuda::cg::step_1<double>",
    "gko::kernels::cuda::cg::step_2<float>",
    "gko::kernels::cuda::bicgstab::step_3<double>",
    "gko::kernels::cuda::common_gmres::hessenberg_qr<float>",
    "gko::kernels::cuda::gmres::multi_dot<float>",
    "gko::kernels::cuda::residual_norm::residual_norm<double>",
    "gko::kernels::cuda::jacobi::simple_scalar_apply<double>",
    "gko::kernels::cuda::jacobi::simple_apply<float, int>",
    "gko::kernels::cuda::jacobi::generate<double, int>",
    "gko::kernels::cuda::jacobi::apply<double, int>",
    "gko::kernels::cuda::jacobi::transpose_jacobi<double, int>",
    "gko::kernels::cuda::jacobi::conj_transpose_jacobi<float, long>",
    "gko::kernels::cuda::jacobi::initialize_precisions",
]


def _disasm_calls(symbol_prefix):
    """names called by the function(s) whose demangled name starts with symbol_prefix"""
    out = subprocess.run(["objdump", "-d", "-C", "--no-show-raw-insn", LIB], capture_output=True, text=True,
                         check=True).stdout
    calls, inside = set(), False
    for line in out.splitlines():
        if line.endswith(">:"):
            inside = ("<void " + symbol_prefix) in line or ("<" + symbol_prefix) in line
        elif inside and "call" in line and "<" in line:
            calls.add(line.split("<", 1)[1].rsplit(">", 1)[0])
    return calls


@needs_build
def test_hot_kernels_resolve_to_the_b200_backend():
    out = subprocess.run(["objdump", "-d", "-C", "--no-show-raw-insn", LIB], capture_output=True, text=True,
                         check=True).stdout
    # split into functions once
    funcs, name = {}, None
    for line in out.splitlines():
        if line.endswith(">:"):
            name = line.split("<", 1)[1][:-2]
            funcs[name] = []
        elif name is not None and "call" in line:
            funcs[name].append(line)
    for sym in HOT:
        bodies = [v for k, v in funcs.items() if sym + "(" in k]
        assert bodies, "symbol not defined: " + sym
        assert any(any("b200_" in c for c in b) for b in bodies), sym + " does not call the B200 C ABI"
    # a kernel outside the path still resolves to the reference's own stub (throws NotCompiled)
    stub = [v for k, v in funcs.items() if "gko::kernels::cuda::csr::spgemm<double, int>(" in k]
    assert stub and not any("b200_" in c for b in stub for c in b)


@needs_build
def test_reference_flow_passes_on_the_omp_executor():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([CHECK, "omp"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DROPIN CHECK PASSED" in r.stdout


@needs_build
@pytest.mark.gpu
def test_reference_solvers_run_on_cuda_executor_through_b200_kernels():
    """gko::solver::Cg::build()...on(CudaExecutor)->generate(A)->apply(b, x) of the REFERENCE, on the B200
    kernels, equals the same on gko::ReferenceExecutor (bit-identical SpMV in all five formats, same
    iteration counts for CG / BiCGStab / GMRES with scalar and block Jacobi)."""
    r = subprocess.run([CHECK, "cuda"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "DROPIN CHECK PASSED" in r.stdout
