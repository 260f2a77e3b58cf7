"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/liboracle.so (the C
restatement of the reference executor).  Imported by tests/, smoke() and
bench.py's cpu_baseline leg -- never by the ginkgo_b200 package."""
import ctypes
import os
import subprocess

from ginkgo_b200 import _cdecl

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

_lib = None


class SolverCfg(ctypes.Structure):
    _fields_ = [
        ("precond", ctypes.c_int32), ("num_blocks", ctypes.c_int64),
        ("block_offset", ctypes.c_int64), ("group_offset", ctypes.c_int64),
        ("group_power", ctypes.c_int32), ("block_ptrs", ctypes.c_void_p),
        ("blocks", ctypes.c_void_p), ("max_iters", ctypes.c_int64),
        ("res_kind", ctypes.c_int32), ("baseline", ctypes.c_int32),
        ("reduction_factor", ctypes.c_double), ("iter_first", ctypes.c_int32),
        ("krylov_dim", ctypes.c_int32), ("ortho", ctypes.c_int32),
        ("relaxation_factor", ctypes.c_double), ("foci_lo", ctypes.c_double),
        ("foci_hi", ctypes.c_double), ("initial_guess", ctypes.c_int32),
    ]


def build():
    subprocess.run(["make", "-C", _HERE, "CC=gcc"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = ctypes.CDLL(LIB_PATH)
        text = _cdecl.preprocess(os.path.join(_HERE, "oracle.c"))
        decls = _cdecl.parse(text, "orc_")
        _cdecl.bind(l, decls)
        _lib = l
    return _lib


def call(name, *args):
    return getattr(lib(), name)(*[_cdecl.as_arg(a) for a in args])
