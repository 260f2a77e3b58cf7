"""Loader for the C-ABI shared library (ginkgo_b200/lib/libginkgo_b200.so).

The library is hand-written CUDA for sm_100a; there is NO fallback: if it is
missing, fails to load or no B200 is visible, importing / creating a context
raises.  Signatures are read from include/ginkgo_b200.h."""
import ctypes
import os

from . import _cdecl

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libginkgo_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "ginkgo_b200.h")
HEADERS = [HEADER]


class B200Error(RuntimeError):
    pass


_lib = None
_decls = None


def declarations():
    global _decls
    if _decls is None:
        d = {}
        for h in HEADERS:
            if os.path.exists(h):
                d.update(_cdecl.parse(_cdecl.preprocess(h), "b200_"))
        _decls = d
    return _decls


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                "%s not found: build it with `make -C ginkgo_b200/csrc` "
                "(__graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        missing = _cdecl.bind(l, declarations())
        if missing:
            raise B200Error("library lacks symbols declared in the header: %s" % missing[:5])
        _lib = l
    return _lib


def check(status):
    if status != 0:
        msg = lib().b200_last_error()
        raise B200Error("b200 status %d: %s" % (status, ctypes.string_at(msg).decode() if msg else ""))


def call(name, *args):
    """Call a status-returning entry point with tensors/arrays/scalars; raise on error."""
    fn = getattr(lib(), name)
    check(fn(*[_cdecl.as_arg(a) for a in args]))
