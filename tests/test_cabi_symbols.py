"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/ginkgo_b200.h declares; nothing here computes."""
import ctypes
import os
import subprocess

from ginkgo_b200 import _cdecl, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_is_plain_c():
    # the boundary must be consumable from C (cgo / JNI / ctypes style FFI)
    src = '#include "ginkgo_b200.h"\nint main(void){return 0;}\n'
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-x", "c", "-"], input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr


def test_library_exports_every_declared_symbol():
    decls = _lib.declarations()
    assert len(decls) > 100
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in decls if not hasattr(lib, n)]
    assert not missing, missing


def test_no_undeclared_exports():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T b200_" in l}
    assert exported == set(_lib.declarations())


def test_library_is_sm100a_cuda():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    assert "sm_100a" in out.stdout
    assert b"sm_100a" in _lib.lib().b200_version()


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    ctx = ctypes.c_void_p()
    st = _lib.lib().b200_ctx_create(0, None, ctypes.byref(ctx))
    assert st != 0 and b"no CPU fallback" in _lib.lib().b200_last_error()


def test_product_does_not_touch_oracle():
    # the oracle is test infrastructure: nothing under ginkgo_b200/ may reference it
    for dp, _, fs in os.walk(os.path.join(ROOT, "ginkgo_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle" not in txt.lower().replace("test oracle", ""), os.path.join(dp, f)


def test_host_library_exports_every_handle_the_python_face_binds():
    # ginkgo_b200/lib/libgko_b200_host.so (the C++ host layer's C handles): api._configure_host_lib
    # and DistMatrix._bind set argtypes on every gkob_* they use, which fails on a missing export
    from ginkgo_b200 import api
    h = api._host()
    api.DistMatrix._bind(h)
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(os.path.dirname(_lib.LIB_PATH),
                                                                     "libgko_b200_host.so")],
                         capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T gkob_" in l}
    assert len(exported) > 50
    import re
    used = set(re.findall(r"\bgkob_\w+", open(os.path.join(ROOT, "ginkgo_b200", "api.py")).read()))
    used = {u for u in used if not u.endswith("_")}  # name prefixes completed with a value type
    assert used <= exported, sorted(used - exported)
