"""Tiny C-declaration reader: turns preprocessed `ret name(type a, type b, ...)`
prototypes into ctypes signatures, so the Python harness binds exactly what
include/ginkgo_b200.h declares (no hand-maintained argtypes)."""
import ctypes
import re
import subprocess

_SCALARS = {
    "int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "uint8_t": ctypes.c_uint8,
    "uint64_t": ctypes.c_uint64, "size_t": ctypes.c_size_t, "double": ctypes.c_double,
    "float": ctypes.c_float, "int": ctypes.c_int, "b200_status": ctypes.c_int32,
}


def _ctype(t):
    t = t.strip()
    if "*" in t:
        return ctypes.c_char_p if "char" in t else ctypes.c_void_p
    t = t.replace("const", "").strip()
    if t == "void":
        return None
    return _SCALARS[t]


def preprocess(path, include_dirs=()):
    cmd = ["gcc", "-E", "-P"] + ["-I" + d for d in include_dirs] + [path]
    return subprocess.run(cmd, check=True, capture_output=True, text=True).stdout


_DECL = re.compile(
    r"(?:^|(?<=[;}]))\s*((?:const\s+)?(?:void|char|int64_t|int32_t|b200_status|int|double)\s*\**)\s*(\w+)\s*\(([^()]*)\)\s*(;|\{)",
    re.M)


def parse(text, prefix):
    """-> {name: (restype, [argtypes], [argnames])} for functions starting with prefix."""
    out = {}
    for m in _DECL.finditer(text):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        if not name.startswith(prefix) or "typedef" in ret or "static" in ret:
            continue
        ret = ret.replace("extern", "").strip()
        if not ret:
            continue
        argtypes, argnames = [], []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                argtypes.append(_ctype(mm.group(1)))
                argnames.append(mm.group(2))
        try:
            out[name] = (_ctype(ret), argtypes, argnames)
        except KeyError:
            continue
    return out


def bind(lib, decls):
    missing = []
    for name, (ret, argtypes, _) in decls.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = ret
        fn.argtypes = argtypes
    return missing


def as_arg(x):
    """torch tensor / numpy array / None / scalar -> ctypes-friendly value."""
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes") and hasattr(x, "dtype"):
        return x.ctypes.data
    return x
