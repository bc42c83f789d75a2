"""ctypes binding of libsg2b200.so (the C ABI declared in include/sg2b200.h).

There is no fallback: if the library is missing or a call is rejected, a RuntimeError is raised.
"""
import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libsg2b200.so")
# tools/ only: SG2_PROBES=1 loads the diagnostics build (product sources + probes, `python -m sg2b200.build --probes`)
_PROBES = os.environ.get("SG2_PROBES", "0") == "1"
if _PROBES:
    LIB_PATH = os.path.join(_PKG, "..", "tools", "libsg2b200_probes.so")

_c_int, _c_vp, _c_float = ctypes.c_int, ctypes.c_void_p, ctypes.c_float
_c_ll = ctypes.c_longlong

HEADER_PATH = os.path.join(_PKG, "..", "include", "sg2b200.h")


def _ctype(decl):
    """Map one C parameter declaration of include/sg2b200.h to a ctypes type."""
    d = decl.strip()
    if "*" in d:
        return _c_vp
    base = " ".join(t for t in d.replace("const", " ").split()[:-1]) or d
    return {"int": _c_int, "float": _c_float, "long long": _c_ll}[base]


def parse_header(path=HEADER_PATH):
    """name -> [ctypes argtypes] for every `int sg2_*(...)` prototype in the public header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    sigs = {}
    for m in re.finditer(r"\bint\s+(sg2_\w+)\s*\(([^)]*)\)\s*;", text):
        args = m.group(2).strip()
        sigs[m.group(1)] = [] if args in ("", "void") else [_ctype(a) for a in args.split(",")]
    return sigs


SIGNATURES = parse_header()
if _PROBES:
    SIGNATURES.update(parse_header(os.path.join(_PKG, "..", "include", "sg2b200_probes.h")))

_lib = None


def lib():
    """Load (once) and return the shared library; fail loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU/PyTorch fallback for the sg2b200 kernels.")
        l = ctypes.CDLL(LIB_PATH)
        l.sg2_last_error.restype = ctypes.c_char_p
        l.sg2_last_error.argtypes = []
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _c_int
        _lib = l
        if os.environ.get("SG2_SM_RESERVE_ALWAYS"):      # tests: run every kernel with the data-parallel grid sizes
            check(l.sg2_set_sm_reserve(int(os.environ["SG2_SM_RESERVE_ALWAYS"])), "sg2_set_sm_reserve")
    return _lib


class NoFuse(RuntimeError):
    """SG2_ENOFUSE: the requested epilogue fusion does not apply to this shape (the caller runs the separate kernel)."""


def check(rc, what):
    if rc != 0:
        msg = lib().sg2_last_error().decode(errors="replace")
        if rc == -3:
            raise NoFuse(f"sg2b200: {what}: {msg}")
        raise RuntimeError(f"sg2b200: {what} failed (rc={rc}): {msg}")


def call(name, *args):
    check(getattr(lib(), name)(*args), name)
