"""Loader for the product's CUDA library (dbox_b200/libdbox_b200.so).  There is no CPU fallback: if the library is
missing or no CUDA device is usable, callers get an exception."""
import ctypes as C
import os
import subprocess

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdbox_b200.so")
_api = None


def build(force=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    deps = [os.path.join(src, f) for f in os.listdir(src) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(_HERE, "..", "include", "dbox_b200.h"))
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        subprocess.check_call(["make", "-C", src, "-s", "-j4", "../libdbox_b200.so"])
    return LIB_PATH


def api():
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libdbox_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                               "dbox_b200 has no CPU fallback")
        _api = A.Api(C.CDLL(LIB_PATH), "dbx_")
    return _api
