"""ctypes binding of the C-ABI shared library (include/eavit_b200.h).

The library is the product: there is NO CPU or PyTorch fallback.  If ``libeavit_b200.so`` is missing
or a call returns a non-zero status a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EAVIT_B200_LIB") or os.path.join(_HERE, "libeavit_b200.so")   # override: instrumented builds (tools/)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "eavit_b200.h")

_lib = None


class GemmArgs(ctypes.Structure):
    """Mirror of ``eavit_gemm_args`` (include/eavit_b200.h)."""
    _fields_ = [
        ("M", ctypes.c_int), ("N", ctypes.c_int), ("K", ctypes.c_int),
        ("A", ctypes.c_void_p), ("lda", ctypes.c_longlong), ("a_mn", ctypes.c_int),
        ("B", ctypes.c_void_p), ("ldb", ctypes.c_longlong), ("b_mn", ctypes.c_int),
        ("bias", ctypes.c_void_p), ("aux_bf16", ctypes.c_void_p), ("residual", ctypes.c_void_p),
        ("out_f32", ctypes.c_void_p), ("out_bf16", ctypes.c_void_p), ("out_pre_bf16", ctypes.c_void_p),
        ("colsum", ctypes.c_void_p),
        ("ldc", ctypes.c_longlong), ("drop_p", ctypes.c_float), ("drop_seed", ctypes.c_ulonglong),
        ("act", ctypes.c_int), ("atomic_f32", ctypes.c_int), ("split_k", ctypes.c_int),
        ("ln_gamma", ctypes.c_void_p), ("ln_beta", ctypes.c_void_p), ("ln_mean", ctypes.c_void_p), ("ln_rstd", ctypes.c_void_p),
        ("ln_eps", ctypes.c_float),
    ]


def declared_symbols():
    """Every ``eavit_*`` function the public header declares."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(eavit_[a-z0-9_]+)\s*\(", txt)))


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a with nvcc (cross-compiles without a GPU)."""
    import subprocess
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("eavit_b200: nvcc build failed\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"eavit_b200: {LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no fallback path")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.eavit_last_error.restype = ctypes.c_char_p
        _lib.eavit_launch_count.restype = ctypes.c_longlong
        _lib.eavit_rms_workspace_bytes.restype = ctypes.c_longlong
        _lib.eavit_rms_workspace_bytes.argtypes = [ctypes.c_longlong, ctypes.c_int]
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"eavit_b200.{what} failed with status {rc}: {lib().eavit_last_error().decode()}")


_replayed = 0


def add_replayed(n: int) -> None:
    """Kernel launches executed by a CUDA-graph replay (the library only sees the launches of the capture)."""
    global _replayed
    _replayed += int(n)


def launch_count(include_replays: bool = True) -> int:
    """Kernels of this library launched so far: direct launches counted inside the library + kernel nodes of replayed graphs."""
    return int(lib().eavit_launch_count()) + (_replayed if include_replays else 0)
