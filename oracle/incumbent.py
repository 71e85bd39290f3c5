"""ctypes front end of oracle/_ref/libfloor_incumbent.so (BENCH INFRASTRUCTURE ONLY): the reference's own CUDA minify kernels
(sm_50 PTX from mmm.fubar, JIT-compiled by the driver) behind the reference's own launch loop, timed on this box."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build_incumbent

_lib = None


def available() -> bool:
    return os.path.exists(build_incumbent.PTX) and os.path.exists(build_incumbent.LIB)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build_incumbent.LIB)
        L.flinc_run.restype = ctypes.c_int
        L.flinc_run.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint32), ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                ctypes.POINTER(ctypes.c_uint64)]
        L.flinc_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def run(dim, image_type: int, level0: np.ndarray | None, warmup: int, steps: int, device: int = 0, out: np.ndarray | None = None):
    """returns (blocking ms per chain, enqueued ms per chain, launches per chain); `out` receives every level in floor's host layout"""
    d = (ctypes.c_uint32 * 4)(*(list(dim) + [0] * (4 - len(dim))))
    a, b, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_uint64()
    rc = lib().flinc_run(build_incumbent.PTX.encode(), device, image_type, d, None if level0 is None else level0.ctypes.data,
                         None if out is None else out.ctypes.data, warmup, steps, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n))
    if rc != 0:
        raise RuntimeError("incumbent: " + lib().flinc_last_error().decode(errors="replace"))
    return a.value, b.value, n.value
