"""ctypes front end of the CPU oracle (oracle/minify_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py -- never by the floor_b200 package.

Pinned against the reference itself: the reference ships no golden vectors for this path, but its Host-Compute minify
kernels, software sampler and image-size helpers compile with g++ through oracle/build_ref.py (-> oracle/_ref, front end
oracle/ref.py); tests/test_reference_pin.py requires this restatement to match them bit for bit, and
tests/golden/golden_ref.json freezes reference-computed chains ("Oracle").
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_minify.so")
_lib = None

FLAG_NO_DOUBLE = 1


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, strict IEEE flags)."""
    src = os.path.join(_HERE, "minify_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "liboracle_minify.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        L.flo_bytes_per_pixel.restype = ctypes.c_uint32
        L.flo_bytes_per_pixel.argtypes = [ctypes.c_uint64]
        L.flo_mip_level_count.restype = ctypes.c_uint32
        L.flo_mip_level_count.argtypes = [u32p, ctypes.c_uint64]
        L.flo_layer_count.restype = ctypes.c_uint32
        L.flo_layer_count.argtypes = [u32p, ctypes.c_uint64]
        L.flo_level_dim.restype = None
        L.flo_level_dim.argtypes = [u32p, ctypes.c_uint64, ctypes.c_uint32, u32p]
        L.flo_level_size.restype = ctypes.c_uint64
        L.flo_level_size.argtypes = [u32p, ctypes.c_uint64, ctypes.c_uint32]
        L.flo_level_offset.restype = ctypes.c_uint64
        L.flo_level_offset.argtypes = [u32p, ctypes.c_uint64, ctypes.c_uint32]
        L.flo_effective_level_count.restype = ctypes.c_uint32
        L.flo_effective_level_count.argtypes = [u32p, ctypes.c_uint64, ctypes.c_uint32]
        L.flo_image_data_size.restype = ctypes.c_uint64
        L.flo_image_data_size.argtypes = [u32p, ctypes.c_uint64, ctypes.c_uint32]
        L.flo_generate_mip_map_chain.restype = ctypes.c_int
        L.flo_generate_mip_map_chain.argtypes = [ctypes.c_void_p, u32p, ctypes.c_uint64, ctypes.c_uint32,
                                                 ctypes.c_uint32, ctypes.c_uint32]
        L.flo_synth_element.restype = ctypes.c_uint32
        L.flo_synth_element.argtypes = [ctypes.c_uint64] * 4
        L.flo_fill_synthetic.restype = None
        L.flo_fill_synthetic.argtypes = [ctypes.c_void_p, u32p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                         ctypes.c_uint32]
        L.flo_half_to_float.restype = ctypes.c_float
        L.flo_half_to_float.argtypes = [ctypes.c_uint16]
        L.flo_float_to_half.restype = ctypes.c_uint16
        L.flo_float_to_half.argtypes = [ctypes.c_float]
        _lib = L
    return _lib


def _dim(dim):
    d = list(dim) + [0] * (4 - len(dim))
    return (ctypes.c_uint32 * 4)(*d)


def bytes_per_pixel(image_type: int) -> int:
    return lib().flo_bytes_per_pixel(image_type)


def mip_level_count(dim, image_type: int, mip_level_limit: int = 0) -> int:
    return lib().flo_effective_level_count(_dim(dim), image_type, mip_level_limit)


def layer_count(dim, image_type: int) -> int:
    return lib().flo_layer_count(_dim(dim), image_type)


def level_dim(dim, image_type: int, level: int):
    out = (ctypes.c_uint32 * 3)()
    lib().flo_level_dim(_dim(dim), image_type, level, out)
    return tuple(out)


def level_size(dim, image_type: int, level: int) -> int:
    return lib().flo_level_size(_dim(dim), image_type, level)


def level_offset(dim, image_type: int, level: int) -> int:
    return lib().flo_level_offset(_dim(dim), image_type, level)


def image_data_size(dim, image_type: int, mip_level_limit: int = 0) -> int:
    return lib().flo_image_data_size(_dim(dim), image_type, mip_level_limit)


def generate_mip_map_chain(level0, dim, image_type: int, mip_level_limit: int = 0, no_double: bool = False,
                           threads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
    """Runs the restated Host-Compute chain.  `level0` = bytes of level 0 (all layers); returns the whole
    level-major image buffer (uint8) with levels >= 1 generated."""
    total = image_data_size(dim, image_type, mip_level_limit)
    l0 = np.ascontiguousarray(level0).view(np.uint8).reshape(-1)
    n0 = level_size(dim, image_type, 0)
    if l0.size < n0:
        raise ValueError(f"level 0 needs {n0} bytes, got {l0.size}")
    buf = out if out is not None else np.zeros(total, dtype=np.uint8)
    if buf.size < total:
        raise ValueError("output buffer too small")
    buf[:n0] = l0[:n0]
    rc = lib().flo_generate_mip_map_chain(buf.ctypes.data, _dim(dim), image_type, mip_level_limit,
                                          FLAG_NO_DOUBLE if no_double else 0, threads)
    if rc != 0:
        raise RuntimeError(f"oracle: unsupported image type {image_type:#x} (rc={rc})")
    return buf


def generate_in_place(buf: np.ndarray, dim, image_type: int, mip_level_limit: int = 0, no_double: bool = False,
                      threads: int = 1) -> None:
    rc = lib().flo_generate_mip_map_chain(buf.ctypes.data, _dim(dim), image_type, mip_level_limit,
                                          FLAG_NO_DOUBLE if no_double else 0, threads)
    if rc != 0:
        raise RuntimeError(f"oracle: unsupported image type {image_type:#x} (rc={rc})")


def fill_synthetic(dim, image_type: int, config_id: int, layer_id0: int = 0, layer_num: int | None = None,
                   out: np.ndarray | None = None) -> np.ndarray:
    """Level-0 bytes of `layer_num` layers whose global ids start at `layer_id0` (counter-based)."""
    layers = layer_count(dim, image_type) if layer_num is None else layer_num
    per_layer = level_size(dim, image_type, 0) // max(layer_count(dim, image_type), 1)
    n = per_layer * layers
    buf = out if out is not None else np.empty(n, dtype=np.uint8)
    lib().flo_fill_synthetic(buf.ctypes.data, _dim(dim), image_type, config_id, layer_id0, layers)
    return buf[:n]
