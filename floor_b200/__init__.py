"""floor_b200 -- B200-native mip-chain generation behind libfloor's image API.

The product is ``libfloor_b200_mip.so`` (C-ABI in include/floor_b200_mip.h; C++ drop-in classes in
include/floor_b200/).  This Python package is glue for tests and bench.py: a ctypes binding that mirrors the
reference's names (device_context / device_queue / device_image, IMAGE_TYPE, MEMORY_FLAG).  There is no CPU
fallback: if the library is not built, or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

from .image_types import IMAGE_TYPE, MEMORY_FLAG  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfloor_b200_mip.so")
CSRC = os.path.join(_HERE, "csrc")

FLMIP_OK = 0
ERR_NO_CUDA, ERR_INVALID, ERR_UNSUPPORTED, ERR_DRIVER, ERR_OUT_OF_MEMORY = -1, -2, -3, -4, -5
IMAGE_NO_DOUBLE, IMAGE_FORCE_GENERIC, IMAGE_UNITS_ALWAYS, IMAGE_UNITS_NEVER, IMAGE_FORCE_TILED, IMAGE_NO_TMA_TILES = 1, 2, 4, 8, 16, 32
HOST_WRITE_COMBINED = 1

# every symbol include/floor_b200_mip.h declares (checked by tests/test_cabi.py without a GPU)
EXPORTS = [
    "flmip_init", "flmip_device_count", "flmip_get_device_info", "flmip_last_error_string", "flmip_launch_count",
    "flmip_stream_create", "flmip_stream_destroy", "flmip_stream_sync", "flmip_stream_set_chain_overlap", "flmip_stream_fence", "flmip_overlap_bookkeeping",
    "flmip_event_create", "flmip_event_record", "flmip_event_sync", "flmip_event_elapsed_ms", "flmip_event_destroy",
    "flmip_host_alloc", "flmip_host_alloc_ex", "flmip_host_free",
    "flmip_device_attach_context", "flmip_image_create_external", "flmip_image_download_layers",
    "flmip_image_create", "flmip_image_destroy", "flmip_image_mip_level_count", "flmip_image_layer_count",
    "flmip_image_data_size", "flmip_image_get_level_info", "flmip_image_device_ptr", "flmip_image_plan", "flmip_image_plan_tma_tile_launches", "flmip_sampler_table_entry",
    "flmip_image_upload", "flmip_image_download", "flmip_image_write", "flmip_image_zero",
    "flmip_mip_chain_generate", "flmip_mip_chain_generate_from", "flmip_image_fill_synthetic",
    "flmip_image_blit", "flmip_image_create_tiled_twin", "flmip_tiled_destroy", "flmip_image_copy_to_tiled",
    "flmip_image_copy_from_tiled", "flmip_tiled_download", "flmip_device_cu_context",
    "flmip_batch_create", "flmip_batch_generate", "flmip_batch_kernel_count", "flmip_batch_destroy",
]


class FlmipError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libfloor_b200_mip error {code}: {msg}")
        self.code = code


class DeviceInfo(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 128), ("global_mem_size", ctypes.c_uint64), ("sm_major", ctypes.c_uint32),
                ("sm_minor", ctypes.c_uint32), ("units", ctypes.c_uint32), ("max_total_local_size", ctypes.c_uint32),
                ("max_image_2d_dim", ctypes.c_uint32 * 2), ("max_image_3d_dim", ctypes.c_uint32 * 3),
                ("max_mip_levels", ctypes.c_uint32), ("driver_version", ctypes.c_uint32), ("clock_mhz", ctypes.c_uint32),
                ("mem_clock_mhz", ctypes.c_uint32), ("mem_bus_width", ctypes.c_uint32), ("l2_cache_size", ctypes.c_uint32)]


class LevelInfo(ctypes.Structure):
    _fields_ = [("dim", ctypes.c_uint32 * 3), ("offset", ctypes.c_uint64), ("size", ctypes.c_uint64),
                ("slice_size", ctypes.c_uint64)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the sm_100a cubin and the C-ABI library in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC] + (["-B"] if force else [])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 or verbose:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libfloor_b200_mip.so failed")
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """Loads the C-ABI library.  Raises if it has not been built -- never falls back to a CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FLMIP_LIB") or LIB_PATH  # FLMIP_LIB: A/B tuning builds of the same library only
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback for the mip-chain path)")
    L = ctypes.CDLL(path)
    vp, u32, u64, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
    u32p = ctypes.POINTER(u32)
    sig = {
        "flmip_init": (i32, []),
        "flmip_device_count": (i32, []),
        "flmip_get_device_info": (i32, [i32, ctypes.POINTER(DeviceInfo)]),
        "flmip_last_error_string": (ctypes.c_char_p, []),
        "flmip_launch_count": (u64, []),
        "flmip_stream_create": (i32, [i32, ctypes.POINTER(vp)]),
        "flmip_stream_destroy": (i32, [i32, vp]),
        "flmip_stream_sync": (i32, [i32, vp]),
        "flmip_stream_set_chain_overlap": (i32, [i32, vp, i32]),
        "flmip_stream_fence": (i32, [i32, vp]),
        "flmip_overlap_bookkeeping": (i32, [u64, u64, ctypes.c_uint32, ctypes.c_uint32]),
        "flmip_event_create": (i32, [i32, ctypes.POINTER(vp)]),
        "flmip_event_record": (i32, [i32, vp, vp]),
        "flmip_event_sync": (i32, [i32, vp]),
        "flmip_event_elapsed_ms": (i32, [i32, vp, vp, ctypes.POINTER(ctypes.c_float)]),
        "flmip_event_destroy": (i32, [i32, vp]),
        "flmip_host_alloc": (i32, [i32, ctypes.c_size_t, ctypes.POINTER(vp)]),
        "flmip_host_free": (i32, [i32, vp]),
        "flmip_host_alloc_ex": (i32, [i32, ctypes.c_size_t, u32, ctypes.POINTER(vp)]),
        "flmip_device_attach_context": (i32, [i32, vp]),
        "flmip_image_create_external": (i32, [i32, u64, u32p, u32, u32, u64, u64, ctypes.POINTER(vp)]),
        "flmip_image_download_layers": (i32, [vp, vp, ctypes.c_size_t, u32, u32, u32, u32, vp]),
        "flmip_image_create": (i32, [i32, u64, u32p, u32, u32, ctypes.POINTER(vp)]),
        "flmip_image_destroy": (i32, [vp]),
        "flmip_image_mip_level_count": (i32, [vp, u32p]),
        "flmip_image_layer_count": (i32, [vp, u32p]),
        "flmip_image_data_size": (i32, [vp, ctypes.POINTER(u64)]),
        "flmip_image_get_level_info": (i32, [vp, u32, ctypes.POINTER(LevelInfo)]),
        "flmip_image_device_ptr": (i32, [vp, ctypes.POINTER(u64)]),
        "flmip_image_plan": (i32, [vp, u32p, u32p, u32p]),
        "flmip_image_plan_tma_tile_launches": (i32, [vp, u32p]),
        "flmip_sampler_table_entry": (i32, [u32, u32, u32p]),
        "flmip_image_upload": (i32, [vp, vp, ctypes.c_size_t, u32, u32, vp]),
        "flmip_image_download": (i32, [vp, vp, ctypes.c_size_t, u32, u32, vp]),
        "flmip_image_write": (i32, [vp, vp, ctypes.c_size_t, u32p, u32p, u32p, u32p, vp]),
        "flmip_image_zero": (i32, [vp, vp]),
        "flmip_mip_chain_generate": (i32, [vp, vp]),
        "flmip_mip_chain_generate_from": (i32, [vp, u32, vp]),
        "flmip_image_fill_synthetic": (i32, [vp, u64, u64, vp]),
        "flmip_image_blit": (i32, [vp, vp, vp]),
        "flmip_image_create_tiled_twin": (i32, [vp, ctypes.POINTER(vp)]),
        "flmip_tiled_destroy": (i32, [i32, vp]),
        "flmip_image_copy_to_tiled": (i32, [vp, vp, u32, u32, vp]),
        "flmip_image_copy_from_tiled": (i32, [vp, vp, u32, u32, vp]),
        "flmip_tiled_download": (i32, [vp, vp, vp, ctypes.c_size_t, u32, u32, vp]),
        "flmip_device_cu_context": (i32, [i32, ctypes.POINTER(vp)]),
        "flmip_batch_create": (i32, [ctypes.POINTER(vp), u32, ctypes.POINTER(vp)]),
        "flmip_batch_generate": (i32, [vp, vp]),
        "flmip_batch_kernel_count": (i32, [vp, ctypes.POINTER(u32)]),
        "flmip_batch_destroy": (i32, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> int:
    if rc < 0:
        raise FlmipError(rc, lib().flmip_last_error_string().decode(errors="replace"))
    return rc


from .host import device_context, device_queue, device_image, pinned_buffer  # noqa: E402,F401
