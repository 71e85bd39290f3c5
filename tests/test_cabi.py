"""CPU checks of the drop-in boundary: the library loads, exports every symbol the header declares, and fails
loudly (no CPU fallback) when CUDA is not usable.  No compute calls here."""
import ctypes
import os
import re

import pytest

import floor_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(built_lib):
    header = open(os.path.join(ROOT, "include", "floor_b200_mip.h")).read()
    declared = sorted(set(re.findall(r"\b(flmip_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in include/floor_b200_mip.h but not exported"
    assert sorted(declared) == sorted(floor_b200.EXPORTS)


def test_no_torch_types_in_the_abi():
    header = open(os.path.join(ROOT, "include", "floor_b200_mip.h")).read()
    assert "torch" not in header and "at::" not in header and "std::" not in header
    assert 'extern "C"' in header


def test_every_entry_point_cites_the_reference():
    header = open(os.path.join(ROOT, "include", "floor_b200_mip.h")).read()
    for ref in ["device_image.cpp:235-328", "mip_map_minify.hpp:89-126", "cuda_image.cpp:158-539", "cuda_image.cpp:588-673",
                "cuda_api.cpp", "cuda_context.cpp", "cuda_queue.cpp"]:
        assert ref in header, ref


def test_no_cpu_fallback_without_cuda(built_lib):
    """On a box without a GPU the compute path must raise, not compute on the CPU."""
    if built_lib.flmip_init() == 0:
        pytest.skip("a CUDA device is present")
    assert built_lib.flmip_device_count() == 0
    assert b"cuda" in built_lib.flmip_last_error_string().lower()
    with pytest.raises(floor_b200.FlmipError):
        floor_b200.device_context()
    img = ctypes.c_void_p()
    dim = (ctypes.c_uint32 * 4)(64, 64, 0, 0)
    rc = built_lib.flmip_image_create(0, 0x802FC12, dim, 0, 0, ctypes.byref(img))
    assert rc == floor_b200.ERR_NO_CUDA and not img.value


def test_product_does_not_import_the_oracle():
    """only tests/, smoke() and bench.py's cpu legs may import / link / execute anything under oracle/"""
    bad = re.compile(r"import\s+oracle|from\s+oracle|liboracle|#include\s*[\"<][^\n]*oracle|dlopen\([^\n]*oracle|flo_generate|flo_fill")
    for top in ("floor_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for fn in files:
                if fn.endswith((".py", ".cpp", ".cu", ".h", ".hpp", ".S", "Makefile")):
                    src = open(os.path.join(dirpath, fn), errors="replace").read()
                    assert not bad.search(src), os.path.join(dirpath, fn)


def test_cubin_is_sm100a_and_uses_tma():
    cubin = os.path.join(ROOT, "floor_b200", "csrc", "mip_kernels.cubin")
    if not os.path.exists(cubin):
        floor_b200.build()
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", "-fun", "flmip_fast2d_k1_c4", cubin], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100a" in out or "EF_CUDA_SM100" in out or "sm_100" in out
    assert "UTMALDG" in out, "the single-pass kernel must load its tile with TMA"
    assert "SYNCS" in out  # mbarrier
