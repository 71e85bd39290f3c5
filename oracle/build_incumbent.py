"""Builds the pieces bench.py --impl incumbent needs (BENCH INFRASTRUCTURE ONLY, never imported by floor_b200):

  oracle/_ref/mmm_incumbent.ptx       the PTX of the reference's prebuilt CUDA minify kernels, extracted from
                                      /root/reference/etc/mip_map_minify/mmm.fubar by oracle/fubar_extract.cpp, which is compiled
                                      together with the reference's own src/core/bcm.cpp where it lies (the archive is BCM-compressed)
  oracle/_ref/libfloor_incumbent.so   oracle/incumbent_harness.cpp: the reference's host side for these kernels through the driver API

Outputs only go to oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).  Nothing of the reference is copied into the
repository.  /root/reference is only needed for the extraction; without it the prebuilt files are used as they are.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("FLOOR_REFERENCE", "/root/reference")
PTX = os.path.join(OUT, "mmm_incumbent.ptx")
LIB = os.path.join(OUT, "libfloor_incumbent.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _newer(dst, *srcs):
    return os.path.exists(dst) and all(os.path.getmtime(dst) >= os.path.getmtime(s) for s in srcs)


def build(force: bool = False) -> bool:
    os.makedirs(OUT, exist_ok=True)
    fubar = os.path.join(REF, "etc", "mip_map_minify", "mmm.fubar")
    ext_src = os.path.join(HERE, "fubar_extract.cpp")
    if os.path.exists(fubar) and (force or not _newer(PTX, ext_src)):
        exe = os.path.join(OUT, "fubar_extract")
        subprocess.check_call(["g++", "-std=c++23", "-O1", "-I" + os.path.join(REF, "include"), ext_src, os.path.join(REF, "src", "core", "bcm.cpp"), "-o", exe])
        subprocess.check_call([exe, fubar, PTX], stdout=subprocess.DEVNULL)
    har_src = os.path.join(HERE, "incumbent_harness.cpp")
    if force or not _newer(LIB, har_src):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I" + os.path.join(CUDA_HOME, "include"), har_src, "-o", LIB, "-ldl"])
    return os.path.exists(PTX) and os.path.exists(LIB)


if __name__ == "__main__":
    print("incumbent:", "ready" if build(force=True) else "unavailable")
