"""The C++ drop-in (include/floor_b200/floor_b200.hpp): builds with g++ -std=c++20 against the C-ABI library, its enums /
size arithmetic match, and on a GPU box it runs the reference's usage pattern against the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


def build(built_lib):
    subprocess.check_call(["make", "-s", "-C", CPP])
    return os.path.join(CPP, "dropin_test")


def test_dropin_builds_and_cpu_checks(built_lib):
    exe = build(built_lib)
    res = subprocess.run([exe, "--cpu"], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "ok" in res.stdout


def test_dropin_headers_cite_the_reference():
    src = open(os.path.join(ROOT, "include", "floor_b200", "floor_b200.hpp")).read()
    for ref in ["device_image.hpp:161-162", "device_image.cpp:235-328", "cuda_image.cpp:588-673", "cuda_image.cpp:703-769", "device_context.hpp:261-267"]:
        assert ref in src, ref


@pytest.mark.gpu
def test_dropin_parity_on_gpu(built_lib, oracle_mod):
    exe = build(built_lib)
    oracle_so = os.path.join(ROOT, "oracle", "liboracle_minify.so")
    res = subprocess.run([exe, oracle_so], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "dropin_test: ok" in res.stdout


@pytest.mark.gpu
def test_cpp_bench_program_runs(built_lib):
    """tests/cpp/mipbench.cpp: the C++ host API (cuda_context -> queue -> image -> generate_mip_map_chain) timed with the queue's own
    profiling; a 8192^2 RGBA16F chain has to stay far above anything a CPU path could do (no silent fallback)"""
    subprocess.check_call(["make", "-s", "-C", CPP, "mipbench"])
    res = subprocess.run([os.path.join(CPP, "mipbench"), "c2", "10"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stdout + res.stderr
    gbs = float(res.stdout.split("enqueued")[1].split("=")[1].split("GB/s")[0])
    assert gbs > 1000.0, res.stdout
    # the overlapped leg (device_queue::set_mip_chain_overlap) must not be slower than stream-ordered chains
    over = float(res.stdout.split("overlapped")[1].split("=")[1].split("GB/s")[0])
    assert over > 0.98 * gbs, res.stdout
