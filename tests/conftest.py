import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def built_lib():
    """libfloor_b200_mip.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    import floor_b200
    if not os.path.exists(floor_b200.LIB_PATH):
        floor_b200.build()
    return floor_b200.lib()


@pytest.fixture(scope="session")
def gpu_ctx(built_lib):
    """device_context + queue on GPU 0; fails loudly (no skip) when the CUDA path is unavailable."""
    import floor_b200
    ctx = floor_b200.device_context()
    dev = ctx.get_device(0)
    q = ctx.create_queue(dev)
    return ctx, dev, q
