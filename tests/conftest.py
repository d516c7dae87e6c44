import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_library():
    """The in-tree CUDA library; built on demand (nvcc cross-compiles without a GPU)."""
    from vehiclemodelvisualodometry_b200 import build as b

    return b.build_library()


@pytest.fixture(scope="session")
def cuda_device(built_library):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("a test marked gpu ran without a CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


@pytest.fixture
def tuning(cuda_device):
    """``tuning(key, value)``: the library's test hook (vmvo_debug_set_tuning) on device 0's ctx;
    every override is back at the library's own choice when the test ends."""
    from vehiclemodelvisualodometry_b200 import _lib

    ctx = _lib.context(0)
    used = set()

    def set_(key, value):
        used.add(key)
        ctx.set_tuning(key, value)

    yield set_
    for key in used:
        ctx.set_tuning(key, -1)
