import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    """GPU tests go through libpsim_b200.so; they fail (not skip) when it cannot run."""
    import ctypes as C
    from particlesim_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.psim_create(0, 16, 16, None, C.byref(h))
    assert rc == 0, "psim_create failed: no CUDA device visible to libpsim_b200.so"
    lib.psim_destroy(h)
    return 0
