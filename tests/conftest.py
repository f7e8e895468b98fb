import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "jit_hw: run-time compiled kernels on hardware (opt-in, not part of -m gpu yet)")
    # make sure the native pieces exist (no-op when up to date; nvcc cross-compiles without a GPU)
    from flecsolve_b200 import build
    build.build_all()
    # the drop-in proof library (reference headers over the device policies): needs /root/reference, no-op elsewhere
    from tests.dropin import build as dropin_build
    dropin_build.main()


def _device_count() -> int:
    from flecsolve_b200 import _lib
    return _lib.device_count()


@pytest.fixture(scope="session")
def ctx():
    from flecsolve_b200 import _lib
    if _lib.device_count() == 0:
        pytest.skip("no CUDA device")
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    return o
