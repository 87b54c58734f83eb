import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def f2d():
    """The product package (fluid-2d_b200/), library built if necessary."""
    import fluid2d_b200

    fluid2d_b200.build()
    return fluid2d_b200


@pytest.fixture(scope="session")
def sfo():
    """The plain-C oracle (TEST INFRASTRUCTURE)."""
    from oracle import sfo as _sfo

    _sfo.build()
    return _sfo


@pytest.fixture(scope="session")
def gpu_ok(f2d):
    if f2d.device_count() < 1:
        pytest.fail("GPU test selected but libf2d sees no CUDA device (there is no CPU fallback)")
    return True
