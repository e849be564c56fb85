import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def _has_gpu():
    try:
        from slslam_b200 import capi
        return os.path.exists(capi.LIB_PATH) and capi.lib().slslam_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not _has_gpu():
        pytest.fail("no sm_100 device / library not built: GPU tests must not silently pass")
    from slslam_b200 import capi
    return capi
