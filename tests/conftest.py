import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    return os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no NVIDIA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _exact_smooth_arithmetic(request):
    """GPU tests compare Smooth with SciPy bit for bit: they run with the exact tap arithmetic
    (GM_SMOOTH_EXACT).  The default (fused multiply-add) and the float32 arithmetic have their own
    tolerance tests in tests/test_spatial_gpu.py."""
    if "gpu" not in request.keywords or not _has_gpu():
        yield
        return
    from dask_geomodeling_b200 import _native

    with _native.smooth_arithmetic("exact"):
        yield
