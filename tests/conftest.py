import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        from motivate_b200 import _ffi
        return _ffi.load().mvb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # A `-m gpu` run on a box without a device must FAIL, not skip: the product has no CPU path.
    if config.getoption("-m") and "not gpu" in config.getoption("-m"):
        return
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here (GPU tests run under gpurun)")
    selected_gpu = config.getoption("-m") == "gpu"
    for item in items:
        if "gpu" in item.keywords and not selected_gpu:
            item.add_marker(skip)
