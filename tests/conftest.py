import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import stereo_b200._lib as L
        return L.lib().sb_device_count() > 0
    except Exception:
        return False


HAVE_GPU = None


def pytest_collection_modifyitems(config, items):
    global HAVE_GPU
    if HAVE_GPU is None:
        HAVE_GPU = _have_gpu()
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_everything():
    """Make sure the product library and the oracle libraries exist (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
