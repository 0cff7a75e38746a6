import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_unavailable():
    try:
        from multiview_stitcher_b200 import _lib

        _lib.load(require_device=True)
        return None
    except Exception as e:  # EngineUnavailable: library missing or no CUDA device
        return str(e)


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without CUDA skips the gpu-marked tests
    instead of failing in them (`-m gpu` on a GPU box runs them all)."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    why = _gpu_unavailable()
    if why is None:
        return
    skip = pytest.mark.skip(reason=f"no usable CUDA engine: {why[:120]}")
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def fusion_golden():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "fusion_golden.npz"))


@pytest.fixture(scope="session")
def registration_golden():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "registration_golden.npz"))
