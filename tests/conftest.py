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


@pytest.fixture(scope="session")
def fusion_golden():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "fusion_golden.npz"))


@pytest.fixture(scope="session")
def registration_golden():
    import numpy as np

    return np.load(os.path.join(GOLDEN, "registration_golden.npz"))
