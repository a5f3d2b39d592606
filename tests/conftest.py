import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def aal90():
    """The reference's input DATA (SC, empirical FCs, NA/ACh maps): data/aal90_inputs.npz, not a golden output."""
    return np.load(os.path.join(ROOT, "data", "aal90_inputs.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import cwrap
    cwrap.build()
    return cwrap
