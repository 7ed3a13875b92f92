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
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    cases = {}
    for k in z.files:
        if "/" in k:
            c, f = k.split("/", 1)
            cases.setdefault(c, {})[f] = z[k]
        else:
            cases[k] = z[k]
    return cases


@pytest.fixture(scope="session")
def msda_golden():
    return load_golden("msda_golden.npz")


@pytest.fixture(scope="session")
def hungarian_golden():
    return load_golden("hungarian_golden.npz")


@pytest.fixture(scope="session")
def lsap_golden():
    return load_golden("lsap_golden.npz")


@pytest.fixture(scope="session")
def ema_golden():
    return load_golden("ema_golden.npz")
