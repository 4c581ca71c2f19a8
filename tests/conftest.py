import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, "golden_%s.npz" % name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_units():
    return load_golden("units")


@pytest.fixture(scope="session")
def golden_global():
    return load_golden("global")


@pytest.fixture(scope="session")
def golden_regional():
    return load_golden("regional")


@pytest.fixture(scope="session")
def golden_seam():
    return load_golden("seam")


def ulp_diff(a, b):
    """Largest distance in units-in-the-last-place between two float64 arrays."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    ia = a.view(np.int64).astype(np.int64)
    ib = b.view(np.int64).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-2**63) - ia, ia)
    ib = np.where(ib < 0, np.int64(-2**63) - ib, ib)
    if a.size == 0:
        return 0
    return int(np.max(np.abs(ia - ib)))
