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


@pytest.fixture(scope="session")
def golden_grid():
    return load_golden("grid")


def grid_case_sources(g, tag):
    """The in-memory data sets of tests/golden/make_golden.py:case_grid, rebuilt from the fixture."""
    from gonzag_b200.sources import ModelArrays, TrackArrays
    if tag == "a":
        lat, lon = g["a_lat"].astype(np.float64), g["a_lon"].astype(np.float64)
        model = ModelArrays(g["a_tm"], lat, lon, fields={"ssh": g["a_f32"]})
        lsm = ModelArrays(g["a_tm"], None, None, masks={"tmask": g["a_maskin"].astype(np.int64)})
        return model, TrackArrays(g["a_t"], g["a_y"], g["a_x"], {"ssh": g["a_ssat"]}), lsm, "tmask", True, len(g["a_t"])
    if tag == "b":
        model = ModelArrays(g["b_tm"], g["b_lat"], g["b_lon"], fields={"ssh": g["b_f64"]},
                            masks={"tmask": g["b_maskin"].astype(np.int64)})
        return model, TrackArrays(g["b_t"], g["b_y"], g["b_x"], {"ssh": g["b_ssat"]}), model, "tmask", False, 100
    fcm = np.ma.masked_array(g["c_f32"], mask=np.broadcast_to(g["c_maskin"] == 0, g["c_f32"].shape))
    model = ModelArrays(g["c_tm"], g["c_lat1"], g["c_lon1"], fields={"ssh": fcm})
    return model, TrackArrays(g["c_t"], g["c_y"], g["c_x"], {"ssh": g["c_ssat"]}), model, "_FillValue@ssh", False, len(g["c_t"])


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
