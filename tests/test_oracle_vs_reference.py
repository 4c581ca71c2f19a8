"""Live differential test of the oracle against the unmodified reference, run only where the
reference checkout is mounted (the build container).  Complements the committed golden vectors
with fresh random cases."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("GONZAG_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "gonzag")),
                                reason="reference checkout not present")

from gonzag_b200 import synthetic as syn
from oracle import gz_oracle as orc


@pytest.fixture(scope="module")
def gz():
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    g, _ = make_golden.load_reference(REF)
    return g


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


@pytest.mark.parametrize("seed", [0, 1])
def test_random_tracks_on_distorted_grid(gz, seed):
    rng = np.random.default_rng(seed)
    lat, lon = syn.orca_like_grid(180, 362, pole_shift_deg=float(rng.uniform(10, 30)))   # 1 degree: r = 2
    kew = 2 if seed == 0 else -1
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    t, y, x = syn.orbit_track(4000, phase=float(rng.uniform(0, 6)), lon0_deg=float(rng.uniform(0, 360)), **syn.SARAL)
    with quiet():
        ang = gz.GridAngle(lat, lon)
        ref = gz.BilinTrack(y, x, lat, lon, src_grid_local_angle=ang, k_ew_per=kew, rd_found_km=rd, np_box_r=r)
    np.testing.assert_allclose(orc.grid_angle(lat, lon), ang, rtol=0, atol=1e-11)
    mine = orc.TrackMapping(y, x, lat, lon, src_grid_local_angle=ang, k_ew_per=kew, rd_found_km=rd, np_box_r=r)
    np.testing.assert_array_equal(mine.NP, ref.NP)
    np.testing.assert_array_equal(mine.SM, ref.SM)
    np.testing.assert_array_equal(mine.WB, ref.WB)          # same machine: bit-identical


def test_scalar_functions(gz):
    rng = np.random.default_rng(5)
    for _ in range(300):
        a = rng.uniform(-89, 89, 2)
        b = rng.uniform(0, 360, 2)
        assert orc.heading(a[0], b[0], a[1], b[1]) == gz.Heading(a[0], b[0], a[1], b[1])
        assert orc.haversine_scalar_km(a[0], b[0], a[1], b[1]) == gz.Haversine_sclr(a[0], b[0], a[1], b[1])
        assert orc.frame_pm180(b[0]) == gz.degE_to_degWE(b[0])
    assert orc.search_box_radius(9.26, 500.) == gz.SearchBoxSize(9.26, 500.)
