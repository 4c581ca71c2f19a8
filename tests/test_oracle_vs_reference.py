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
def ref():
    """(the imported reference package, the store behind its in-memory `ncio` stand-in)"""
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    return make_golden.load_reference(REF)


@pytest.fixture(scope="module")
def gz(ref):
    return ref[0]


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


@pytest.mark.parametrize("seed,dtype,kew", [(3, np.float32, 2), (4, np.float64, -1)])
def test_model2sat_products_live(ref, seed, dtype, kew):
    """Everything `Model2SatTrack` exposes (gonzag/mod2sat.py:20-186) on a fresh random case: unmodified reference
    (in-memory `ncio` stand-in of make_golden.py) against the oracle's `track_products`, same machine => same bits."""
    import make_golden
    gz, store = ref
    rng = np.random.default_rng(seed)
    lat, lon = syn.orca_like_grid(120, 182, pole_shift_deg=float(rng.uniform(10, 30)))       # 2 degrees: r = 1
    mask = syn.land_mask(lat, lon, seed=seed)
    with quiet():
        ang = gz.GridAngle(lat, lon)
        res = float(gz.GridResolution(lon))
    t, y, x = syn.orbit_track(6000, t0=500.0, phase=float(rng.uniform(0, 6)), lon0_deg=float(rng.uniform(0, 360)), **syn.JASON)
    tmod = np.array([0.0, 2500.0, 5000.0, 9000.0])
    fld = syn.ssh_records(lat, lon, 4, dtype=np.float32).astype(dtype)
    ssat = make_golden.sat_ssh(len(t), seed + 20)
    R = make_golden.run_model2sat(gz, store, lat, lon, mask, ang, kew, tmod, fld, t, y, x, ssat.copy(), res)
    P = orc.track_products(lat, lon, mask, tmod, lambda k: fld[k], t, y, x, ssat.copy(), res, angle=ang, k_ew_per=kew)
    np.testing.assert_array_equal(P["mask"], R.mask)
    for name in ("ssh_mod_np", "ssh_mod", "ssh_sat", "XNPtrack"):
        a, b = P[name], getattr(R, name)
        np.testing.assert_array_equal(np.ma.getmaskarray(a), np.ma.getmaskarray(b), err_msg=name)
        np.testing.assert_array_equal(np.ma.getdata(a), np.ma.getdata(b), err_msg=name)
    np.testing.assert_array_equal(P["distance"], R.distance)
    assert (P["NP"][:, 0] >= 0).sum() > 3000 and (R.mask == 1).sum() > 1000
