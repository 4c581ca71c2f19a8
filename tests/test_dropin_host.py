"""Host side of the rows SURVEY.md 8f lists next to the hot path: the grid heuristics and the
time-window scan (pinned to the reference's own answers, tests/golden/golden_grid.npz, which
include the arrays of the reference's tests/test_functions.py), `SatTrack` on in-memory sources,
the reader stand-ins and the NetCDF writers.  No GPU needed."""
import os

import numpy as np
import pytest

import gonzag_b200 as gz
from gonzag_b200.sources import ModelArrays, TrackArrays
from tests.conftest import grid_case_sources


def test_longitude_heuristics_known_answers(golden_grid):
    g = golden_grid
    for k in range(8):
        x = g["kat%d_lon" % k]
        res, lglob, l360, xmin, xmax, ew = g["kat%d_out" % k]
        r = gz.GridResolution(x) if k < 3 else 1.0
        assert r == res
        got = gz.IsGlobalLongitudeWise(x, resd=r)
        assert (bool(got[0]), bool(got[1]), float(got[2]), float(got[3])) == (bool(lglob), bool(l360), xmin, xmax)
        if k < 3:
            assert gz.IsEastWestPeriodic(x) == int(ew)
    # the three answers the reference's test prints for its periodic arrays: 0, 1 and 2 overlapping points
    assert [int(g["kat%d_out" % k][5]) for k in range(3)] == [0, 1, 2]
    # and the tuples written in its comments (tests/test_functions.py:66-84)
    assert tuple(g["kat4_out"][1:5]) == (0.0, 0.0, -20.0, 6.0)
    assert tuple(g["kat5_out"][1:5]) == (0.0, 1.0, 169.0, 201.0)


def test_scan_idx_matches_reference_loops(golden_grid):
    g = golden_grid
    for (a, b), want in zip(g["scan_q"], g["scan_out"]):
        assert gz.scan_idx(g["scan_vt"], a, b) == tuple(int(v) for v in want)
    with pytest.raises(ValueError):
        gz.scan_idx(np.array([1.0]), 0.0, 2.0)


def test_reader_slicing_quirk():
    """kt1 / kt2 are honoured only when both are > 0 (gonzag/ncio.py:119-125)."""
    t = np.arange(10.0)
    src = TrackArrays(t, t, t, {"ssh": np.ma.masked_array(t, mask=t == 3)})
    assert len(src.GetTimeEpochVector(kt1=0, kt2=4)) == 10
    assert src.GetTimeEpochVector(kt1=2, kt2=4).tolist() == [2.0, 3.0, 4.0]
    assert src.GetTimeEpochVector(isubsamp=4).tolist() == [0.0, 4.0, 8.0]
    assert src.GetSatCoor("latitude", 5, 6).tolist() == [5.0, 6.0]
    v = src.GetSatSSH("ssh", kt1=2, kt2=5, ikeep=(np.array([0, 1, 3]),))
    assert v.tolist() == [2.0, -9999.0, 5.0]
    assert src.GetTimeInfo() == (10, (0.0, 9.0))
    m = ModelArrays(t, np.zeros(3), np.zeros(4), fields={"f": np.ma.masked_array(np.ones((10, 3, 4)), mask=False)})
    assert m.GetModelLSM("_FillValue@f").dtype == int and m.GetModelLSM("_FillValue@f").sum() == 12
    m2 = ModelArrays(t, np.zeros(3), np.zeros(4), fields={"f": lambda k: np.full((3, 4), np.nan if k == 0 else 1.0)})
    assert m2.GetModelLSM("_FillValue@f").sum() == 0 and m2.field_array("f") is None


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_sattrack_matches_reference(golden_grid, tag):
    g = golden_grid
    model, track, lsm, varlsm, dist, Np = grid_case_sources(g, tag)
    (it1, it2), (nts, ntm) = gz.GetEpochTimeOverlap(track, model)
    assert [it1, it2] == g[tag + "_it"].tolist() and [nts, ntm] == g[tag + "_nts_ntm"].tolist()
    bounds = g[tag + "_mg_bounds"].tolist()
    l360 = bool(g[tag + "_mg_scal"][3])
    ST = gz.SatTrack(track, it1, it2, Np=Np, domain_bounds=bounds, l_0_360=l360)
    assert [ST.jt1, ST.jt2, ST.size] == g[tag + "_st_j"].tolist()
    np.testing.assert_array_equal(ST.keepit[0], g[tag + "_st_keepit"])
    np.testing.assert_array_equal(ST.time, g[tag + "_st_time"])
    np.testing.assert_array_equal(ST.lat, g[tag + "_st_lat"])
    np.testing.assert_array_equal(ST.lon, g[tag + "_st_lon"])


def test_no_time_overlap_exits():
    a = TrackArrays(np.arange(10.0), np.zeros(10), np.zeros(10))
    b = ModelArrays(100.0 + np.arange(3.0), np.zeros(3), np.zeros(3))
    with pytest.raises(SystemExit):
        gz.GetEpochTimeOverlap(a, b)


def test_file_names_need_the_reference_io():
    with pytest.raises(RuntimeError):
        gz.SatTrack("/nonexistent/track.nc", 0.0, 1.0)


def test_netcdf_writers(tmp_path):
    from scipy.io import netcdf_file
    t = 1.0e9 + np.arange(50.0)
    xd = np.ma.masked_array(np.random.default_rng(0).normal(size=(3, 50)), mask=False)
    xd[1, 7] = np.ma.masked
    path = str(tmp_path / "result.nc")
    assert gz.SaveTimeSeries(t, xd, ["latitude", "ssh_np", "ssh_bl"], path, time_units="seconds since 1970-01-01 00:00:00",
                             vunits=["deg.N", "m", "m"], vlnm=["Latitude", "a", "b"], missing_val=-9999.) == 0
    with netcdf_file(path, "r", mmap=False) as f:
        assert f.variables["time"].units == b"seconds since 1970-01-01 00:00:00"
        np.testing.assert_array_equal(f.variables["time"][:], t)
        assert f.variables["ssh_np"][7] == np.float32(-9999.)
        np.testing.assert_allclose(f.variables["ssh_bl"][:], xd[2].astype(np.float32))
        assert f.variables["ssh_np"].units == b"m"
    fld = np.arange(12.0).reshape(3, 4)
    msk = np.ones((3, 4)); msk[0, 0] = 0
    p2 = str(tmp_path / "xnp_msk.nc")
    gz.Save2Dfield(p2, fld, xlon=fld, xlat=fld, name="track", mask=msk)
    with netcdf_file(p2, "r", mmap=False) as f:
        v = f.variables["track"][:]
        assert np.isnan(v[0, 0]) and v[2, 3] == 11.0
        assert f.variables["nav_lon"].shape == (3, 4)
    with pytest.raises(ValueError):
        gz.SaveTimeSeries(t[:3], xd, ["a", "b", "c"], path)


def test_periodicity_test_hardening():
    """`IsEastWestPeriodic(X, tol)`: tol = 0 is the reference's exact equality (kept as the default, pinned by the
    known answers above); a tolerance recovers the overlap of grids the exact test misses."""
    from gonzag_b200 import synthetic as syn
    for kew in (0, 1, 2):
        nx = 90 + kew
        lon = np.tile((4.0 * np.arange(nx)) % 360.0, (5, 1))              # exact in binary: both tests agree
        assert gz.IsEastWestPeriodic(lon) == kew and gz.IsEastWestPeriodic(lon, tol=0.01) == kew
    step = 360.0 / 4320.0                                                  # 1/12 degree: not a binary fraction
    lon = np.tile(step * np.arange(4322), (3, 1))
    lon32 = lon.astype(np.float32).astype(np.float64)                      # "file-like": through float32
    assert gz.IsEastWestPeriodic(lon32) == -1                              # the reference's answer
    assert gz.IsEastWestPeriodic(lon32, tol=0.01) == 2
    assert gz.IsEastWestPeriodic(lon, tol=0.01) == 2
    assert gz.IsEastWestPeriodic(np.tile(np.linspace(100.0, 140.0, 81), (3, 1)), tol=0.01) == -1      # regional: still none
    lat, glon = syn.orca_like_grid(60, 122)
    assert gz.IsEastWestPeriodic(glon, tol=0.01) in (-1, 0, 1, 2)
