"""The built-in NetCDF-3 readers (`gonzag_b200.nc3io`, call surface of gonzag/ncio.py:68-270) against
the in-memory sources on the same data, the time conversion of `ToEpochTime` (gonzag/ncio.py:20-64)
against known dates, and `GetEpochTimeOverlap` / `SatTrack` from FILE NAMES against the outputs of the
unmodified reference (tests/golden/golden_grid.npz).  No GPU needed."""
import calendar
import datetime as dt

import numpy as np
import pytest

import gonzag_b200 as gz
from gonzag_b200 import nc3io
from tests.conftest import grid_case_sources, write_grid_case_files


US = 1.0e-6 + 5e-7      # num2date's resolution (+ float rounding at 2e9 s)


def _epoch(*a):
    return float(calendar.timegm(dt.datetime(*a).timetuple()))


def test_num2epoch_known_dates():
    assert nc3io.num2epoch(1.5e9, "seconds since 1970-01-01 00:00:00", "gregorian") == 1.5e9
    assert nc3io.num2epoch(0, "seconds since 1970-01-01", "standard") == 0.0
    assert nc3io.num2epoch(613608, "hours since 1950-01-01 00:00:00", "gregorian") == _epoch(2020, 1, 1)
    assert nc3io.num2epoch(7305.5, "days since 2000-01-01 00:00:00", "proleptic_gregorian") == _epoch(2020, 1, 1, 12)
    assert nc3io.num2epoch(90, "minutes since 2001-02-03T04:05:06Z", "standard") == _epoch(2001, 2, 3, 5, 35, 6)
    assert nc3io.num2epoch(0, "seconds since 2000-01-01 00:00:00 +01:00", "standard") == _epoch(1999, 12, 31, 23)
    # microseconds survive, finer parts are rounded away (num2date's resolution)
    assert nc3io.num2epoch(10.25, "seconds since 1970-01-01 00:00:00", "standard") == 10.25
    assert nc3io.num2epoch(10.2500004, "seconds since 1970-01-01 00:00:00", "standard") == 10.25
    # model calendars: the date is counted in the file's calendar, then read as a real date
    assert nc3io.num2epoch(59, "days since 2000-01-01", "noleap") == _epoch(2000, 3, 1)       # no 29 Feb 2000
    assert nc3io.num2epoch(59, "days since 2000-01-01", "standard") == _epoch(2000, 2, 29)
    with pytest.raises(ValueError):                                                         # 2001-02-29 is not a real date
        nc3io.num2epoch(59, "days since 2001-01-01", "366_day")
    assert nc3io.num2epoch(365 * 20 + 31, "days since 1980-01-01", "365_day") == _epoch(2000, 2, 1)
    assert nc3io.num2epoch(360 * 3 + 45.5, "days since 2000-01-01", "360_day") == _epoch(2003, 2, 16, 12)
    with pytest.raises(ValueError):
        nc3io.num2epoch(1, "fortnights since 2000-01-01", "standard")
    with pytest.raises(ValueError):
        nc3io.num2epoch(1, "days since 2000-01-01", "julian")


def test_to_epoch_time_vector_rule():
    """vet[0] from the calendar, vet[1:] = vet[0] + (X[1:] - X[0]) * seconds per unit; minutes are
    refused for vectors as in the reference (gonzag/ncio.py:47-55)."""
    x = np.array([613608.0, 613609.0, 613611.5])
    v = nc3io.to_epoch_time(x, "hours since 1950-01-01 00:00:00", "gregorian")
    t0 = _epoch(2020, 1, 1)
    assert v.tolist() == [t0, t0 + 3600., t0 + 3.5 * 3600.]
    v = nc3io.to_epoch_time(np.ma.masked_array([1.0, 3.0]), "days since 1970-01-01", "standard")
    assert v.tolist() == [86400., 3 * 86400.]
    with pytest.raises(ValueError):
        nc3io.to_epoch_time(x, "minutes since 1950-01-01 00:00:00", "gregorian")


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_file_readers_equal_memory_sources(golden_grid, tmp_path, tag):
    g = golden_grid
    model, track, lsm, varlsm, dist, Np = grid_case_sources(g, tag)
    fm, ft, fl, fvarlsm, fdist, fNp = write_grid_case_files(g, tag, tmp_path)
    assert (fvarlsm, fdist, fNp) == (varlsm, dist, Np)
    assert nc3io.is_netcdf3(fm) and nc3io.is_netcdf3(ft) and not nc3io.is_netcdf3(str(tmp_path / "nothing.nc"))
    M, T, Lm = nc3io.NetCDF3Source(fm), nc3io.NetCDF3Source(ft), nc3io.NetCDF3Source(fl)
    # times: the first value of every read goes through num2date, i.e. is rounded to the microsecond
    # (gonzag/ncio.py:40-46); the in-memory sources hand the epoch seconds through untouched
    for F, S in ((M, model), (T, track)):
        (n1, (a1, b1)), (n2, (a2, b2)) = F.GetTimeInfo(), S.GetTimeInfo()
        assert n1 == n2 and abs(a1 - a2) <= US and abs(b1 - b2) <= US
    for kw in (dict(), dict(kt1=1, kt2=3), dict(kt1=0, kt2=3), dict(isubsamp=2)):
        np.testing.assert_allclose(M.GetTimeEpochVector(**kw), model.GetTimeEpochVector(**kw), rtol=0, atol=US)
    np.testing.assert_allclose(T.GetTimeEpochVector(kt1=500, kt2=4000, isubsamp=500),
                               track.GetTimeEpochVector(kt1=500, kt2=4000, isubsamp=500), rtol=0, atol=US)
    for what in ("latitude", "longitude"):
        a, b = M.GetModelCoor(what), model.GetModelCoor(what)
        assert np.ma.getdata(a).dtype == np.asarray(b).dtype and np.ma.getdata(a).dtype.isnative
        np.testing.assert_array_equal(np.ma.getdata(a), b)
        np.testing.assert_array_equal(np.ma.getdata(T.GetSatCoor(what, 5, 900)), track.GetSatCoor(what, 5, 900))
        np.testing.assert_array_equal(np.ma.getdata(T.GetSatCoor(what)), track.GetSatCoor(what))
    got = Lm.GetModelLSM(varlsm)
    assert got.dtype == int
    np.testing.assert_array_equal(got, lsm.GetModelLSM(varlsm))
    for k in (0, len(model.time) - 1):
        a, b = M.GetModel2DVar("ssh", kt=k), model.GetModel2DVar("ssh", kt=k)
        assert np.ma.getdata(a).dtype == np.ma.getdata(b).dtype and np.ma.getdata(a).dtype.isnative
        ocean = ~np.ma.getmaskarray(b)
        np.testing.assert_array_equal(np.ma.getmaskarray(a), np.ma.getmaskarray(b))
        np.testing.assert_array_equal(np.ma.getdata(a)[ocean], np.ma.getdata(b)[ocean])
    keep = (np.array([0, 3, 4, 10]),)
    np.testing.assert_array_equal(T.GetSatSSH("ssh", kt1=2, kt2=40, ikeep=keep), track.GetSatSSH("ssh", kt1=2, kt2=40, ikeep=keep))
    assert M.field_array("ssh") is None
    with pytest.raises(ValueError):
        M.GetModelCoor("depth")
    with pytest.raises(ValueError):
        T.GetSatSSH("ssh", kt1=5, kt2=5)
    for s in (M, T, Lm):
        s.close()


def test_fill_values_scaling_and_default_fill(tmp_path):
    from scipy.io import netcdf_file
    p = str(tmp_path / "packed.nc")
    with netcdf_file(p, "w") as f:
        f.createDimension("time", None)
        f.createDimension("y", 2)
        f.createDimension("x", 3)
        t = f.createVariable("time", "f8", ("time",))
        t.units = "days since 2000-01-01 00:00:00"
        t.calendar = "noleap"
        t[:] = [59.0, 60.0]
        v = f.createVariable("sossheig", "i2", ("time", "y", "x"))
        v.scale_factor = np.float32(0.001)
        v.add_offset = np.float32(1.0)
        v.missing_value = np.int16(-32000)
        v[:] = np.array([[[0, 1000, -32000], [5, 6, 7]], [[1, 2, 3], [4, 5, -32000]]], dtype=np.int16)
        w = f.createVariable("sla", "f4", ("time",))          # no fill attribute: the library default applies
        w[:] = np.array([0.5, 9.969209968386869e+36], dtype=np.float32)
    S = nc3io.NetCDF3Source(p)
    x = S.GetModel2DVar("sossheig", kt=0)
    assert np.ma.getmaskarray(x).tolist() == [[False, False, True], [False, False, False]]
    np.testing.assert_allclose(np.ma.getdata(x)[0, :2], [1.0, 2.0], rtol=1e-6)
    assert S.GetModelLSM("_FillValue@sossheig").tolist() == [[1, 1, 0], [1, 1, 1]]
    assert S.GetSatSSH("sla").tolist() == [0.5, -9999.0]
    t = S.GetTimeEpochVector()
    assert t.tolist() == [_epoch(2000, 3, 1), _epoch(2000, 3, 1) + 86400.]
    S.close()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_overlap_and_sattrack_from_file_names(golden_grid, tmp_path, tag):
    """`GetEpochTimeOverlap(file_sat, file_mod)` and `SatTrack(file_sat, ...)` given FILE NAMES, as
    alongtrack_sat_vs_nemo.py:58-71 calls them, against the unmodified reference's outputs."""
    g = golden_grid
    fm, ft, fl, varlsm, dist, Np = write_grid_case_files(g, tag, tmp_path)
    (it1, it2), (nts, ntm) = gz.GetEpochTimeOverlap(ft, fm)
    np.testing.assert_allclose([it1, it2], g[tag + "_it"], rtol=0, atol=US)      # microsecond rounding of num2date
    assert [nts, ntm] == g[tag + "_nts_ntm"].tolist()
    ST = gz.SatTrack(ft, it1, it2, Np=Np, domain_bounds=g[tag + "_mg_bounds"].tolist(), l_0_360=bool(g[tag + "_mg_scal"][3]))
    assert ST.file == ft
    assert [ST.jt1, ST.jt2, ST.size] == g[tag + "_st_j"].tolist()
    np.testing.assert_array_equal(ST.keepit[0], g[tag + "_st_keepit"])
    np.testing.assert_allclose(ST.time, g[tag + "_st_time"], rtol=0, atol=US)
    np.testing.assert_array_equal(ST.lat, g[tag + "_st_lat"])
    np.testing.assert_array_equal(ST.lon, g[tag + "_st_lon"])
    ssh = ST.source.GetSatSSH("ssh", kt1=ST.jt1, kt2=ST.jt2, ikeep=ST.keepit)
    np.testing.assert_array_equal(ssh, g[tag + "_ssat"][ST.jt1:ST.jt2 + 1][ST.keepit] if ST.jt1 > 0 and ST.jt2 > 0
                                  else g[tag + "_ssat"][ST.keepit])


def test_dimension_variants(tmp_path):
    """The shapes gonzag/ncio.py accepts: 3-D coordinates (first record taken, :156), 4-D fields (surface level,
    :204), 3-D masks (:187), mask from a 4-D field's fill value (:182), `TIME` as the record name (:77)."""
    from scipy.io import netcdf_file
    p = str(tmp_path / "nemo_like.nc")
    lat2 = np.arange(6.0).reshape(2, 3)
    fld = np.arange(2 * 2 * 2 * 3, dtype=np.float32).reshape(2, 2, 2, 3)
    fld[:, :, 1, 2] = 1.e20
    with netcdf_file(p, "w") as f:
        f.createDimension("TIME", None)
        f.createDimension("deptht", 2)
        f.createDimension("y", 2)
        f.createDimension("x", 3)
        t = f.createVariable("TIME", "f8", ("TIME",))
        t.units = "hours since 2000-01-01 00:00:00"
        t.calendar = "gregorian"
        t[:] = [0.0, 6.0]
        v = f.createVariable("nav_lat", "f4", ("TIME", "y", "x"))
        v[0] = lat2
        v[1] = lat2 + 100.0
        v = f.createVariable("nav_lon", "f4", ("TIME", "y", "x"))
        v[0] = lat2 * 2
        v[1] = 0.0
        v = f.createVariable("thetao", "f4", ("TIME", "deptht", "y", "x"))
        v._FillValue = np.float32(1.e20)
        v[:] = fld
        v = f.createVariable("tmask3", "i1", ("TIME", "y", "x"))
        v[0] = [[1, 1, 0], [1, 0, 1]]
        v[1] = 0
    S = nc3io.NetCDF3Source(p)
    assert S.GetTimeInfo() == (2, (_epoch(2000, 1, 1), _epoch(2000, 1, 1, 6)))
    np.testing.assert_array_equal(np.ma.getdata(S.GetModelCoor("latitude")), lat2.astype(np.float32))
    np.testing.assert_array_equal(np.ma.getdata(S.GetModelCoor("longitude")), (lat2 * 2).astype(np.float32))
    x = S.GetModel2DVar("thetao", kt=1)
    assert x.shape == (2, 3) and np.ma.getmaskarray(x).tolist() == [[False, False, False], [False, False, True]]
    np.testing.assert_array_equal(np.ma.getdata(x)[0], fld[1, 0, 0])
    assert S.GetModelLSM("_FillValue@thetao").tolist() == [[1, 1, 1], [1, 1, 0]]
    assert S.GetModelLSM("tmask3").tolist() == [[1, 1, 0], [1, 0, 1]]
    with pytest.raises(ValueError):
        S.GetModelLSM("_FillValue@TIME")                    # a 1-D variable cannot give a mask
    S.close()


def test_record_into_equals_the_reader(tmp_path):
    """`NetCDF3Source.record_into` (one threaded byte-swapping pass into a staging slot) against the data of
    `GetModel2DVar` -- fill values included --, for 3-D and 4-D variables of several types; variables that need
    scaling, and slots of another dtype / shape / layout, are declined."""
    from scipy.io import netcdf_file
    from gonzag_b200.nc3io import NetCDF3Source
    from gonzag_b200.mod2sat import _record_writer
    import types
    rng = np.random.default_rng(3)
    fn = str(tmp_path / "rec.nc")
    nj, ni, nt = 37, 53, 4
    with netcdf_file(fn, "w", version=2) as f:
        f.createDimension("time_counter", None)
        f.createDimension("deptht", 2)
        f.createDimension("y", nj)
        f.createDimension("x", ni)
        v = f.createVariable("time_counter", "f8", ("time_counter",))
        v.units = "seconds since 1970-01-01 00:00:00"
        v[:] = 3600. * np.arange(nt)
        a = f.createVariable("ssh", "f4", ("time_counter", "y", "x"))
        a._FillValue = np.float32(1.e20)
        b = f.createVariable("ssh8", "f8", ("time_counter", "y", "x"))
        c = f.createVariable("temp", "f4", ("time_counter", "deptht", "y", "x"))
        d = f.createVariable("packed", "i2", ("time_counter", "y", "x"))
        d.scale_factor = 0.001
        e = f.createVariable("counts", "i2", ("time_counter", "y", "x"))
        for k in range(nt):
            x = rng.standard_normal((nj, ni)).astype(np.float32)
            x[::5, ::3] = 1.e20
            a[k] = x
            b[k] = rng.standard_normal((nj, ni))
            c[k] = rng.standard_normal((2, nj, ni)).astype(np.float32)
            d[k] = rng.integers(-3000, 3000, (nj, ni)).astype(np.int16)
            e[k] = rng.integers(-3000, 3000, (nj, ni)).astype(np.int16)
    src = NetCDF3Source(fn)
    try:
        for name, dt in (("ssh", np.float32), ("ssh8", np.float64), ("temp", np.float32), ("counts", np.int16)):
            stage = np.full((3, nj, ni), 7, dtype=dt)
            for k in (0, 2, 3):
                want = np.ma.getdata(src.GetModel2DVar(name, kt=k))
                assert want.dtype == dt
                assert src.record_into(name, k, stage[1]) is True
                assert stage[1].tobytes() == want.tobytes(), (name, k)
                assert (stage[0] == 7).all() and (stage[2] == 7).all()
        slot = np.zeros((nj, ni), np.float64)
        assert src.record_into("packed", 1, slot) is False and not slot.any()           # needs scaling: ordinary reader
        assert src.record_into("ssh", 1, slot) is False                                    # other dtype
        assert src.record_into("ssh", 1, np.zeros((nj, ni + 1), np.float32)) is False      # other shape
        assert src.record_into("ssh", 1, np.zeros((ni, nj), np.float32).T) is False        # not C-contiguous
        assert src.record_into("ssh", 1, np.zeros((nj, ni), ">f4")) is False               # not native
        with pytest.raises(ValueError):
            src.record_into("time_counter", 0, slot)
        # who gets the fast path: a model grid bound to a NetCDF-3 file, not one carrying arrays / callables
        assert _record_writer(types.SimpleNamespace(source=src)) is not None
        assert _record_writer(types.SimpleNamespace(source=src, field=np.zeros((2, 3, 3)))) is None
        assert _record_writer(types.SimpleNamespace(source=object())) is None
        assert _record_writer(types.SimpleNamespace()) is None
    finally:
        src.close()
