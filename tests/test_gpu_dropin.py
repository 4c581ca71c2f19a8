"""The sequence alongtrack_sat_vs_nemo.py runs -- GetEpochTimeOverlap, ModGrid, SatTrack,
Model2SatTrack -- through the drop-in classes on in-memory sources, against the outputs of the
UNMODIFIED reference on the same data (tests/golden/golden_grid.npz, case_grid of make_golden.py):
global periodic grid with -D, regional grid across Greenwich, regional grid from 1-D coordinates
with the mask taken from _FillValue."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gonzag_b200 as gz
from gonzag_b200 import _lib as L
from gonzag_b200.sources import ModelArrays
from tests.conftest import grid_case_sources

RTOL, ATOL = 1e-6, 1e-9


def _pipeline(g, tag, **kw):
    model, track, lsm, varlsm, dist, Np = grid_case_sources(g, tag)
    (it1, it2), (nts, ntm) = gz.GetEpochTimeOverlap(track, model)
    MG = gz.ModGrid(model, it1, it2, lsm, varlsm, distorded_grid=dist)
    ST = gz.SatTrack(track, it1, it2, Np=Np, domain_bounds=MG.domain_bounds, l_0_360=MG.l360)
    R = gz.Model2SatTrack(MG, "ssh", ST, "ssh", **kw)
    return MG, ST, R


def _check_results(R, g, tag):
    np.testing.assert_array_equal(R.mask, g[tag + "_r_mask"])
    for nm in ("ssh_mod_np", "ssh_mod", "ssh_sat"):
        got = getattr(R, nm)
        np.testing.assert_array_equal(np.ma.getmaskarray(got), g[tag + "_r_" + nm + "_mask"])
        np.testing.assert_allclose(np.ma.getdata(got), g[tag + "_r_" + nm + "_data"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(R.distance, g[tag + "_r_dist"], rtol=1e-9)
    np.testing.assert_array_equal(np.ma.getmaskarray(R.XNPtrack), g[tag + "_r_XNPtrack_mask"])
    np.testing.assert_array_equal(np.ma.getdata(R.XNPtrack), g[tag + "_r_XNPtrack_data"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cli_sequence_matches_reference(golden_grid, tag):
    g = golden_grid
    MG, ST, R = _pipeline(g, tag)
    try:
        # ---- ModGrid
        assert MG.size == int(g[tag + "_mg_size"]) and MG.shape == g[tag + "_mg_lat"].shape
        np.testing.assert_array_equal(MG.time, g[tag + "_mg_time"])
        np.testing.assert_array_equal(MG.lat, g[tag + "_mg_lat"])
        np.testing.assert_array_equal(MG.lon, g[tag + "_mg_lon"])
        np.testing.assert_array_equal(MG.mask, g[tag + "_mg_mask"])
        res, reskm, lglob, l360, ew = g[tag + "_mg_scal"]
        assert (MG.HResDeg, MG.HResKM, bool(MG.IsLonGlobal), bool(MG.l360), MG.EWPer) == (res, reskm, bool(lglob), bool(l360), int(ew))
        assert MG.domain_bounds == g[tag + "_mg_bounds"].tolist()
        np.testing.assert_allclose(MG.xangle, g[tag + "_mg_xangle"], rtol=0, atol=1e-9)
        # the device reductions agree with the host function the reference exports
        assert gz.IsGlobalLongitudeWise(MG.lon if MG.l360 else np.mod(MG.lon, 360.), resd=MG.HResDeg)[:2] == (MG.IsLonGlobal, MG.l360)
        # ---- SatTrack (also covered without a GPU) and the products
        assert [ST.jt1, ST.jt2, ST.size] == g[tag + "_st_j"].tolist()
        _check_results(R, g, tag)
        assert R.fused_path
        # ---- another variable on the same track: the mapping is reused, only K4 runs
        n0 = MG.ctx.stats()["kernel_launches"]
        R2 = gz.Model2SatTrack(MG, "ssh", ST, "ssh", mapping=R.BT)
        _check_results(R2, g, tag)
        assert MG.ctx.stats()["kernel_launches"] - n0 < 12
    finally:
        MG.close()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_cli_sequence_from_netcdf3_files(golden_grid, tmp_path, tag):
    """The same sequence given FILE NAMES, as alongtrack_sat_vs_nemo.py:58-74 does: NetCDF-3 files read by
    the built-in readers (gonzag_b200/nc3io.py), model records streamed one at a time through pinned
    memory into K4.  Times pass through num2date's microsecond rounding (gonzag/ncio.py:40-46), hence
    the 1.5e-6 s tolerance on them; indices and masks stay bit-exact."""
    from tests.conftest import write_grid_case_files
    g = golden_grid
    fm, ft, fl, varlsm, dist, Np = write_grid_case_files(g, tag, tmp_path)
    (it1, it2), (nts, ntm) = gz.GetEpochTimeOverlap(ft, fm)
    MG = gz.ModGrid(fm, it1, it2, fl, varlsm, distorded_grid=dist)
    try:
        ST = gz.SatTrack(ft, it1, it2, Np=Np, domain_bounds=MG.domain_bounds, l_0_360=MG.l360)
        assert MG.file == fm and ST.file == ft
        assert MG.size == int(g[tag + "_mg_size"])
        np.testing.assert_allclose(MG.time, g[tag + "_mg_time"], rtol=0, atol=1.5e-6)
        np.testing.assert_array_equal(MG.lat, g[tag + "_mg_lat"])
        np.testing.assert_array_equal(MG.lon, g[tag + "_mg_lon"])
        np.testing.assert_array_equal(MG.mask, g[tag + "_mg_mask"])
        assert MG.domain_bounds == g[tag + "_mg_bounds"].tolist()
        assert [ST.jt1, ST.jt2, ST.size] == g[tag + "_st_j"].tolist()
        R = gz.Model2SatTrack(MG, "ssh", ST, "ssh")
        assert not R.fused_path                       # per-record provider: gzb_mapping + slabs of gzb_interp
        _check_results(R, g, tag)
        path = str(tmp_path / "result.nc")            # and out again through the writers, read back by the reader
        gz.SaveTimeSeries(R.time, np.ma.array([R.lat, R.lon, R.ssh_mod, R.ssh_sat]), ["latitude", "longitude", "ssh_mod", "ssh_sat"],
                          path, time_units="seconds since 1970-01-01 00:00:00", missing_val=gz.rmissval)
        back = gz.NetCDF3Source(path)
        np.testing.assert_allclose(back.GetTimeEpochVector(), R.time, rtol=0, atol=1.5e-6)
        np.testing.assert_array_equal(back.GetSatSSH("ssh_mod"), np.ma.filled(R.ssh_mod, gz.rmissval).astype(np.float32))
        back.close()
    finally:
        MG.close()


@pytest.mark.parametrize("tag", ["a", "c"])
def test_cli_main_from_files(golden_grid, tmp_path, tag):
    """`python -m gonzag_b200.cli` (flags and steps of alongtrack_sat_vs_nemo.py) from NetCDF-3 files to
    result.nc / xnp_msk.nc: -D with a mesh-mask file (a), `-l 0` = mask from _FillValue (c)."""
    from scipy.io import netcdf_file
    from gonzag_b200 import cli
    from tests.conftest import write_grid_case_files
    g = golden_grid
    fm, ft, fl, varlsm, dist, Np = write_grid_case_files(g, tag, tmp_path)
    out = tmp_path / "out"
    argv = ["-s", ft, "-n", "ssh", "-m", fm, "-v", "ssh", "-o", str(out)]
    if tag == "a":
        argv += ["-l", fl, "-k", "tmask", "-D"]
    R = cli.main(argv)
    _check_results(R, g, tag)
    back = gz.NetCDF3Source(str(out / "result.nc"))
    assert back.GetTimeInfo()[0] == len(R.time)
    np.testing.assert_array_equal(back.GetSatSSH("ssh_bl"), np.ma.getdata(R.ssh_mod).astype(np.float32))
    np.testing.assert_array_equal(back.GetSatSSH("ssh_np"), np.ma.getdata(R.ssh_mod_np).astype(np.float32))
    np.testing.assert_array_equal(back.GetSatSSH("distance"), R.distance.astype(np.float32))
    np.testing.assert_array_equal(np.ma.getdata(back.GetSatCoor("latitude")), R.lat.astype(np.float32))
    back.close()
    with netcdf_file(str(out / "xnp_msk.nc"), "r", mmap=False) as f:
        trk = f.variables["track"][:]
        assert trk.shape == R.XNPtrack.shape and f.variables["nav_lon"].shape == trk.shape
        np.testing.assert_array_equal(np.isnan(trk), np.ma.getmaskarray(R.XNPtrack))


@pytest.mark.parametrize("provider", ["array", "callable", "netcdf3"])
def test_model_window_not_starting_at_record_0(golden_window, tmp_path, provider):
    """The track starts two model records into the model file (jt1 = 2).  The reference keeps MG.time = time[2:6] and
    asks its reader for record k = index into THAT vector (gonzag/mod2sat.py:94-112): every provider has to follow the
    same rule -- a whole in-memory array (more records than MG.time), a per-record callable, a NetCDF-3 file."""
    from tests.conftest import write_grid_case_files
    g = golden_window
    if provider == "netcdf3":
        fm, ft, fl, varlsm, dist, Np = write_grid_case_files(g, "w", tmp_path)
        model, track, lsm = fm, ft, fl
    else:
        model, track, lsm, varlsm, dist, Np = grid_case_sources(g, "w")
        if provider == "callable":
            arr = model.fields["ssh"]
            model.fields["ssh"] = lambda k: arr[k]
    (it1, it2), _ = gz.GetEpochTimeOverlap(track, model)
    MG = gz.ModGrid(model, it1, it2, lsm, varlsm, distorded_grid=dist)
    try:
        assert MG.jt1 == 2 and MG.size == int(g["w_mg_size"])
        np.testing.assert_allclose(MG.time, g["w_mg_time"], rtol=0, atol=1.5e-6)
        ST = gz.SatTrack(track, it1, it2, Np=Np, domain_bounds=MG.domain_bounds, l_0_360=MG.l360)
        assert [ST.jt1, ST.jt2, ST.size] == g["w_st_j"].tolist()
        R = gz.Model2SatTrack(MG, "ssh", ST, "ssh")
        assert R.fused_path == (provider == "array")
        assert type(R.mask) is np.ndarray and R.mask.dtype == np.int8
        _check_results(R, g, "w")
    finally:
        MG.close()


def test_streaming_provider_and_staged_paths(golden_grid):
    """Records delivered one at a time (a file-like provider) and a whole pageable array too big
    to stay resident go through gzb_mapping + gzb_interp slab by slab: same products."""
    g = golden_grid
    model, track, lsm, varlsm, dist, Np = grid_case_sources(g, "a")
    arr = model.fields["ssh"]
    calls = []

    def one_record(k):
        calls.append(k)
        return arr[k]

    streamed = ModelArrays(model.time, model.lat, model.lon, fields={"ssh": one_record})
    (it1, it2), _ = gz.GetEpochTimeOverlap(track, model)
    MG = gz.ModGrid(streamed, it1, it2, lsm, varlsm, distorded_grid=dist)
    try:
        ST = gz.SatTrack(track, it1, it2, Np=Np, domain_bounds=MG.domain_bounds, l_0_360=MG.l360)
        R = gz.Model2SatTrack(MG, "ssh", ST, "ssh", slab_bytes=3 * arr[0].nbytes)      # 3 records per slab
        assert not R.fused_path and len(calls) >= len(model.time)
        _check_results(R, g, "a")
        MG.source = model
        R = gz.Model2SatTrack(MG, "ssh", ST, "ssh", resident_bytes=0)                   # pageable, not resident
        assert not R.fused_path
        _check_results(R, g, "a")
        pin = L.pinned_empty(arr.shape, arr.dtype)
        pin[...] = arr
        MG.source = ModelArrays(model.time, model.lat, model.lon, fields={"ssh": pin})
        R = gz.Model2SatTrack(MG, "ssh", ST, "ssh", resident_bytes=0)                   # pinned + mapped host records
        _check_results(R, g, "a")
        MG.ctx.set_option("zero_copy", 2)                                               # ... gathered in place over PCIe
        R = gz.Model2SatTrack(MG, "ssh", ST, "ssh", resident_bytes=0)
        _check_results(R, g, "a")
        L.pinned_free(pin)
    finally:
        MG.close()


def test_frame_switch_on_device():
    """gzb_lon_to_pm180 == degE_to_degWE; gzb_lon_stats == the NumPy reductions, on a domain the
    reference handles in the [-180,180] frame (its own test array, tests/test_functions.py:64)."""
    import ctypes as C
    lon = np.array([[340., 358., 5.], [341., 359., 6.], [342., 359.5, 7.]])
    lat = np.array([[10., 10., 10.], [11., 11., 11.], [12., 12., 12.]])
    with gz.GridContext(lat, lon) as ctx:
        vals, cols = np.empty(4), np.empty(2, dtype=np.int64)
        ctx._ck(L.lib().gzb_lon_stats(ctx._h, L.ptr(vals), L.ptr(cols)))
        xb = gz.degE_to_degWE(lon)
        assert vals.tolist() == [lon.min(), lon.max(), xb.min(), xb.max()]
        assert cols.tolist() == [np.argmin(lon) % 3, np.argmax(lon) % 3]
        out = np.empty_like(lon)
        ctx._ck(L.lib().gzb_lon_to_pm180(ctx._h, L.ptr(out)))
        np.testing.assert_array_equal(out, xb)
        NP = ctx.nearest(np.array([11.1]), np.array([-0.8]), 200.0, 2)      # the context now lives in that frame
        assert NP.tolist() == [[1, 1]] or NP.tolist() == [[-1, -1]]


def test_multi_mission_on_one_resident_grid(golden_grid):
    """BASELINE config 5 in small: several missions mapped onto ONE model grid kept on the GPU
    (`ModGrid.ctx`): each mission gets exactly what a fresh grid context gives it."""
    import types
    from gonzag_b200 import synthetic as syn
    g = golden_grid
    model, track, lsm, varlsm, dist, Np = grid_case_sources(g, "a")
    (it1, it2), _ = gz.GetEpochTimeOverlap(track, model)
    MG = gz.ModGrid(model, it1, it2, lsm, varlsm, distorded_grid=dist)
    try:
        missions = []
        for k, orb in enumerate((syn.SARAL, syn.JASON, syn.SARAL)):
            t, y, x = syn.orbit_track(9000, t0=1500.0 + 700.0 * k, **orb)
            t = t + float(model.time[0])
            missions.append(types.SimpleNamespace(size=len(t), lat=y, lon=x, time=t, file="mem", jt1=0, jt2=0, keepit=[],
                                                  ssh=np.zeros(len(t))))
        n0 = MG.ctx.stats()["kernel_launches"]
        shared = [gz.Model2SatTrack(MG, "ssh", ST, "ssh") for ST in missions]
        assert MG.ctx.stats()["kernel_launches"] - n0 < 3 * 40       # the path only: no grid upload, no pyramid / certificate build
        for ST, R in zip(missions, shared):
            fresh = types.SimpleNamespace(shape=MG.shape, HResDeg=MG.HResDeg, lat=MG.lat, lon=MG.lon, xangle=MG.xangle,
                                          EWPer=MG.EWPer, time=MG.time, file="mem", mask=MG.mask,
                                          field=model.fields["ssh"])
            R2 = gz.Model2SatTrack(fresh, "ssh", ST, "ssh")
            np.testing.assert_array_equal(R.BT.NP, R2.BT.NP)
            np.testing.assert_array_equal(R.BT.SM, R2.BT.SM)
            np.testing.assert_array_equal(R.BT.WB, R2.BT.WB)
            np.testing.assert_array_equal(np.ma.getdata(R.ssh_mod), np.ma.getdata(R2.ssh_mod))
            np.testing.assert_array_equal(R.mask, R2.mask)
    finally:
        MG.close()
