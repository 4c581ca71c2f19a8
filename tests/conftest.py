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


@pytest.fixture(scope="session")
def golden_window():
    return load_golden("window")           # case_window of make_golden.py: model window starting at record 2


@pytest.fixture(scope="session")
def golden_surface():
    return load_golden("surface")          # tests/golden/make_golden_surface.py


def grid_case_sources(g, tag):
    """The in-memory data sets of tests/golden/make_golden.py:case_grid, rebuilt from the fixture."""
    from gonzag_b200.sources import ModelArrays, TrackArrays
    if tag == "a":
        lat, lon = g["a_lat"].astype(np.float64), g["a_lon"].astype(np.float64)
        model = ModelArrays(g["a_tm"], lat, lon, fields={"ssh": g["a_f32"]})
        lsm = ModelArrays(g["a_tm"], None, None, masks={"tmask": g["a_maskin"].astype(np.int64)})
        return model, TrackArrays(g["a_t"], g["a_y"], g["a_x"], {"ssh": g["a_ssat"]}), lsm, "tmask", True, len(g["a_t"])
    if tag in ("b", "w"):
        model = ModelArrays(g[tag + "_tm"], g[tag + "_lat"], g[tag + "_lon"], fields={"ssh": g[tag + "_f64"]},
                            masks={"tmask": g[tag + "_maskin"].astype(np.int64)})
        return model, TrackArrays(g[tag + "_t"], g[tag + "_y"], g[tag + "_x"], {"ssh": g[tag + "_ssat"]}), model, "tmask", False, 100
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


def write_grid_case_files(g, tag, directory):
    """The data sets of `grid_case_sources` written as NetCDF-3 files (scipy), laid out like the files
    the reference's examples read: NEMO-style model file (+ separate mesh-mask file for case a),
    along-track satellite file.  Returns (model_path, track_path, lsm_path, varlsm, distorded, Np)."""
    import os
    from scipy.io import netcdf_file
    units = "seconds since 1970-01-01 00:00:00"
    d = str(directory)
    fm, ft, fl = (os.path.join(d, "%s_%s.nc" % (tag, n)) for n in ("model", "track", "mesh_mask"))

    def time_var(f, name, values, dim):
        f.createDimension(dim, None)
        v = f.createVariable(name, "f8", (dim,))
        v.units = units
        v.calendar = "gregorian"
        v[:] = values
        return v

    with netcdf_file(ft, "w", version=2) as f:
        time_var(f, "time", g[tag + "_t"], "time")
        for name, key in (("latitude", "_y"), ("longitude", "_x")):
            v = f.createVariable(name, "f8", ("time",))
            v[:] = g[tag + key]
        v = f.createVariable("ssh", "f8", ("time",))
        v._FillValue = np.float64(1.e20)
        v[:] = g[tag + "_ssat"]
    if tag == "a":
        with netcdf_file(fm, "w", version=2) as f:
            nj, ni = g["a_lat"].shape
            time_var(f, "time_counter", g["a_tm"], "time_counter")      # the record dimension comes first
            f.createDimension("y", nj)
            f.createDimension("x", ni)
            for name, key in (("nav_lat", "a_lat"), ("nav_lon", "a_lon")):
                v = f.createVariable(name, "f8", ("y", "x"))        # float32 values held as float64, as the
                v[:] = g[key].astype(np.float64)                     # golden generator handed them to the reference
            v = f.createVariable("ssh", "f4", ("time_counter", "y", "x"))
            v[:] = g["a_f32"]
        with netcdf_file(fl, "w", version=1) as f:
            f.createDimension("t", 1)
            f.createDimension("z", 2)
            f.createDimension("y", nj)
            f.createDimension("x", ni)
            v = f.createVariable("tmask", "i1", ("t", "z", "y", "x"))
            v[0, 0] = g["a_maskin"]
            v[0, 1] = 0
        return fm, ft, fl, "tmask", True, len(g["a_t"])
    if tag in ("b", "w"):
        with netcdf_file(fm, "w", version=2) as f:
            nj, ni = g[tag + "_lat"].shape
            time_var(f, "time", g[tag + "_tm"], "time")
            f.createDimension("y", nj)
            f.createDimension("x", ni)
            for name, key in (("gphit", tag + "_lat"), ("glamt", tag + "_lon")):
                v = f.createVariable(name, "f8", ("y", "x"))
                v[:] = g[key]
            v = f.createVariable("ssh", "f8", ("time", "y", "x"))
            v[:] = g[tag + "_f64"]
            v = f.createVariable("tmask", "i2", ("y", "x"))
            v[:] = g[tag + "_maskin"]
        return fm, ft, fm, "tmask", False, 100
    with netcdf_file(fm, "w", version=1) as f:
        nj, ni = len(g["c_lat1"]), len(g["c_lon1"])
        time_var(f, "time", g["c_tm"], "time")
        f.createDimension("lat", nj)
        f.createDimension("lon", ni)
        v = f.createVariable("lat", "f8", ("lat",))
        v[:] = g["c_lat1"]
        v = f.createVariable("lon", "f8", ("lon",))
        v[:] = g["c_lon1"]
        v = f.createVariable("ssh", "f4", ("time", "lat", "lon"))
        v._FillValue = np.float32(1.e20)
        v[:] = np.where(g["c_maskin"][None] == 0, np.float32(1.e20), g["c_f32"])
    return fm, ft, fm, "_FillValue@ssh", False, len(g["c_t"])
