"""The small helpers the reference package exports beside the hot-path functions (`from .utils import *` and
gonzag/__init__.py:13) against golden vectors of the unmodified reference (tests/golden/make_golden_surface.py):
the index bookkeeping is host code in the drop-in as it is in the reference; `IQH` is pinned here for the oracle
and in tests/test_zz_gpu_surface.py for the CUDA path."""
import numpy as np
import pytest

import gonzag_b200 as gzb
from oracle import gz_oracle as orc


def test_surrounding_j_i(golden_surface):
    g = golden_surface
    for (jP, iP, Ny, Nx, kew), want in zip(g["sji_calls"], g["sji_out"]):
        args = (int(jP), int(iP), int(Ny), int(Nx))
        assert tuple(gzb.surrounding_j_i(*args, k_ew_per=int(kew))) == tuple(want)
        assert tuple(orc.neighbours(*args, k_ew_per=int(kew))) == tuple(want)
    assert gzb.surrounding_j_i(3, 8, 6, 9) == (2, 4, 7, -1)          # default: no periodicity


def test_iquadran2srcmesh(golden_surface):
    g = golden_surface
    for (jP, iP, Ny, Nx, iqd, kew), want in zip(g["q2m_calls"], g["q2m_out"]):
        args = (int(jP), int(iP), int(Ny), int(Nx), int(iqd))
        got = gzb.Iquadran2SrcMesh(*args, k_ew_per=int(kew))
        assert [tuple(p) for p in got] == [tuple(p) for p in want]
        assert [tuple(p) for p in orc.mesh_from_quadrant(*args, k_ew_per=int(kew))] == [tuple(p) for p in want]
    # the reference exits (code 0) on a bad quadrant and on an incomplete neighbourhood
    for bad in ((3, 4, 6, 9, 5, 2), (5, 4, 6, 9, 1, 2), (3, 8, 6, 9, 1, -1)):
        with pytest.raises(SystemExit) as e:
            gzb.Iquadran2SrcMesh(*bad[:5], k_ew_per=bad[5])
        assert e.value.code == 0


def test_find_j_i(golden_surface):
    g = golden_surface
    for a, wmin, wmax in zip(g["fji_in"], g["fji_min"], g["fji_max"]):
        assert tuple(gzb.find_j_i_min(a)) == tuple(wmin)
        assert tuple(gzb.find_j_i_max(a)) == tuple(wmax)
        assert tuple(orc.argmin_ji(a)) == tuple(wmin)


def test_oracle_side_of_exit(golden_surface):
    g = golden_surface
    got = [orc.side_of_exit(*(g["iqh_" + k][q] for k in ("yA", "xA", "yB", "xB", "vY", "vX")))
           for q in range(len(g["iqh_out"]))]
    np.testing.assert_array_equal(got, g["iqh_out"])
    assert set(np.unique(g["iqh_out"])) == {-1, 3}      # what the reference's four tests can answer (see the oracle)


def test_chck4f(tmp_path):
    f = tmp_path / "there.nc"
    f.write_bytes(b"x")
    gzb.chck4f(str(f))
    with pytest.raises(SystemExit):
        gzb.chck4f(str(tmp_path / "absent.nc"))


def test_package_exports_what_the_reference_package_does():
    """Names a script written against `import gonzag as gz` can use for this path (gonzag/__init__.py:11-14)."""
    for name in ("BilinTrack", "NearestPoint", "Iquadran", "IDSourceMesh", "AlfaBeta", "WeightBL", "Heading", "IQH",
                 "Model2SatTrack", "ModGrid", "SatTrack", "MsgExit", "chck4f", "degE_to_degWE", "EpochT2Str",
                 "find_j_i_min", "find_j_i_max", "GridResolution", "IsEastWestPeriodic", "SearchBoxSize",
                 "IsGlobalLongitudeWise", "GetEpochTimeOverlap", "scan_idx", "GridAngle", "RadiusEarth",
                 "Haversine_sclr", "Haversine", "ldebug", "ivrb", "deg2km", "rfactor", "search_box_w_km",
                 "l_save_track_on_model_grid", "rmissval"):
        assert hasattr(gzb, name), name
