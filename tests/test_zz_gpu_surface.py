"""CUDA side of the helpers in tests/test_surface_host.py: `Haversine`, `Haversine_sclr`, `RadiusEarth` (gonzag/utils.py:
267-316) and `IQH` (gonzag/bilin_mapping.py:267-312) through the drop-in functions, against the reference's golden
vectors.  fp64 results within 1e-12 relative (device libm vs the host's: last-place differences); the IQH cases keep
0.008 degrees between any two compared headings, so the side is an exact answer."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gonzag_b200 as gzb


def test_haversine_family(golden_units):
    g = golden_units
    hv = gzb.Haversine(np.float64(g["hav_pt"][0]), np.float64(g["hav_pt"][1]), g["hav_glat"], g["hav_glon"])
    assert hv.shape == g["hav_out"].shape and hv.dtype == np.float64
    np.testing.assert_allclose(hv, g["hav_out"], rtol=1e-12)
    a = g["hs_in"]
    d = gzb.Haversine_sclr(*[np.ascontiguousarray(a[:, k]) for k in range(4)])
    np.testing.assert_allclose(d, g["hs_out"], rtol=1e-12, atol=1e-12)
    for row, want in zip(a[:10], g["hs_out"][:10]):
        got = gzb.Haversine_sclr(*[float(v) for v in row])
        assert isinstance(got, float) and abs(got - want) <= 1e-12 * max(1.0, abs(want))
    np.testing.assert_allclose(gzb.RadiusEarth(np.ascontiguousarray(a[:, 0])), g["re_out"], rtol=1e-13)
    r = gzb.RadiusEarth(0.0)
    assert isinstance(r, float) and abs(r - 6378.137) < 1e-9
    assert abs(gzb.RadiusEarth(90.0) - 6356.752) < 1e-9
    # the along-track distance of Model2SatTrack is the running sum of these steps
    assert gzb.Haversine_sclr(10.0, 20.0, 10.0, 20.0) == 0.0


def test_iqh(golden_surface):
    g = golden_surface
    names = {1: "down", 2: "right", 3: "up", 4: "left"}
    for q in range(len(g["iqh_out"])):
        args = [g["iqh_" + k][q] for k in ("yA", "xA", "yB", "xB", "vY", "vX")]
        want = int(g["iqh_out"][q])
        if want < 0:
            with pytest.raises(SystemExit):
                gzb.IQH(*args)
        else:
            assert gzb.IQH(*args) == (want, names[want])
