"""GPU fuzz: the CUDA path against the CPU oracle on the seeded small odd cases of
tests/test_oracle_vs_reference.py (`_fuzz_case` / `_fuzz_fields`: four grid kinds, every overlap convention, box
half-widths 0-5, radii 0.5-3 x the resolution rule, drifting targets with outliers and exact node hits; fields
with NaN patches, 1e20 fill values, values outside (-100, 100), track times exactly on a record) -- the cases on
which the oracle equals the unmodified reference bit for bit.  Indices and masks identical, weights and series
within 1e-6 relative (+1e-9)."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gonzag_b200 as gzb
from oracle import gz_oracle as orc
from tests.test_oracle_vs_reference import _fuzz_case, _fuzz_fields


@pytest.mark.parametrize("first,count", [(5000, 160), (7000, 120)])
def test_fuzz_mapping_and_products(first, count):
    compared = 0
    for seed in range(first, first + count):
        lat, lon, y, x, kew, rd, r, use_angle = _fuzz_case(seed)
        ang = orc.grid_angle(lat, lon) if use_angle else []
        try:
            want = orc.TrackMapping(y, x, lat, lon, src_grid_local_angle=ang, k_ew_per=kew, rd_found_km=rd, np_box_r=r)
        except ValueError:
            continue          # the reference (and the oracle) slice an empty box: undefined there
        got = gzb.BilinTrack(y, x, lat, lon, src_grid_local_angle=ang, k_ew_per=kew, rd_found_km=rd, np_box_r=r)
        try:
            tag = "seed %d grid %s kew %d r %d rd %.1f angle %s" % (seed, lat.shape, kew, r, rd, use_angle)
            np.testing.assert_array_equal(got.NP, want.NP, err_msg=tag)
            np.testing.assert_array_equal(got.SM, want.SM, err_msg=tag)
            np.testing.assert_allclose(got.WB, want.WB, rtol=1e-6, atol=1e-9, equal_nan=True, err_msg=tag)
        finally:
            got.close()
        compared += 1
        # ---- the whole of Model2SatTrack on the same case
        res, tmod, t, fld, mask, ssat = _fuzz_fields(seed, lat, len(y))
        angf = ang if use_angle else np.zeros(lat.shape)
        with np.errstate(all="ignore"):
            P = orc.track_products(lat, lon, mask, tmod, lambda k: fld[k], t, y, x, ssat.copy(), res, angle=angf, k_ew_per=kew)
        MG = types.SimpleNamespace(shape=lat.shape, HResDeg=res, lat=lat, lon=lon, xangle=angf, EWPer=kew, time=tmod,
                                   file="mem", mask=mask, field=fld)
        ST = types.SimpleNamespace(size=len(y), lat=y, lon=x, time=t, file="mem", jt1=0, jt2=0, keepit=[], ssh=ssat.copy())
        R = gzb.Model2SatTrack(MG, "ssh", ST, "ssh")
        np.testing.assert_array_equal(R.mask, P["mask"], err_msg=tag)
        np.testing.assert_allclose(R.distance, P["distance"], rtol=1e-12, atol=1e-12, err_msg=tag)
        for name in ("ssh_mod_np", "ssh_mod", "ssh_sat", "XNPtrack"):
            a, b = getattr(R, name), P[name]
            np.testing.assert_array_equal(np.ma.getmaskarray(a), np.ma.getmaskarray(b), err_msg=name + " " + tag)
            np.testing.assert_allclose(np.ma.getdata(a), np.ma.getdata(b), rtol=1e-6, atol=1e-9, equal_nan=True,
                                       err_msg=name + " " + tag)
    assert compared > 0.6 * count
