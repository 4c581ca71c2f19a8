"""GPU fuzz of the mapping (K1-K3) against the CPU oracle on small odd cases: the generator of
tests/test_oracle_vs_reference.py::test_fuzz_small_odd_cases (grid kind, overlap convention, radius, box size,
drifting targets with outliers and exact node hits), on which the oracle was held to the unmodified reference.
    python scripts/gpu_fuzz.py [first seed] [count]
Indices must be identical, weights within 1e-6 relative (+1e-9).  Each case then goes through the whole of
`Model2SatTrack` with the fields of `_fuzz_fields` (NaN patches, fill values, track times exactly on a record):
masks identical, values within 1e-6 relative.  Prints every differing seed; exit code 1 if any."""
import types
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gonzag_b200 as gzb
from oracle import gz_oracle as orc
from tests.test_oracle_vs_reference import _fuzz_case, _fuzz_fields

first = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 400
bad = compared = 0
for seed in range(first, first + count):
    lat, lon, y, x, kew, rd, r, use_angle = _fuzz_case(seed)
    ang = orc.grid_angle(lat, lon) if use_angle else []
    try:
        want = orc.TrackMapping(y, x, lat, lon, src_grid_local_angle=ang, k_ew_per=kew, rd_found_km=rd, np_box_r=r)
    except ValueError:
        continue          # the reference (and the oracle) slice an empty box: undefined there
    got = gzb.BilinTrack(y, x, lat, lon, src_grid_local_angle=ang, k_ew_per=kew, rd_found_km=rd, np_box_r=r)
    compared += 1
    ok_np = np.array_equal(got.NP, want.NP)
    ok_sm = np.array_equal(got.SM, want.SM)
    ok_wb = np.allclose(got.WB, want.WB, rtol=1e-6, atol=1e-9, equal_nan=True)
    if not (ok_np and ok_sm and ok_wb):
        bad += 1
        rows = np.where((got.NP != want.NP).any(axis=1) | (got.SM != want.SM).reshape(len(y), -1).any(axis=1))[0]
        print("seed %d  grid %s  kew %d  r %d  rd %.1f  angle %s:  NP %s  SM %s  WB %s  first rows %s" % (
            seed, lat.shape, kew, r, rd, use_angle, ok_np, ok_sm, ok_wb, rows[:5].tolist()), flush=True)
    got.close()
    # ---- the whole of Model2SatTrack on the same case
    res, tmod, t, fld, mask, ssat = _fuzz_fields(seed, lat, len(y))
    angf = ang if use_angle else np.zeros(lat.shape)
    with np.errstate(all="ignore"):
        P = orc.track_products(lat, lon, mask, tmod, lambda k: fld[k], t, y, x, ssat.copy(), res, angle=angf, k_ew_per=kew)
    MG = types.SimpleNamespace(shape=lat.shape, HResDeg=res, lat=lat, lon=lon, xangle=angf, EWPer=kew, time=tmod, file="mem",
                               mask=mask, field=fld)
    ST = types.SimpleNamespace(size=len(y), lat=y, lon=x, time=t, file="mem", jt1=0, jt2=0, keepit=[], ssh=ssat.copy())
    R = gzb.Model2SatTrack(MG, "ssh", ST, "ssh")
    okm = np.array_equal(R.mask, P["mask"]) and np.allclose(R.distance, P["distance"], rtol=1e-12, atol=1e-12)
    for name in ("ssh_mod_np", "ssh_mod", "ssh_sat", "XNPtrack"):
        a, b = getattr(R, name), P[name]
        okm = okm and np.array_equal(np.ma.getmaskarray(a), np.ma.getmaskarray(b))
        okm = okm and np.allclose(np.ma.getdata(a), np.ma.getdata(b), rtol=1e-6, atol=1e-9, equal_nan=True)
    if not okm:
        bad += 1
        print("seed %d  Model2SatTrack products differ (grid %s, kew %d, field %s)" % (seed, lat.shape, kew, fld.dtype), flush=True)
print("compared %d cases, %d differ" % (compared, bad))
sys.exit(1 if bad else 0)
