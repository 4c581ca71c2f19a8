"""GPU fuzz of the mapping (K1-K3) against the CPU oracle on small odd cases: the generator of
tests/test_oracle_vs_reference.py::test_fuzz_small_odd_cases (grid kind, overlap convention, radius, box size,
drifting targets with outliers and exact node hits), on which the oracle was held to the unmodified reference.
    python scripts/gpu_fuzz.py [first seed] [count]
Indices must be identical, weights within 1e-6 relative (+1e-9).  Prints every differing seed; exit code 1 if any."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gonzag_b200 as gzb
from oracle import gz_oracle as orc
from tests.test_oracle_vs_reference import _fuzz_case

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
print("compared %d cases, %d differ" % (compared, bad))
sys.exit(1 if bad else 0)
