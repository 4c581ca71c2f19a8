"""GPU experiment: repeated end-to-end Model2SatTrack calls on the bench workload, per-call breakdown."""
import sys, os, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gonzag_b200 as gzb
from gonzag_b200 import _lib as L
from gonzag_b200 import synthetic as syn
import bench

w = bench.WORKLOADS["C3"]
lat, lon, mask, res, rd, r = bench.make_grid(w)
t, y, x = bench.make_track(w, 0, 1)
k_hi = int(t[-1] // 3600.0) + 1
nrec = k_hi + 1
host_field = L.pinned_empty((nrec, w["nj"], w["ni"]), np.float32)
rec = syn.ssh_record(lat, lon, 0).astype(np.float32)
for k in range(nrec):
    host_field[k] = rec + 0.01 * k
with gzb.GridContext(lat, lon) as c:
    ang = c.grid_angle()
MG = types.SimpleNamespace(shape=lat.shape, HResDeg=res, lat=lat, lon=lon, xangle=ang, EWPer=2,
                           time=3600.0 * np.arange(nrec), file="mem", mask=mask, field=host_field)
ST = types.SimpleNamespace(size=len(t), lat=y, lon=x, time=t, file="mem", jt1=0, jt2=0, keepit=[], ssh=np.zeros(len(t)))
for it in range(8):
    t0 = time.perf_counter()
    R = gzb.Model2SatTrack(MG, "ssh", ST, "ssh", keep_mapping=False)
    dt = time.perf_counter() - t0
    print("call %d: %.1f ms  %s" % (it, 1e3 * dt, {k: round(1e3 * v, 1) for k, v in R.timing.items()}), flush=True)
