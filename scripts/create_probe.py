"""GPU experiment: time of a grid context (upload + search structures) against the number of host copy threads.
    python scripts/create_probe.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gonzag_b200 as gzb
from gonzag_b200 import _lib as L
import bench

w = bench.WORKLOADS["C3"]
lat, lon, mask, res, rd, r = bench.make_grid(w)
with gzb.GridContext(lat, lon) as c0:
    ang = c0.grid_angle()
print("host cores:", os.cpu_count())
for n in (0, 8, 12, 16, 20, 24, 12, 24):
    L.lib().gzb_set_host_threads(n)
    ts = []
    for it in range(4):
        t0 = time.perf_counter()
        ctx = gzb.GridContext(lat, lon, angle=ang, mask=mask, k_ew_per=2)
        ts.append(time.perf_counter() - t0)
        ctx.close()
    print("host threads %2d: context %s ms" % (n, " ".join("%.2f" % (1e3 * v) for v in ts)), flush=True)
