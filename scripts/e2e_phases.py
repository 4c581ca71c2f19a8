"""GPU experiment: where do the milliseconds of an end-to-end Model2SatTrack call go once the grid is resident?
The steps of the constructor (gonzag_b200/mod2sat.py) issued one by one with a device synchronisation after each, so that
host issue time and device time of every phase can be told apart.  C3 bench workload, records in pinned host memory.
    python scripts/e2e_phases.py [reps]"""
import sys, os, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gonzag_b200 as gzb
from gonzag_b200 import _lib as L
from gonzag_b200 import synthetic as syn
from gonzag_b200 import config
from gonzag_b200.mod2sat import TrackJob, satellite_ssh
import bench

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
w = bench.WORKLOADS["C3"]
lat, lon, mask, res, rd, r = bench.make_grid(w)
t, y, x = bench.make_track(w, 0, 1)
k_hi = int(t[-1] // 3600.0) + 1
nrec = k_hi + 1
host_field = L.pinned_empty((nrec, w["nj"], w["ni"]), np.float32)
rec = syn.ssh_record(lat, lon, 0).astype(np.float32)
for k in range(nrec):
    host_field[k] = rec + 0.01 * k
tmod = 3600.0 * np.arange(nrec)
ssh = np.zeros(len(t))
with gzb.GridContext(lat, lon) as c0:
    ang = c0.grid_angle()
MG = types.SimpleNamespace(shape=lat.shape, HResDeg=res, lat=lat, lon=lon, xangle=ang, EWPer=2, time=tmod, file="mem", mask=mask, field=host_field)
ST = types.SimpleNamespace(size=len(t), lat=y, lon=x, time=t, file="mem", jt1=0, jt2=0, keepit=[], ssh=ssh)
pc = time.perf_counter
for it in range(reps):
    tm = {}
    t0 = pc()
    ctx = gzb.GridContext(lat, lon, angle=ang, mask=mask, k_ew_per=2)
    tm["ctx_create_return"] = pc() - t0
    ctx.sync(); tm["ctx_create_synced"] = pc() - t0
    for sub in range(2):          # second pass: everything warm (heading table built, workspace sized)
        p = "w%d_" % sub
        t1 = pc()
        job = TrackJob(ctx, L.as_f64(y), L.as_f64(x), L.as_f64(t), L.as_f64(tmod))
        tm[p + "job_alloc_h2d_issue"] = pc() - t1; ctx.sync(); tm[p + "job_alloc_h2d_synced"] = pc() - t1
        t1 = pc()
        job.run(MG, "ssh", rd, r, (r, r))
        tm[p + "run_issue"] = pc() - t1; ctx.sync(); tm[p + "run_synced"] = pc() - t1
        t1 = pc()
        vs = satellite_ssh(ST, "ssh")
        tm[p + "sat_host"] = pc() - t1
        t1 = pc()
        job.distance_and_mask(vs)
        tm[p + "distmask_issue"] = pc() - t1; ctx.sync(); tm[p + "distmask_synced"] = pc() - t1
        t1 = pc()
        out = job.download_series()
        tm[p + "download_issue"] = pc() - t1; ctx.sync(); tm[p + "download_synced"] = pc() - t1
        t1 = pc()
        job.free()
        tm[p + "free"] = pc() - t1
    t1 = pc(); ctx.close(); tm["close"] = pc() - t1
    print("rep %d: " % it + "  ".join("%s %.2f" % (k, 1e3 * v) for k, v in tm.items()), flush=True)
# the plain call, for comparison
for it in range(3):
    t0 = pc()
    R = gzb.Model2SatTrack(MG, "ssh", ST, "ssh", keep_mapping=False)
    print("Model2SatTrack: %.2f ms  %s" % (1e3 * (pc() - t0), {k: round(1e3 * v, 2) for k, v in R.timing.items()}), flush=True)
