"""GPU experiment: K1 with the per-thread search vs the warp-cooperative search on a bench workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gonzag_b200 as gzb
from gonzag_b200 import synthetic as syn
from gonzag_b200 import _lib as L
from gonzag_b200.multi import CudaEngine

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
cfg = {"C1": (292, 362, 1.0, syn.SARAL), "C3": (3059, 4322, 10.0, syn.JASON), "C3saral": (3059, 4322, 10.0, syn.SARAL)}[wl]
lat, lon = syn.orca_like_grid(cfg[0], cfg[1])
res = syn.grid_resolution_deg(lon)
rd, r = syn.search_params(res)
t, y, x = syn.orbit_track(int(cfg[2] * 86400), t0=100.0, **cfg[3])
torch.cuda.set_device(0)
t0 = time.perf_counter()
ctx = gzb.GridContext(lat, lon, k_ew_per=2)
ctx.sync()
print("context %.1f ms, stats %s" % ((time.perf_counter() - t0) * 1e3, ctx.stats()))
eng = CudaEngine(ctx, torch)
yd, xd = eng.tensor(y), eng.tensor(x)
out = {}
for path in (1, 0):
    ctx.set_option("nearest_path", path)
    NP = eng.empty((len(y), 2), torch.int64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for it in range(6):
        torch.cuda.synchronize()
        e0.record()
        eng.nearest(yd, xd, rd, r, (r, r), NP)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    st = ctx.stats()
    out[path] = NP.cpu().numpy()
    print("path %d: K1 %.3f ms (min %.3f)  unresolved %d / %d  breaks %d steps %d" % (
        path, np.mean(ms[1:]), np.min(ms[1:]), st["last_unresolved"], len(y), st["last_breaks"], st["last_chain_steps"]))
print("identical:", np.array_equal(out[0], out[1]))
