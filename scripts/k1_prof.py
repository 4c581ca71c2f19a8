"""K1 alone on a bench workload (for ncu): python scripts/k1_prof.py [C3] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gonzag_b200 as gzb
from gonzag_b200 import synthetic as syn
from gonzag_b200.multi import CudaEngine

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = {"C1": (292, 362, 1.0, syn.SARAL), "C3": (3059, 4322, 10.0, syn.JASON), "C3saral": (3059, 4322, 10.0, syn.SARAL)}[wl]
lat, lon = syn.orca_like_grid(cfg[0], cfg[1])
res = syn.grid_resolution_deg(lon)
rd, r = syn.search_params(res)
t, y, x = syn.orbit_track(int(cfg[2] * 86400), t0=100.0, **cfg[3])
torch.cuda.set_device(0)
ctx = gzb.GridContext(lat, lon, k_ew_per=2)
eng = CudaEngine(ctx, torch)
yd, xd = eng.tensor(y), eng.tensor(x)
NP = eng.empty((len(y), 2), torch.int64)
for it in range(reps):
    eng.nearest(yd, xd, rd, r, (r, r), NP)
torch.cuda.synchronize()
print("ok", ctx.stats())
