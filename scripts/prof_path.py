"""One gzb_map_interp call per repetition on a bench-shaped workload (for an ncu launch list of a whole step):
    python scripts/prof_path.py [C3|C3saral|C5w] [reps]
C3: 10-day Jason-like track; C3saral: 10-day SARAL-like (polar) track; C5w: three missions (Jason-, SARAL-, S3-like), 5 days
each, as ONE multi-chain call.  Two model records bracket the whole track (K4's access pattern is not the point here)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gonzag_b200 as gzb
from gonzag_b200 import synthetic as syn
from gonzag_b200.multi import CudaEngine

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lat, lon = syn.orca_like_grid(3059, 4322)
mask = syn.land_mask(lat, lon)
res = syn.grid_resolution_deg(lon)
rd, r = syn.search_params(res)
if wl == "C5w":
    parts = [syn.orbit_track(5 * 86400, t0=100.0, incl_deg=a, period_s=b, phase=c, lon0_deg=d)
             for (a, b, c, d) in [(66.04, 6745.7, 0.37, 11.3), (98.55, 6035.0, 1.91, 140.2), (98.65, 6059.0, 4.20, 251.7)]]
    starts = np.cumsum([0] + [len(p[0]) for p in parts[:-1]])
    t, y, x = (np.concatenate([p[q] for p in parts]) for q in range(3))
else:
    t, y, x = syn.orbit_track(10 * 86400, t0=100.0, **(syn.SARAL if wl == "C3saral" else syn.JASON))
    starts = np.array([0])
torch.cuda.set_device(0)
ctx = gzb.GridContext(lat, lon, mask=mask, k_ew_per=2)
ctx.grid_angle(out=False)
ctx.set_option("heading_table", 2)
eng = CudaEngine(ctx, torch)
yd, xd, td = eng.tensor(y), eng.tensor(x), eng.tensor(t)
tmod = eng.tensor(np.array([0.0, t.max() + 1000.0]))
slab = eng.tensor(syn.ssh_records(lat, lon, 2))
n = len(y)
NP, SM, WB = eng.empty((n, 2), torch.int64), eng.empty((n, 4, 2), torch.int64), eng.empty((n, 4), torch.float64)
vnp, vbl = eng.empty((n,), torch.float64), eng.empty((n,), torch.float64)
seeds = np.array([[r, r]] * len(starts))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(reps):
    torch.cuda.synchronize()
    e0.record()
    eng.map_interp_multi(starts, seeds, yd, xd, td, rd, r, tmod, slab, 0, NP, SM, WB, vnp, vbl)
    e1.record()
    torch.cuda.synchronize()
    print("call %d: %.1f us for %d points" % (it, 1e3 * e0.elapsed_time(e1), n))
st = ctx.stats()
print({k: st[k] for k in ("last_breaks", "last_chain_steps", "last_unresolved", "last_unres_nocert", "last_unres_noconv", "last_unres_noseed", "last_max_chain")})
