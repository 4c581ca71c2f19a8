"""GPU experiment: timeline of one gzb_map_interp call (option "trace")."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gonzag_b200 as gzb
from gonzag_b200.multi import CudaEngine
import bench
w = bench.WORKLOADS["C3"]
lat, lon, mask, res, rd, r = bench.make_grid(w)
t, y, x = bench.make_track(w, 0, 1)
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
ctx = gzb.GridContext(lat, lon, mask=mask, k_ew_per=2)
ctx.grid_angle(out=False); ctx.set_option("heading_table", 2)
eng = CudaEngine(ctx, torch)
k_hi = int(t[-1] // 3600.0) + 1
slab = bench.field_on_device(torch, dev, torch.as_tensor(lat, device=dev), torch.as_tensor(lon, device=dev), 0, k_hi)
tmod = eng.tensor(3600.0 * np.arange(k_hi + 1))
yd, xd, td = eng.tensor(y), eng.tensor(x), eng.tensor(t)
n = len(y)
NP = eng.empty((n, 2), torch.int64); SM = eng.empty((n, 4, 2), torch.int64); WB = eng.empty((n, 4), torch.float64)
v_np = eng.empty((n,), torch.float64); v_bl = eng.empty((n,), torch.float64)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(4):
    eng.map_interp(yd, xd, td, rd, r, (r, r), tmod, slab, 0, NP, SM, WB, v_np, v_bl)
torch.cuda.synchronize()
for it in range(2):
    ctx.set_option("trace", 1)
    flush.zero_(); flush.zero_()
    eng.map_interp(yd, xd, td, rd, r, (r, r), tmod, slab, 0, NP, SM, WB, v_np, v_bl)
    ctx.sync()
    print("----", file=sys.stderr)
