"""The C3 bench workload for an `ncu --set full` capture of one step and of the stand-alone stage kernels.
    ncu --profile-from-start off --set full --clock-control none --import-source on -o X python scripts/ncu_case.py
Three warm gzb_map_interp calls, then (between cudaProfilerStart / Stop) one gzb_map_interp, one K4 alone (k_interp),
one fused bulk kernel alone (k_mesh_weights_interp from known nearest points), one K2+K3 (k_mesh_weights).
Same grid, track and hourly records as bench.py's N = 1 workload (the records matter for the gather's traffic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gonzag_b200 as gzb
from gonzag_b200.multi import CudaEngine
import bench

w = bench.WORKLOADS["C3"]
lat, lon, mask, res, rd, r = bench.make_grid(w)
t, y, x = bench.make_track(w, 0, 1)
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
ctx = gzb.GridContext(lat, lon, mask=mask, k_ew_per=2)
ctx.grid_angle(out=False)
ctx.set_option("heading_table", 2)
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
eng = CudaEngine(ctx, torch)
k_hi = int(t[-1] // 3600.0) + 1
slab = bench.field_on_device(torch, dev, torch.as_tensor(lat, device=dev), torch.as_tensor(lon, device=dev), 0, k_hi)
tmod = eng.tensor(3600.0 * np.arange(k_hi + 1))
yd, xd, td = eng.tensor(y), eng.tensor(x), eng.tensor(t)
n = len(y)
NP = eng.empty((n, 2), torch.int64); SM = eng.empty((n, 4, 2), torch.int64); WB = eng.empty((n, 4), torch.float64)
v_np = eng.empty((n,), torch.float64); v_bl = eng.empty((n,), torch.float64)
for _ in range(3):
    eng.map_interp(yd, xd, td, rd, r, (r, r), tmod, slab, 0, NP, SM, WB, v_np, v_bl)
ctx.sync()
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.map_interp(yd, xd, td, rd, r, (r, r), tmod, slab, 0, NP, SM, WB, v_np, v_bl)
ctx.sync()
eng.interp(td, SM, WB, tmod, slab, 0, v_np, v_bl)
eng.mesh_weights_interp(yd, xd, td, NP, tmod, slab, 0, SM, WB, v_np, v_bl)
eng.mesh_weights(yd, xd, NP, SM, WB)
ctx.sync()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ncu_case done: %d points" % n)
