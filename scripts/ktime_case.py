"""KTIME build only (GZB_NVCC_EXTRA="-DGZB_KTIME=1" python -m gonzag_b200.build --force): the in-kernel time marks of whole
gzb_map_interp steps on the C3 bench workload; the library prints the table at every gzb_sync.
    python scripts/ktime_case.py [reps] [name=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import gonzag_b200 as gzb
from gonzag_b200.multi import CudaEngine
import bench

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
w = bench.WORKLOADS["C3"]
lat, lon, mask, res, rd, r = bench.make_grid(w)
t, y, x = bench.make_track(w, 0, 1)
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
ctx = gzb.GridContext(lat, lon, mask=mask, k_ew_per=2)
ctx.grid_angle(out=False)
ctx.set_option("heading_table", 2)
for kv in sys.argv[2:]:
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush2 = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
sink = torch.zeros(1, dtype=torch.int64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ctx.sync()
for it in range(reps):
    torch.cuda.synchronize()
    flush.zero_(); flush.zero_()
    torch.sum(flush2, dim=0, keepdim=True, out=sink)
    e0.record()
    eng.map_interp(yd, xd, td, rd, r, (r, r), tmod, slab, 0, NP, SM, WB, v_np, v_bl)
    e1.record()
    torch.cuda.synchronize()
    print("step %d: %.1f us between events" % (it, 1e3 * e0.elapsed_time(e1)), file=sys.stderr, flush=True)
    ctx.sync()
# K1 alone
for it in range(2):
    torch.cuda.synchronize()
    flush.zero_(); flush.zero_()
    torch.sum(flush2, dim=0, keepdim=True, out=sink)
    e0.record()
    eng.nearest(yd, xd, rd, r, (r, r), NP)
    e1.record()
    torch.cuda.synchronize()
    print("K1 alone %d: %.1f us between events" % (it, 1e3 * e0.elapsed_time(e1)), file=sys.stderr, flush=True)
    ctx.sync()
