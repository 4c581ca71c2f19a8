"""GPU experiment on the bench workload (C3): any integer option of gzb_set_option, A/B, on the stages of a step.
    python scripts/ab_opt.py pdl=0 pdl=1 pdl=0 pdl=1          (each argument: comma-separated name=value pairs)
Every setting must reproduce the outputs of the first one bit for bit."""
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
eng = CudaEngine(ctx, torch)
k_hi = int(t[-1] // 3600.0) + 1
slab = bench.field_on_device(torch, dev, torch.as_tensor(lat, device=dev), torch.as_tensor(lon, device=dev), 0, k_hi)
tmod = eng.tensor(3600.0 * np.arange(k_hi + 1))
yd, xd, td = eng.tensor(y), eng.tensor(x), eng.tensor(t)
n = len(y)
NP = eng.empty((n, 2), torch.int64); SM = eng.empty((n, 4, 2), torch.int64); WB = eng.empty((n, 4), torch.float64)
v_np = eng.empty((n,), torch.float64); v_bl = eng.empty((n,), torch.float64)
eng.nearest(yd, xd, rd, r, (r, r), NP)
eng.mesh_weights(yd, xd, NP, SM, WB)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stages = (("K4", lambda: eng.interp(td, SM, WB, tmod, slab, 0, v_np, v_bl)),
          ("bulk", lambda: eng.mesh_weights_interp(yd, xd, td, NP, tmod, slab, 0, SM, WB, v_np, v_bl)),
          ("K2K3", lambda: eng.mesh_weights(yd, xd, NP, SM, WB)),
          ("K1", lambda: eng.nearest(yd, xd, rd, r, (r, r), NP)),
          ("path", lambda: eng.map_interp(yd, xd, td, rd, r, (r, r), tmod, slab, 0, NP, SM, WB, v_np, v_bl)))
if os.environ.get("AB_STAGES"):
    stages = tuple(st for st in stages if st[0] in os.environ["AB_STAGES"].split(","))
ref = {}
flush2 = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
sink = torch.zeros(1, dtype=torch.int64, device=dev)
for cfg in sys.argv[1:]:
    for kv in cfg.split(","):
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    print(cfg, flush=True)
    for name, fn in stages:
        ms = []
        for it in range(11):
            torch.cuda.synchronize()
            flush.zero_(); flush.zero_()
            torch.sum(flush2, dim=0, keepdim=True, out=sink)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ctx.sync()
        got = [a.clone() for a in (NP, SM, WB.view(torch.int64), v_np.view(torch.int64), v_bl.view(torch.int64))]
        if name not in ref:
            ref[name] = got
        same = all(bool((a == b).all()) for a, b in zip(got, ref[name]))
        print("  %-5s %7.2f us (min %7.2f)  outputs %s" % (name, 1e3 * np.mean(ms[2:]), 1e3 * np.min(ms[2:]),
                                                         "identical" if same else "DIFFER"), flush=True)
