"""CPU model of the DRAM traffic of the gathers on the bench workload: unique 32-byte sectors / 64-byte atoms /
128-byte lines touched per track point by each per-cell array (perfect reuse inside the kernel assumed).
Compared with ncu (K4: 286 B/point read, bulk kernel: 516 B/point) the measured traffic matches the 128-byte figure:
L2 fills whole lines from DRAM.  Needs the oracle (test infrastructure): a measurement helper, not product code."""
import sys, time; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gonzag_b200 import synthetic as syn
from oracle import gz_oracle as orc
import bench
w = bench.WORKLOADS["C3"]
lat, lon, mask, res, rd, r = bench.make_grid(w)
t, y, x = bench.make_track(w, 0, 1)
n = 40000
t0 = time.time()
NP = orc.nearest_chain(y[:n], x[:n], lat, lon, 2, rd, r)
SM = orc.source_meshes(y[:n], x[:n], lat, lon, NP, 2, angle=None)
print("oracle", time.time() - t0, "s")
Nj, Ni = lat.shape
ok = SM[:, 0, 0] >= 0
flat = (SM[ok, :, 0] * Ni + SM[ok, :, 1]).ravel()      # cells touched (4 per point)
npts = ok.sum()
def uniq(item_bytes, gran, stride=None):
    stride = stride or item_bytes
    a = flat * stride
    b = a + item_bytes - 1
    blocks = np.unique(np.concatenate([a // gran, b // gran]))
    return len(blocks) * gran / npts
print("points", npts)
for name, ib, st in (("field f32 (per record)", 4, None), ("ll double2", 16, None), ("angle f64", 8, None)):
    print("%-24s unique 32B-sector traffic %.1f B/pt   unique 64B-atom traffic %.1f B/pt   128B-line %.1f" % (name, uniq(ib, 32, st), uniq(ib, 64, st), uniq(ib,128,st)))
p1 = (NP[ok, 0] * Ni + NP[ok, 1])
for name, ib, st in (("heading table 48B", 48, 48), ("heading table 48B @64", 48, 64)):
    a = p1 * st; b = a + ib - 1
    for gran in (32, 64):
        blocks = np.unique(np.concatenate([a // gran, (a+16)//gran, (a+32)//gran, b // gran]))
        print("%-24s gran %d: %.1f B/pt" % (name, gran, len(blocks) * gran / npts))
