"""GPU experiment: where do the host-side stalls of an end-to-end call come from?"""
import sys, os, time, types, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gonzag_b200 as gzb
from gonzag_b200 import _lib as L
from gonzag_b200 import synthetic as syn
import bench

big_pinned = "--nopin" not in sys.argv
w = bench.WORKLOADS["C3"]
lat, lon, mask, res, rd, r = bench.make_grid(w)
Nj, Ni = lat.shape
if big_pinned:
    host_field = L.pinned_empty((242, Nj, Ni), np.float32)
    print("12.8 GB pinned allocated", flush=True)
lib = L.lib()
T = lambda: time.perf_counter()
for it in range(6):
    t0 = T()
    ctx = gzb.GridContext(lat, lon, mask=mask, k_ew_per=2)
    t1 = T()
    d = ctx.dev_alloc(Nj * Ni * 9)
    t2 = T()
    xnp = np.empty((Nj, Ni)); xmk = np.empty((Nj, Ni), np.bool_)
    t3 = T()
    ctx._ck(lib.gzb_d2h(ctx._h, L.ptr(xnp), C.c_void_p(d.ptr), xnp.nbytes))
    t4 = T()
    ctx._ck(lib.gzb_d2h(ctx._h, L.ptr(xmk), C.c_void_p(d.ptr + Nj * Ni * 8), xmk.nbytes))
    ctx.sync()
    t5 = T()
    d.free()
    t6 = T()
    ctx.close()
    t7 = T()
    del xnp, xmk
    t8 = T()
    print("it %d: create %.1f  dev_alloc %.1f  np.empty %.2f  d2h xnp %.1f  d2h mask+sync %.1f  dev_free %.1f  close %.1f  del %.1f ms"
          % (it, *(1e3 * v for v in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t8 - t7))), flush=True)
