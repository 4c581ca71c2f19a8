import numpy as np, time, sys
try:
    from numpy._core.multiarray import _set_madvise_hugepage
except Exception:
    from numpy.core.multiarray import _set_madvise_hugepage
print(open('/sys/kernel/mm/transparent_hugepage/enabled').read().strip(), '|', open('/sys/kernel/mm/transparent_hugepage/defrag').read().strip())
src = np.ones((3059, 4322))
for mode in (True, False, True, False):
    _set_madvise_hugepage(mode)
    ts = []
    for it in range(10):
        t0 = time.perf_counter()
        a = np.empty((3059, 4322))
        a[:] = src
        ts.append(1e3 * (time.perf_counter() - t0))
        del a
    print("madvise_hugepage=%s: alloc+fill 106 MB ms:" % mode, " ".join("%.0f" % v for v in ts))
