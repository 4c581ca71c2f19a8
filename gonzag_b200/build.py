"""Build libgonzag_b200.so in-tree with nvcc for sm_100a.

    python -m gonzag_b200.build [--force]

No GPU is needed to build (nvcc cross-compiles).  -fmad=false keeps every multiply-add as
two roundings, which the parity-critical kernels rely on ("numerics").
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgonzag_b200.so")
SOURCES = ["gzb_grid.cu", "gzb_nearest.cu", "gzb_mesh.cu", "gzb_interp.cu", "gzb_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# experiment builds: GZB_NVCC_EXTRA="-DGZB_EXP_L64=1 -DGZB_EXP_HDG_STRIDE=8" python -m gonzag_b200.build --force
# (compile-time variants listed in csrc/gzb_common.cuh; without the variable the build is the product)
NVCC_FLAGS += os.environ.get("GZB_NVCC_EXTRA", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libgonzag_b200.so cannot be built")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "flags.txt")          # objects compiled with other flags (an experiment build) are stale
    flags_now = " ".join(NVCC_FLAGS)
    if not os.path.exists(stamp) or open(stamp).read() != flags_now:
        force = True
    with open(stamp, "w") as f:
        f.write(flags_now)
    headers = [os.path.join(CSRC, "gzb_common.cuh"), os.path.join(CSRC, "gzb_point.cuh"),
               os.path.join(os.path.dirname(HERE), "include", "gonzag_b200.h"), __file__]
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(o + ".log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (s, log))
        return log

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        logs = list(ex.map(compile_one, jobs))
    if verbose:
        for l in logs:
            print(l)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-cudart", "static", "-o", LIB] + objs,
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
