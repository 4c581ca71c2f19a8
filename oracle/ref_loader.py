"""TEST INFRASTRUCTURE -- locating and importing the UNMODIFIED reference (brodeau/gonzag, pure Python).

Only tests/, __graft_entry__.smoke()/build() and bench.py's reference / cpu_baseline legs may import this module;
nothing under gonzag_b200/ does.

Where the reference comes from, in this order:
  1. $GONZAG_REFERENCE or /root/reference   (the checkout, present in the build container only);
  2. oracle/_ref/                            (a verbatim copy of the package directory made by `make -C oracle`,
     i.e. oracle/make_ref.py; git-ignored build output that travels to the GPU box like the built .so --
     reference SOURCES are never committed).

Two things the reference imports are not installed offline and are injected through ``sys.modules`` first:
  * ``gonzag.ncio`` (netCDF4 readers) -> an in-memory stand-in (`GetModel2DVar`, `GetSatSSH`, ...), so that
    `Model2SatTrack` (gonzag/mod2sat.py:20) runs without files;
  * ``shapely.geometry`` -> the planar strict-interior point-in-quad stand-in of oracle/gz_oracle.py (only the
    heavy-duty quadrant branch, gonzag/bilin_mapping.py:415-426, reaches it: that branch is "parity unpinned").
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY = os.path.join(HERE, "_ref")


def find_reference():
    """Directory that holds the `gonzag` package of the reference, or None."""
    for root in (os.environ.get("GONZAG_REFERENCE"), "/root/reference", REF_COPY):
        if root and os.path.isfile(os.path.join(root, "gonzag", "bilin_mapping.py")):
            return root
    return None


def reference_kind(root):
    return "copy (oracle/_ref)" if os.path.abspath(root) == os.path.abspath(REF_COPY) else "checkout"


def load_reference(ref_root=None):
    """(the imported reference package, the store behind its in-memory `ncio` stand-in)."""
    from oracle.gz_oracle import point_strictly_in_quad      # shapely stand-in only
    if ref_root is None:
        ref_root = find_reference()
    if ref_root is None:
        raise RuntimeError("the reference is neither mounted (/root/reference) nor copied (oracle/_ref: run `make -C oracle`)")
    sys.dont_write_bytecode = True          # the checkout is read-only
    shp = types.ModuleType("shapely")
    geo = types.ModuleType("shapely.geometry")
    pol = types.ModuleType("shapely.geometry.polygon")

    class Point:
        def __init__(self, a, b):
            self.a, self.b = a, b

    class Polygon:
        def __init__(self, pts):
            self.qa = tuple(p[0] for p in pts)
            self.qb = tuple(p[1] for p in pts)

        def contains(self, p):
            return point_strictly_in_quad(p.a, p.b, self.qa, self.qb)

    geo.Point = Point
    pol.Polygon = Polygon
    geo.polygon = pol
    shp.geometry = geo
    sys.modules["shapely"] = shp
    sys.modules["shapely.geometry"] = geo
    sys.modules["shapely.geometry.polygon"] = pol

    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import gonzag  # noqa
    store = {"fields": {}, "sat": {}, "src": {}}
    nc = types.ModuleType("gonzag.ncio")

    def _rd(fn):
        # files registered as in-memory sources answer through them; the two legacy stand-ins below serve the
        # SimpleNamespace cases
        def call(f, *a, **k):
            return getattr(store["src"][f], fn)(*a, **k)
        return call

    def get2d(f, v, kt=0):
        if f in store["src"]:
            return np.array(store["src"][f].GetModel2DVar(v, kt=kt), copy=True)
        return store["fields"][f][kt].copy()   # a file read returns a fresh array

    def getsat(f, v, kt1=0, kt2=0, ikeep=[]):
        if f in store["src"]:
            return store["src"][f].GetSatSSH(v, kt1=kt1, kt2=kt2, ikeep=ikeep)
        return store["sat"][f]

    nc.GetModel2DVar = get2d
    nc.GetSatSSH = getsat
    for fn in ("GetTimeEpochVector", "GetTimeInfo", "GetModelCoor", "GetModelLSM", "GetSatCoor"):
        setattr(nc, fn, _rd(fn))
    nc.Save2Dfield = lambda *a, **k: None
    nc.SaveTimeSeries = lambda *a, **k: 0
    sys.modules["gonzag.ncio"] = nc
    gonzag.ncio = nc
    return gonzag, store
