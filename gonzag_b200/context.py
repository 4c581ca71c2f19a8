"""`GridContext`: one model grid resident on one B200, plus thin typed wrappers over the C ABI.

Host NumPy in / host NumPy out methods mirror the stages of the reference's `BilinTrack`
(gonzag/bilin_mapping.py:566-617); the `dev_*` helpers let `Model2SatTrack` keep a whole track
on the device between the stages.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .config import rmissval as _RMISS


class DevBuf:
    """A device allocation owned by a `GridContext`."""

    __slots__ = ("ctx", "ptr", "nbytes")

    def __init__(self, ctx, nbytes):
        p = C.c_void_p()
        L.check(L.lib().gzb_dev_alloc(ctx._h, int(max(nbytes, 1)), C.byref(p)), ctx._h)
        self.ctx, self.ptr, self.nbytes = ctx, p.value, int(nbytes)

    def free(self):
        if self.ptr is not None and self.ctx._h is not None:
            L.lib().gzb_dev_free(self.ctx._h, C.c_void_p(self.ptr))
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class GridContext:
    def __init__(self, Ys, Xs, angle=None, mask=None, k_ew_per=-1, device=0):
        Ys = L.as_f64(Ys)
        Xs = L.as_f64(Xs)
        if Ys.ndim != 2 or Ys.shape != Xs.shape:
            raise ValueError("Ys and Xs must be 2-D arrays of the same shape")
        self.Nj, self.Ni = Ys.shape
        self.device = int(device)
        self._h = None
        ang = None
        if angle is not None and np.shape(angle) == Ys.shape:
            ang = L.as_f64(angle)
        msk, isz = None, 1
        if mask is not None:
            m = np.asarray(np.ma.getdata(mask))
            if m.shape != (self.Nj, self.Ni):
                raise ValueError("mask has shape %s, grid is %s" % (m.shape, (self.Nj, self.Ni)))
            if m.dtype.kind in "iu" and m.dtype.isnative and m.flags["C_CONTIGUOUS"] and m.dtype.itemsize in (2, 4, 8):
                msk, isz = m, m.dtype.itemsize          # the reference's int64 {0,1} mask: packed to int8 by the upload itself
            else:
                msk = self._mask8(mask)
        h = C.c_void_p()
        L.check(L.lib().gzb_create_imask(self.device, self.Nj, self.Ni, L.ptr(Ys), L.ptr(Xs), L.ptr(ang),
                                         L.ptr(msk), isz, int(k_ew_per), C.byref(h)))
        self._h = h
        self.k_ew_per = int(k_ew_per)
        self.has_mask = msk is not None
        self.has_angle = ang is not None

    @classmethod
    def from_device(cls, Nj, Ni, lat_ptr, lon_ptr, angle_ptr=None, mask_ptr=None, k_ew_per=-1, device=0):
        """A context built from arrays that already sit in DEVICE memory of GPU ``device`` (raw addresses: lat, lon
        float64 (Nj, Ni); angle float64 or None; mask int8 or None) -- how the ranks of a sharded job get their copy of
        a grid that one rank uploaded and the others received over NVLink (`gonzag_b200.sharded.replicate_grid`)."""
        self = cls.__new__(cls)
        self.Nj, self.Ni = int(Nj), int(Ni)
        self.device = int(device)
        self._h = None
        h = C.c_void_p()
        vp = lambda p: C.c_void_p(int(p)) if p else None
        L.check(L.lib().gzb_create(self.device, self.Nj, self.Ni, vp(lat_ptr), vp(lon_ptr), vp(angle_ptr), vp(mask_ptr),
                                   int(k_ew_per), C.byref(h)))
        self._h = h
        self.k_ew_per = int(k_ew_per)
        self.has_mask = bool(mask_ptr)
        self.has_angle = bool(angle_ptr)
        return self

    def grid_arrays(self):
        """Device addresses (lat, lon, angle or None, int8 mask or None) of the arrays this context holds as uploaded."""
        p = [C.c_void_p() for _ in range(4)]
        self._ck(L.lib().gzb_grid_arrays(self._h, *[C.byref(q) for q in p]))
        return tuple(q.value for q in p)

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if self._h is not None:
            L.lib().gzb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _mask8(self, mask):
        m = np.asarray(np.ma.getdata(mask))
        if m.shape != (self.Nj, self.Ni):
            raise ValueError("mask has shape %s, grid is %s" % (m.shape, (self.Nj, self.Ni)))
        if m.dtype in (np.int8, np.uint8, np.bool_):
            return np.ascontiguousarray(m).view(np.int8)
        if m.dtype.kind in "iu" and m.dtype.isnative and m.flags["C_CONTIGUOUS"]:
            # the reference's int64 {0,1} mask (gonzag/ncio.py:164-191): astype(int8) on the library's host threads
            m8 = np.empty(m.shape, dtype=np.int8)
            L.check(L.lib().gzb_pack_mask(L.ptr(m), m.dtype.itemsize, m.size, L.ptr(m8)))
            return m8
        return np.ascontiguousarray(m.astype(np.int8))

    def _ck(self, code):
        L.check(code, self._h)

    # ------------------------------------------------------------------ configuration
    def set_mask(self, mask):
        m8 = self._mask8(mask)          # keep the temporary alive across the call
        self._ck(L.lib().gzb_set_mask(self._h, L.ptr(m8), L.GZB_HOST))
        self.has_mask = True

    def set_angle(self, angle):
        if angle is None or np.shape(angle) != (self.Nj, self.Ni):
            self._ck(L.lib().gzb_set_angle(self._h, None, L.GZB_HOST))
            self.has_angle = False
        else:
            a64 = L.as_f64(angle)       # keep the temporary alive across the call
            self._ck(L.lib().gzb_set_angle(self._h, L.ptr(a64), L.GZB_HOST))
            self.has_angle = True

    def set_ew_per(self, k):
        self._ck(L.lib().gzb_set_ew_per(self._h, int(k)))
        self.k_ew_per = int(k)

    def set_option(self, name, value):
        self._ck(L.lib().gzb_set_option(self._h, name.encode(), int(value)))

    def set_stream(self, cuda_stream_ptr):
        """Run on a caller-provided cudaStream_t (0 / None = the legacy default stream)."""
        self._ck(L.lib().gzb_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr)) if cuda_stream_ptr else None))

    def use_own_stream(self):
        self._ck(L.lib().gzb_use_own_stream(self._h))

    def sync(self):
        self._ck(L.lib().gzb_sync(self._h))

    def stats(self):
        s = L.Stats()
        self._ck(L.lib().gzb_get_stats(self._h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in L.Stats._fields_}

    # ------------------------------------------------------------------ host in / host out stages
    def grid_angle(self, out=True):
        """K0 -- `GridAngle` (gonzag/utils.py:226-262); also installs the angle in the context."""
        a = np.empty((self.Nj, self.Ni), dtype=np.float64) if out else None
        self._ck(L.lib().gzb_grid_angle(self._h, L.ptr(a), L.GZB_HOST))
        self.has_angle = True
        return a

    def nearest(self, Yt, Xt, rd_found_km, np_box_r, seed=None):
        Yt, Xt = L.as_f64(Yt), L.as_f64(Xt)
        n = Yt.shape[0]
        sj, si = (np_box_r, np_box_r) if seed is None else seed
        NP = np.empty((n, 2), dtype=np.int64)
        self._ck(L.lib().gzb_nearest(self._h, n, L.ptr(Yt), L.ptr(Xt), float(rd_found_km), int(np_box_r),
                                     int(sj), int(si), L.ptr(NP), L.GZB_HOST))
        return NP

    def source_mesh(self, Yt, Xt, NP):
        Yt, Xt, NP = L.as_f64(Yt), L.as_f64(Xt), L.as_i64(NP)
        n = Yt.shape[0]
        SM = np.empty((n, 4, 2), dtype=np.int64)
        self._ck(L.lib().gzb_source_mesh(self._h, n, L.ptr(Yt), L.ptr(Xt), L.ptr(NP), L.ptr(SM), L.GZB_HOST))
        return SM

    def weights(self, Yt, Xt, SM):
        Yt, Xt, SM = L.as_f64(Yt), L.as_f64(Xt), L.as_i64(SM)
        n = Yt.shape[0]
        WB = np.empty((n, 4), dtype=np.float64)
        self._ck(L.lib().gzb_weights(self._h, n, L.ptr(Yt), L.ptr(Xt), L.ptr(SM), L.ptr(WB), L.GZB_HOST))
        return WB

    def mapping(self, Yt, Xt, rd_found_km, np_box_r, seed=None):
        Yt, Xt = L.as_f64(Yt), L.as_f64(Xt)
        n = Yt.shape[0]
        sj, si = (np_box_r, np_box_r) if seed is None else seed
        NP = np.empty((n, 2), dtype=np.int64)
        SM = np.empty((n, 4, 2), dtype=np.int64)
        WB = np.empty((n, 4), dtype=np.float64)
        self._ck(L.lib().gzb_mapping(self._h, n, L.ptr(Yt), L.ptr(Xt), float(rd_found_km), int(np_box_r),
                                     int(sj), int(si), L.ptr(NP), L.ptr(SM), L.ptr(WB), L.GZB_HOST))
        return NP, SM, WB

    @staticmethod
    def _field_dtype(slab):
        if slab.dtype == np.float32:
            return L.GZB_F32
        if slab.dtype == np.float64:
            return L.GZB_F64
        raise TypeError("model field must be float32 or float64, got %s" % slab.dtype)

    def interp(self, t_sat, SM, WB, t_mod, slab, k0=0, rmissval=_RMISS, out=None):
        """K4 on host arrays.  ``slab`` holds records ``k0 .. k0+len(slab)-1``; pass ``out`` (the
        pair returned by a previous call) to accumulate over several slabs."""
        t_sat, WB, SM, t_mod = L.as_f64(t_sat), L.as_f64(WB), L.as_i64(SM), L.as_f64(t_mod)
        slab = np.ascontiguousarray(np.ma.getdata(slab))
        n = t_sat.shape[0]
        init = 1 if out is None else 0
        v_np, v_bl = (np.empty(n), np.empty(n)) if out is None else out
        self._ck(L.lib().gzb_interp(self._h, n, L.ptr(t_sat), L.ptr(SM), L.ptr(WB), t_mod.shape[0], L.ptr(t_mod),
                                    L.ptr(slab), self._field_dtype(slab), int(k0), slab.shape[0],
                                    float(rmissval), init, L.ptr(v_np), L.ptr(v_bl), L.GZB_HOST))
        return v_np, v_bl

    def track_distance(self, Yt, Xt):
        Yt, Xt = L.as_f64(Yt), L.as_f64(Xt)
        d = np.empty(Yt.shape[0], dtype=np.float64)
        self._ck(L.lib().gzb_track_distance(self._h, Yt.shape[0], L.ptr(Yt), L.ptr(Xt), L.ptr(d), L.GZB_HOST))
        return d

    def xnp_track(self, NP, rmissval=_RMISS):
        NP = L.as_i64(NP)
        x = np.empty((self.Nj, self.Ni), dtype=np.float64)
        self._ck(L.lib().gzb_xnp_track(self._h, NP.shape[0], L.ptr(NP), float(rmissval), L.ptr(x), L.GZB_HOST))
        return x

    # ------------------------------------------------------------------ device-resident helpers
    def dev_alloc(self, nbytes):
        return DevBuf(self, nbytes)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        b = DevBuf(self, arr.nbytes)
        self._ck(L.lib().gzb_h2d(self._h, C.c_void_p(b.ptr), L.ptr(arr), arr.nbytes))
        self.sync()          # the host array may be a temporary
        return b

    def download(self, buf, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        self._ck(L.lib().gzb_d2h(self._h, L.ptr(out), C.c_void_p(buf.ptr), out.nbytes))
        self.sync()
        return out
