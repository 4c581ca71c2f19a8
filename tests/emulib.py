"""TEST INFRASTRUCTURE: an emulation of the C ABI (include/gonzag_b200.h) backed by the CPU oracle, so that the
HOST logic of the drop-in classes -- `ModGrid`, `SatTrack`, `Model2SatTrack`, the CLI, the record providers and
file readers, slab planning, packaging of the outputs -- runs in the CPU test suite against the reference's golden
outputs.  The CUDA kernels are NOT under test here (that is what `-m gpu` is for) and nothing under `gonzag_b200/`
knows this file exists: tests install it with ``monkeypatch.setattr(gonzag_b200._lib, "lib", lambda: EmuLib())``.

"Device memory" is host memory (`ctypes` buffers), so a device pointer is an address NumPy can view; host-only entry
points (`gzb_pack_mask`, `gzb_copy_swap`, `gzb_version`) go to the real library.  Entry points the host classes do
not use raise `NotImplementedError`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from gonzag_b200 import _lib as L
from oracle import gz_oracle as orc

_REAL_LIB = L.lib            # captured before any test replaces it
_MEMO = {}                   # the test bodies map the same few tracks again and again: oracle results by input digest


def _memo(tag, fn, *arrays_and_scalars):
    import hashlib
    h = hashlib.blake2b(tag.encode(), digest_size=16)
    for a in arrays_and_scalars:
        h.update(np.ascontiguousarray(a).tobytes() if isinstance(a, np.ndarray) else repr(a).encode())
    key = h.digest()
    if key not in _MEMO:
        _MEMO[key] = fn()
    return _MEMO[key].copy()


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, (int, np.integer)):
        return int(p)
    if hasattr(p, "_obj"):                      # ctypes.byref(x)
        return C.addressof(p._obj)
    if isinstance(p, C.c_void_p):
        return p.value or 0
    return C.cast(p, C.c_void_p).value or 0


def _view(p, shape, dtype):
    a = _addr(p)
    n = int(np.prod(shape))
    if n == 0:
        return np.empty(shape, dtype)
    assert a, "null pointer where an array was expected"
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(a)
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def _target(p):
    """the ctypes object behind a byref() argument"""
    return p._obj


class _Ctx:
    def __init__(self, nj, ni, lat, lon, angle, mask, kew):
        self.nj, self.ni = nj, ni
        self.lat, self.lon = lat.copy(), lon.copy()
        self.angle = None if angle is None else angle.copy()
        self.mask = None if mask is None else mask.astype(np.int64)
        self.kew = kew
        self.opt = {"edge_exclusion": 1, "force_heavy": 0}
        self.launches = self.h2d = self.d2h = 0
        self.err = b""
        self.time_error = False


class EmuLib:
    def __init__(self):
        self._real = _REAL_LIB()
        self._ctx = {}
        self._next = 0x1000
        self._blocks = {}            # address -> (ctypes buffer, kind)  kind 1 = pinned host, 2 = device
        self._gerr = b""

    # ------------------------------------------------------------------ anything not emulated
    def __getattr__(self, name):
        if name in ("gzb_pack_mask", "gzb_copy_swap", "gzb_version"):
            return getattr(self._real, name)
        if name.startswith("gzb_"):
            def missing(*a, **k):
                raise NotImplementedError("tests/emulib.py does not emulate " + name)
            return missing
        raise AttributeError(name)

    def _c(self, h):
        return self._ctx[_addr(h)]

    def _fail(self, ctx, code, msg):
        if ctx is None:
            self._gerr = msg.encode()
        else:
            ctx.err = msg.encode()
        return code

    # ------------------------------------------------------------------ library level
    def gzb_last_global_error(self):
        return self._gerr

    def gzb_last_error(self, h):
        return self._c(h).err

    def gzb_device_count(self, n_out):
        _target(n_out).value = 1
        return L.GZB_OK

    def _kind(self, p):
        a = _addr(p)
        for base, (buf, k) in self._blocks.items():
            if base <= a < base + len(buf):
                return k
        return 0

    def gzb_pointer_kind(self, p, kind_out):
        _target(kind_out).value = self._kind(p)
        return L.GZB_OK

    def _alloc(self, nbytes, kind):
        buf = (C.c_char * max(int(nbytes), 1))()
        a = C.addressof(buf)
        self._blocks[a] = (buf, kind)
        return a

    def gzb_host_alloc(self, nbytes, p_out):
        _target(p_out).value = self._alloc(nbytes, 1)
        return L.GZB_OK

    def gzb_host_free(self, p):
        self._blocks.pop(_addr(p), None)
        return L.GZB_OK

    # ------------------------------------------------------------------ context
    def gzb_create(self, device, nj, ni, lat, lon, angle, mask, kew, h_out):
        _target(h_out).value = None
        if nj < 3 or ni < 3 or not _addr(lat) or not _addr(lon):
            return self._fail(None, L.GZB_ERR_ARG, "gzb_create: need lat/lon of shape (nj>=3, ni>=3)")
        la, lo = _view(lat, (nj, ni), np.float64), _view(lon, (nj, ni), np.float64)
        if not (np.isfinite(la).all() and np.isfinite(lo).all()):
            return self._fail(None, L.GZB_ERR_ARG, "gzb_create: lat/lon must be finite")
        ang = _view(angle, (nj, ni), np.float64) if _addr(angle) else None
        msk = _view(mask, (nj, ni), np.int8) if _addr(mask) else None
        ctx = _Ctx(nj, ni, la, lo, ang, msk, int(kew))
        ctx.h2d += 16 * nj * ni + (8 * nj * ni if ang is not None else 0) + (nj * ni if msk is not None else 0)
        self._next += 0x100
        self._ctx[self._next] = ctx
        _target(h_out).value = self._next
        return L.GZB_OK

    def gzb_create_imask(self, device, nj, ni, lat, lon, angle, mask, isz, kew, h_out):
        if _addr(mask) and isz != 1:
            m8 = np.ascontiguousarray(_view(mask, (nj, ni), {2: np.int16, 4: np.int32, 8: np.int64}[isz]).astype(np.int8))
            self._keep_mask = m8
            return self.gzb_create(device, nj, ni, lat, lon, angle, C.c_void_p(m8.ctypes.data), kew, h_out)
        return self.gzb_create(device, nj, ni, lat, lon, angle, mask, kew, h_out)

    def gzb_destroy(self, h):
        self._ctx.pop(_addr(h), None)
        return L.GZB_OK

    def gzb_sync(self, h):
        if not _addr(h):
            return L.GZB_ERR_ARG
        ctx = self._c(h)
        if ctx.time_error:
            ctx.time_error = False
            return self._fail(ctx, L.GZB_ERR_TIME, "gzb_interp: track time outside the model records [tmod[0], tmod[nrec-1])")
        return L.GZB_OK

    def gzb_set_host_threads(self, n):
        return L.GZB_OK

    def gzb_grid_arrays(self, h, lat, lon, angle, mask):
        ctx = self._c(h)
        if ctx.mask is not None and getattr(ctx, "mask8", None) is None:
            ctx.mask8 = np.ascontiguousarray(ctx.mask.astype(np.int8))
        for p, a in ((lat, ctx.lat), (lon, ctx.lon), (angle, ctx.angle), (mask, getattr(ctx, "mask8", None))):
            if p is not None:
                _target(p).value = None if a is None else a.ctypes.data
        return L.GZB_OK

    def gzb_set_stream(self, h, s):
        return L.GZB_OK

    def gzb_use_own_stream(self, h):
        return L.GZB_OK

    def gzb_get_stats(self, h, out):
        ctx, s = self._c(h), _target(out)
        C.memset(C.addressof(s), 0, C.sizeof(s))
        s.kernel_launches, s.h2d_bytes, s.d2h_bytes = ctx.launches, ctx.h2d, ctx.d2h
        return L.GZB_OK

    def gzb_set_mask(self, h, mask, where):
        ctx = self._c(h)
        ctx.mask = _view(mask, (ctx.nj, ctx.ni), np.int8).astype(np.int64)
        ctx.mask8 = None
        ctx.h2d += ctx.nj * ctx.ni
        return L.GZB_OK

    def gzb_set_angle(self, h, angle, where):
        ctx = self._c(h)
        ctx.angle = _view(angle, (ctx.nj, ctx.ni), np.float64).copy() if _addr(angle) else None
        return L.GZB_OK

    def gzb_set_ew_per(self, h, kew):
        self._c(h).kew = int(kew)
        return L.GZB_OK

    def gzb_set_option(self, h, name, value):
        ctx = self._c(h)
        name = name.decode() if isinstance(name, bytes) else name
        known = ("zero_copy", "nearest_path", "heading_table", "overlap", "twin_chain", "fuse_k4", "k4_early", "mask_bits",
                 "prefetch", "pdl", "l2_fetch", "trace", "edge_exclusion", "force_heavy", "bulk_chunks", "ix64", "chain_first", "bulk_resident", "carveout", "push_blocks", "push_mode")
        if name not in known:
            return self._fail(ctx, L.GZB_ERR_ARG, "gzb_set_option: unknown option " + name)
        ctx.opt[name] = int(value)
        return L.GZB_OK

    # ------------------------------------------------------------------ device memory = host memory
    def gzb_dev_alloc(self, h, nbytes, p_out):
        _target(p_out).value = self._alloc(nbytes, 2)
        return L.GZB_OK

    def gzb_dev_free(self, h, p):
        self._blocks.pop(_addr(p), None)
        return L.GZB_OK

    def gzb_h2d(self, h, dst, src, nbytes):
        C.memmove(_addr(dst), _addr(src), int(nbytes))
        self._c(h).h2d += int(nbytes)
        return L.GZB_OK

    def gzb_d2h(self, h, dst, src, nbytes):
        C.memmove(_addr(dst), _addr(src), int(nbytes))
        self._c(h).d2h += int(nbytes)
        return L.GZB_OK

    def gzb_d2d(self, h, dst, src, nbytes):
        C.memmove(_addr(dst), _addr(src), int(nbytes))
        return L.GZB_OK

    def gzb_memset(self, h, dst, byte, nbytes):
        C.memset(_addr(dst), int(byte), int(nbytes))
        return L.GZB_OK

    # ------------------------------------------------------------------ grid heuristics, K0
    def gzb_lon_stats(self, h, vals4, cols2):
        ctx = self._c(h)
        X = np.mod(ctx.lon, 360.)
        XB = orc.frame_pm180(X.copy())
        v, c = _view(vals4, (4,), np.float64), _view(cols2, (2,), np.int64)
        v[:] = [X.min(), X.max(), XB.min(), XB.max()]
        c[:] = [int(X.argmin()) % ctx.ni, int(X.argmax()) % ctx.ni]
        ctx.launches += 2
        return L.GZB_OK

    def gzb_lon_to_pm180(self, h, lon_out):
        ctx = self._c(h)
        ctx.lon = orc.frame_pm180(ctx.lon.copy())
        if _addr(lon_out):
            _view(lon_out, (ctx.nj, ctx.ni), np.float64)[:] = ctx.lon
        ctx.launches += 1
        return L.GZB_OK

    def gzb_lat_range(self, h, lat_min, lat_max):
        ctx = self._c(h)
        _target(lat_min).value = float(ctx.lat.min())
        _target(lat_max).value = float(ctx.lat.max())
        return L.GZB_OK

    def gzb_grid_angle(self, h, out, where):
        ctx = self._c(h)
        ctx.angle = orc.grid_angle(ctx.lat, ctx.lon)
        if _addr(out):
            _view(out, (ctx.nj, ctx.ni), np.float64)[:] = ctx.angle
        ctx.launches += 1
        return L.GZB_OK

    def gzb_get_angle(self, h, out, where):
        ctx = self._c(h)
        if ctx.angle is None:
            return self._fail(ctx, L.GZB_ERR_STATE, "gzb_get_angle: the context holds no angle (no -D)")
        _view(out, (ctx.nj, ctx.ni), np.float64)[:] = ctx.angle
        return L.GZB_OK

    # ------------------------------------------------------------------ K1 - K3
    def _nearest(self, ctx, nt, yt, xt, rd, r, sj, si):
        Y, X = _view(yt, (nt,), np.float64), _view(xt, (nt,), np.float64)
        ctx.launches += 8
        if ctx.opt["edge_exclusion"]:
            return _memo("np", lambda: orc.nearest_chain(Y, X, ctx.lat, ctx.lon, ctx.kew, rd, int(r), seed=(int(sj), int(si))),
                         Y, X, ctx.lat, ctx.lon, ctx.kew, float(rd), int(r), int(sj), int(si))
        out = np.zeros((nt, 2), np.int64)
        jj, ji = int(sj), int(si)
        for t in range(nt):
            jj, ji = orc.nearest_point((Y[t], X[t]), ctx.lat, ctx.lon, rd_found_km=rd, j_prv=jj, i_prv=ji, np_box_r=int(r))
            out[t] = [jj, ji]
        return out

    def _mesh(self, ctx, nt, yt, xt, NP):
        if ctx.opt["force_heavy"]:
            raise NotImplementedError("tests/emulib.py: option force_heavy")
        Y, X = _view(yt, (nt,), np.float64), _view(xt, (nt,), np.float64)
        ctx.launches += 1
        return _memo("sm", lambda: orc.source_meshes(Y, X, ctx.lat, ctx.lon, NP, ctx.kew, angle=ctx.angle),
                     Y, X, ctx.lat, ctx.lon, NP, ctx.kew, ctx.angle if ctx.angle is not None else "no angle")

    def _wght(self, ctx, nt, yt, xt, SM):
        Y, X = _view(yt, (nt,), np.float64), _view(xt, (nt,), np.float64)
        ctx.launches += 1
        return _memo("wb", lambda: orc.weights(Y, X, ctx.lat, ctx.lon, SM), Y, X, ctx.lat, ctx.lon, SM)

    def gzb_nearest(self, h, nt, yt, xt, rd, r, sj, si, np_out, where):
        ctx = self._c(h)
        if nt:
            _view(np_out, (nt, 2), np.int64)[:] = self._nearest(ctx, nt, yt, xt, rd, r, sj, si)
        return L.GZB_OK

    def gzb_source_mesh(self, h, nt, yt, xt, np_in, sm_out, where):
        ctx = self._c(h)
        if nt:
            _view(sm_out, (nt, 4, 2), np.int64)[:] = self._mesh(ctx, nt, yt, xt, _view(np_in, (nt, 2), np.int64))
        return L.GZB_OK

    def gzb_weights(self, h, nt, yt, xt, sm_in, wb_out, where):
        ctx = self._c(h)
        if nt:
            _view(wb_out, (nt, 4), np.float64)[:] = self._wght(ctx, nt, yt, xt, _view(sm_in, (nt, 4, 2), np.int64))
        return L.GZB_OK

    def gzb_mesh_weights(self, h, nt, yt, xt, np_in, sm_out, wb_out, where):
        rc = self.gzb_source_mesh(h, nt, yt, xt, np_in, sm_out, where)
        return rc or self.gzb_weights(h, nt, yt, xt, sm_out, wb_out, where)

    def gzb_mapping(self, h, nt, yt, xt, rd, r, sj, si, np_out, sm_out, wb_out, where):
        rc = self.gzb_nearest(h, nt, yt, xt, rd, r, sj, si, np_out, where)
        return rc or self.gzb_mesh_weights(h, nt, yt, xt, np_out, sm_out, wb_out, where)

    # ------------------------------------------------------------------ K4
    def gzb_interp(self, h, nt, t, sm, wb, nrec, tmod, slab, dtype, k0, nk, rmiss, init, np_out, bl_out, where):
        ctx = self._c(h)
        if ctx.mask is None:
            return self._fail(ctx, L.GZB_ERR_STATE, "gzb_interp: the context has no land-sea mask (gzb_set_mask)")
        if nt == 0:
            return L.GZB_OK
        T, TM = _view(t, (nt,), np.float64), _view(tmod, (nrec,), np.float64)
        SM, WB = _view(sm, (nt, 4, 2), np.int64), _view(wb, (nt, 4), np.float64)
        S = _view(slab, (nk, ctx.nj, ctx.ni), np.float32 if dtype == L.GZB_F32 else np.float64)
        vnp, vbl = _view(np_out, (nt,), np.float64), _view(bl_out, (nt,), np.float64)
        if init:
            vnp[:] = rmiss
            vbl[:] = rmiss
        kt = np.searchsorted(TM, T, side="right") - 1
        if ((T < TM[0]) | (T >= TM[-1])).any():
            ctx.time_error = True
            if where == L.GZB_HOST:
                return self.gzb_sync(h)
        inside = np.where((kt >= k0) & (kt + 1 <= k0 + nk - 1) & (T >= TM[0]) & (T < TM[-1]))[0]
        if inside.size:
            a, b = orc.interp_track(T[inside], TM, lambda k: S[k - k0], ctx.mask, SM[inside], WB[inside], rmissval=rmiss)
            vnp[inside], vbl[inside] = a, b
        ctx.launches += 1
        if self._kind(slab) == 0:
            ctx.h2d += S.nbytes
        return L.GZB_OK

    def gzb_map_interp(self, h, nt, yt, xt, tt, rd, r, sj, si, nrec, tmod, slab, dtype, k0, nk, rmiss,
                       np_out, sm_out, wb_out, vnp_out, vbl_out):
        rc = self.gzb_mapping(h, nt, yt, xt, rd, r, sj, si, np_out, sm_out, wb_out, L.GZB_DEVICE)
        return rc or self.gzb_interp(h, nt, tt, sm_out, wb_out, nrec, tmod, slab, dtype, k0, nk, rmiss, 1, vnp_out, vbl_out,
                                     L.GZB_DEVICE)

    def gzb_map_interp_multi(self, h, nseg, seg_start, seg_seed, nt, yt, xt, tt, rd, r, nrec, tmod, slab, dtype, k0, nk, rmiss,
                             np_out, sm_out, wb_out, vnp_out, vbl_out):
        st = list(_view(seg_start, (nseg,), np.int64)) + [int(nt)]
        sd = _view(seg_seed, (nseg, 2), np.int64)
        off = lambda p, k: _addr(p) + int(k)
        for q in range(nseg):
            a_, b_ = int(st[q]), int(st[q + 1])
            rc = self.gzb_map_interp(h, b_ - a_, off(yt, a_ * 8), off(xt, a_ * 8), off(tt, a_ * 8), rd, r, int(sd[q, 0]), int(sd[q, 1]),
                                     nrec, tmod, slab, dtype, k0, nk, rmiss, off(np_out, a_ * 16), off(sm_out, a_ * 64),
                                     off(wb_out, a_ * 32), off(vnp_out, a_ * 8), off(vbl_out, a_ * 8))
            if rc:
                return rc
        return L.GZB_OK

    # ------------------------------------------------------------------ a14
    def gzb_track_distance(self, h, nt, yt, xt, out, where):
        ctx = self._c(h)
        if nt:
            _view(out, (nt,), np.float64)[:] = orc.along_track_distance(_view(yt, (nt,), np.float64), _view(xt, (nt,), np.float64))
        ctx.launches += 3
        return L.GZB_OK

    def gzb_track_mask(self, h, nt, bl, sat, imask_out, bad_out):
        if nt:
            a, b = _view(bl, (nt,), np.float64), _view(sat, (nt,), np.float64)
            with np.errstate(invalid="ignore"):
                good = (a > -100.) & (a < 100.) & (b > -100.) & (b < 100.)      # gonzag/mod2sat.py:157-158
            _view(imask_out, (nt,), np.int8)[:] = good
            if _addr(bad_out):
                _view(bad_out, (nt,), np.uint8)[:] = ~good
        self._c(h).launches += 1
        return L.GZB_OK

    def gzb_xnp_track_masked(self, h, nt, np_in, rmiss, xnp_out, missing_out, where):
        ctx = self._c(h)
        if ctx.mask is None:
            return self._fail(ctx, L.GZB_ERR_STATE, "gzb_xnp_track: the context has no land-sea mask (gzb_set_mask)")
        x = orc.xnp_track(_view(np_in, (nt, 2), np.int64), ctx.mask, rmissval=rmiss)
        _view(xnp_out, (ctx.nj, ctx.ni), np.float64)[:] = np.ma.getdata(x)
        if _addr(missing_out):
            _view(missing_out, (ctx.nj, ctx.ni), np.uint8)[:] = np.ma.getmaskarray(x)
        ctx.launches += 2
        return L.GZB_OK

    def gzb_xnp_track(self, h, nt, np_in, rmiss, xnp_out, where):
        return self.gzb_xnp_track_masked(h, nt, np_in, rmiss, xnp_out, None, where)

    # ------------------------------------------------------------------ batched scalar helpers
    def gzb_heading(self, device, n, a, b, c, d, out):
        A, B, Cc, D, O = [_view(p, (n,), np.float64) for p in (a, b, c, d, out)]
        for k in range(n):
            O[k] = orc.heading(A[k], B[k], Cc[k], D[k])
        return L.GZB_OK

    def gzb_alfabeta(self, device, n, vy, vx, ab):
        Y, X, O = _view(vy, (n, 5), np.float64), _view(vx, (n, 5), np.float64), _view(ab, (n, 2), np.float64)
        for k in range(n):
            O[k] = orc.local_coords(Y[k].copy(), X[k].copy())
        return L.GZB_OK

    def gzb_haversine(self, device, n, plat, plon, la, lo, out):
        _view(out, (n,), np.float64)[:] = orc.haversine_km(np.float64(plat), np.float64(plon), _view(la, (n,), np.float64),
                                                           _view(lo, (n,), np.float64))
        return L.GZB_OK

    def gzb_haversine_sclr(self, device, n, a, b, c, d, out):
        A, B, Cc, D, O = [_view(p, (n,), np.float64) for p in (a, b, c, d, out)]
        for k in range(n):
            O[k] = orc.haversine_scalar_km(A[k], B[k], Cc[k], D[k])
        return L.GZB_OK

    def gzb_radius_earth(self, device, n, a, out):
        A, O = _view(a, (n,), np.float64), _view(out, (n,), np.float64)
        for k in range(n):
            O[k] = orc.earth_radius_km(A[k])
        return L.GZB_OK
