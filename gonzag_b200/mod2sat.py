"""Drop-in for the reference's gonzag/mod2sat.py (`Model2SatTrack`, gonzag/mod2sat.py:18-186).

Same constructor, same attributes (`time, lat, lon, mask, ssh_mod_np, ssh_mod, ssh_sat, distance,
XNPtrack`), but the track stays on the GPU from the mapping to the interpolated series: the host
loops only over *slabs of model records*, never over track points.

Where the model records and the satellite SSH come from, in order of preference:
  * ``MG.field``      (nrec, Nj, Ni) array in memory, or a `RecordWindow` (records k0 .. of a longer series: what a
    rank of a sharded job holds)  /  ``MG.get_record(k)`` callable,
    ``ST.ssh``        (Nt,) array in memory;
  * ``MG.source`` / ``ST.source``: the in-memory sources of `gonzag_b200.sources` (what
    `gonzag_b200.ModGrid` / `SatTrack` were built from), whole array or one record at a time;
  * otherwise the reference's own readers ``GetModel2DVar`` / ``GetSatSSH`` from
    ``gonzag.ncio`` / ``gonzag.zarrio`` (gonzag/mod2sat.py:27-31) when that package is importable,
    which is the drop-in situation inside ``alongtrack_sat_vs_nemo.py``.

Record index convention, kept from the reference: record ``k`` is the one whose time is
``MG.time[k]`` *as the reader numbers it*, i.e. the reference asks its reader for ``kt = k`` with
``k`` an index into the (possibly sliced) ``MG.time`` (gonzag/mod2sat.py:94-112).

How the work is issued (`TrackJob`: one contiguous piece of a track on one GPU -- the whole track here, one time
segment per rank in `gonzag_b200.sharded`):
  * a field held as ONE array that the GPU can read (pinned + mapped host memory: gathered in
    place over PCIe; or small enough to be copied to HBM first) goes through `gzb_map_interp`:
    K1-K4 in one call, the chain replays of K1 beside K2-K4;
  * anything else (pageable big arrays, per-record providers) goes through `gzb_mapping` and then
    `gzb_interp` slab by slab, records staged through pinned memory, copies overlapped with K4;
  * the along-track distance, the validity mask (`gzb_track_mask`) and -- only when somebody reads it -- the
    nearest-point map `XNPtrack` are computed on the GPU as well; the host wraps the arrays.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import _lib as L
from . import config
from .context import GridContext
from .utils import SearchBoxSize


def _reference_readers():
    try:
        from gonzag.config import IsZarr
        if not IsZarr:
            from gonzag.ncio import GetModel2DVar, GetSatSSH
        else:
            from gonzag.zarrio import GetModel2DVar, GetSatSSH
        return GetModel2DVar, GetSatSSH
    except Exception as exc:   # pragma: no cover - depends on the environment
        raise RuntimeError("no in-memory field provider (MG.field / MG.get_record / MG.source, ST.ssh / ST.source) "
                           "and the reference's I/O readers are not importable: %s" % exc)


class MappingView:
    """What the reference keeps in its local `BT` (gonzag/mod2sat.py:49): NP, SM, WB on the host."""

    def __init__(self, NP, SM, WB):
        self.NP, self.SM, self.WB = NP, SM, WB


class RecordWindow:
    """Records ``k0 .. k0 + len(array) - 1`` of a model series held as one array (float32 / float64, C order): what
    ``MG.field`` may be when no process holds the whole series -- each rank of a sharded job passes the window that
    brackets its own time segment.  ``k`` counts like the reference's reader argument ``kt`` (index into MG.time)."""

    def __init__(self, k0, array):
        self.k0 = int(k0)
        self.array = array

    @property
    def shape(self):
        return self.array.shape


def plan_slabs(t_sat, t_mod, max_records):
    """Host-side planning: which model records are needed, grouped in slabs of at most
    ``max_records`` consecutive records (consecutive slabs share one record).  Mirrors the
    bracket rule of gonzag/mod2sat.py:94-96.  Returns a list of (k_first, k_last_inclusive)."""
    t_sat = np.asarray(t_sat, dtype=np.float64)
    t_mod = np.asarray(t_mod, dtype=np.float64)
    if t_sat.size == 0:
        return []
    if t_sat.min() < t_mod[0] or not (t_sat.max() < t_mod[-1]):
        raise ValueError("track times must lie inside [MG.time[0], MG.time[-1])")
    kt = np.unique(np.searchsorted(t_mod, t_sat, side="right") - 1)
    slabs = []
    a = int(kt[0])
    b = a + 1
    for k in kt[1:]:
        k = int(k)
        if k <= b and (k + 1 - a) < max_records:
            b = k + 1
        else:
            slabs.append((a, b))
            a, b = k, k + 1
    slabs.append((a, b))
    return slabs


def _providers(MG, name_ssh_mod):
    """(whole field array or None, first record of that array, per-record callable or None) for the model variable."""
    field = getattr(MG, "field", None)
    get_record = getattr(MG, "get_record", None)
    k0 = 0
    if field is None and get_record is None:
        src = getattr(MG, "source", None)
        if src is not None and hasattr(src, "field_array"):
            field = src.field_array(name_ssh_mod)
            if field is None:
                get_record = lambda k: src.GetModel2DVar(name_ssh_mod, kt=k)
        elif src is not None:
            get_record = lambda k: src.GetModel2DVar(name_ssh_mod, kt=k)
        else:
            reader, _ = _reference_readers()
            get_record = lambda k: reader(MG.file, name_ssh_mod, kt=k)
    if isinstance(field, RecordWindow):
        k0, field = field.k0, field.array
    return field, k0, get_record


def _record_writer(MG):
    """`into(name, k, out) -> bool` of the model source when it can write a record straight into a staging slot
    (`NetCDF3Source.record_into`), else None."""
    if getattr(MG, "field", None) is not None or getattr(MG, "get_record", None) is not None:
        return None
    return getattr(getattr(MG, "source", None), "record_into", None)


def satellite_ssh(ST, name_ssh_sat):
    """The satellite series of the track (gonzag/mod2sat.py:151), float64, plain ndarray."""
    ssh_attr = getattr(ST, "ssh", None)
    if ssh_attr is not None:
        vssh_s = np.array(ssh_attr, dtype=np.float64, copy=True)
    elif getattr(ST, "source", None) is not None:
        vssh_s = ST.source.GetSatSSH(name_ssh_sat, kt1=ST.jt1, kt2=ST.jt2, ikeep=ST.keepit)
    else:
        _, sat_reader = _reference_readers()
        vssh_s = sat_reader(ST.file, name_ssh_sat, kt1=ST.jt1, kt2=ST.jt2, ikeep=ST.keepit)
    # netCDF4 hands back a MaskedArray under auto-masking; the reference's nmp.where drops the masked entries
    # (gonzag/mod2sat.py:157-158), i.e. they end up with imask = 0 in a PLAIN int8 array: fill them with the
    # missing value first so that the mask below is a plain ndarray too
    if isinstance(vssh_s, np.ma.MaskedArray):
        vssh_s = np.ma.filled(np.ma.asarray(vssh_s, dtype=np.float64), float(config.rmissval))
    return np.ascontiguousarray(vssh_s, dtype=np.float64)


def _al(n):
    return (int(n) + 255) & ~255


class TrackJob:
    """One contiguous piece of a track on one GPU: device buffers, K1-K4, distance and validity mask.

    Device layout (one allocation).  ``y, x, dist`` have ONE extra leading slot: the point before the piece (the
    last point of the previous time segment in a sharded job) so that the along-track distance of the first own
    point includes the step from it; the own points start at ``P(name)``, the leading slot at ``P0(name)``."""

    def __init__(self, ctx, yt, xt, tt, tm, prev_point=None, n_alloc=None):
        self.ctx = ctx
        self.lib = L.lib()
        self.n = n = int(len(yt))
        self.nrec = int(len(tm))
        self.yt, self.xt, self.tt, self.tm = yt, xt, tt, tm
        self.prev_point = prev_point
        self.off, self.nbytes = self.layout(n if n_alloc is None else max(int(n_alloc), n), self.nrec)
        total = self.nbytes
        self.d_all = ctx.dev_alloc(total)
        self.d_field = None
        self.fused = False
        h, ck = ctx._h, ctx._ck
        for k, a in (("y", yt), ("x", xt), ("t", tt), ("tm", tm)):
            if a.nbytes:
                ck(self.lib.gzb_h2d(h, self.P(k), L.ptr(a), a.nbytes))
        if prev_point is not None and n:
            self._prev = np.array([prev_point[0], prev_point[1]], dtype=np.float64)      # kept alive: the copies are asynchronous
            ck(self.lib.gzb_h2d(h, self.P0("y"), L.ptr(self._prev[0:1]), 8))
            ck(self.lib.gzb_h2d(h, self.P0("x"), L.ptr(self._prev[1:2]), 8))

    @staticmethod
    def layout(n, nrec):
        """Byte offsets of the device arrays for a capacity of n points: inputs, then the products in the order a gather
        moves them (the two series, distance, mask, its complement, NP, then SM and WB), every array 256-byte aligned.
        The ranks of a sharded job allocate for the LONGEST segment, so the layout is the same on all of them."""
        sizes = [("y", (n + 1) * 8), ("x", (n + 1) * 8), ("t", n * 8), ("tm", nrec * 8), ("sat", n * 8),
                 ("vnp", n * 8), ("vbl", n * 8), ("dist", (n + 1) * 8), ("imask", n), ("bad", n),
                 ("np", n * 16), ("sm", n * 64), ("wb", n * 32), ("end", 0)]
        off, total = {}, 0
        for k, nb in sizes:
            off[k] = total
            total += _al(max(nb, 8)) if k != "end" else 0
        return off, total

    # ---- device addresses
    def P0(self, k):
        return C.c_void_p(self.d_all.ptr + self.off[k])

    def P(self, k):
        return C.c_void_p(self.d_all.ptr + self.off[k] + (8 if k in ("y", "x", "dist") else 0))

    def upload_mapping(self, mapping):
        mNP, mSM, mWB = L.as_i64(mapping.NP), L.as_i64(mapping.SM), L.as_f64(mapping.WB)
        n = self.n
        if mNP.shape != (n, 2) or mSM.shape != (n, 4, 2) or mWB.shape != (n, 4):
            raise ValueError("`mapping` does not belong to this track (Nt = %d)" % n)
        self._keep = (mNP, mSM, mWB)
        for k, a in (("np", mNP), ("sm", mSM), ("wb", mWB)):
            if a.nbytes:
                self.ctx._ck(self.lib.gzb_h2d(self.ctx._h, self.P(k), L.ptr(a), a.nbytes))

    # ---- K1 - K4
    def run(self, MG, name_ssh_mod, d_found_km, r, seed, mapping=None, slab_bytes=256 << 20, resident_bytes=2 << 30):
        """Mapping + interpolation of the piece (gonzag/mod2sat.py:49-139); ``seed`` = the (j, i) handed to the first
        point's search (the reference: (r, r); a later time segment: the last nearest point of its predecessor)."""
        ctx, lib, n, nrec = self.ctx, self.lib, self.n, self.nrec
        h, ck, P = ctx._h, ctx._ck, self.P
        rmiss = float(config.rmissval)
        if mapping is not None:
            self.upload_mapping(mapping)
        field, fk0, get_record = _providers(MG, name_ssh_mod)
        fused = False
        dt = None
        if field is not None and n > 0:
            field = np.ma.getdata(field)
            if not field.flags["C_CONTIGUOUS"]:
                field = np.ascontiguousarray(field)
            dt = GridContext._field_dtype(field)
            # the reference asks its reader for record k = index into the (possibly sliced) MG.time
            # (gonzag/mod2sat.py:94-112): a window starting after record 0 still reads records 0 .. nrec-1
            if fk0 + field.shape[0] > nrec:
                field = field[:nrec - fk0]
            if field.shape[0] < 2:
                raise ValueError("MG.field holds %d usable records, MG.time %d" % (field.shape[0], nrec))
            self._field = field          # kept alive: the GPU may gather from it in place
        if field is not None and n > 0 and mapping is None:
            kind = L.pointer_kind(field)
            sparse = n * 1024.0 < field.nbytes          # same rule as gzb_interp's zero-copy gather
            if kind >= 1 and (kind == 2 or sparse):
                slab_ptr = L.ptr(field)
                fused = True
            elif field.nbytes <= resident_bytes:
                self.d_field = ctx.dev_alloc(field.nbytes)    # small field: copy it to HBM once
                ck(lib.gzb_h2d(h, C.c_void_p(self.d_field.ptr), L.ptr(field), field.nbytes))
                slab_ptr = C.c_void_p(self.d_field.ptr)
                fused = True
        self.fused = fused
        if fused:
            # ---- K1-K4 in one call                                              (mod2sat.py:49-139)
            ck(lib.gzb_map_interp(h, n, P("y"), P("x"), P("t"), float(d_found_km), int(r), int(seed[0]), int(seed[1]), nrec,
                                  P("tm"), slab_ptr, dt, fk0, field.shape[0], rmiss, P("np"), P("sm"), P("wb"), P("vnp"),
                                  P("vbl")))
            return
        # ---- the mapping: K1 + K2 + K3                                      (mod2sat.py:49-50)
        if n > 0 and mapping is None:
            ck(lib.gzb_mapping(h, n, P("y"), P("x"), float(d_found_km), int(r), int(seed[0]), int(seed[1]), P("np"), P("sm"),
                               P("wb"), L.GZB_DEVICE))
        self.interp_slabs(MG, name_ssh_mod, field, fk0, get_record, slab_bytes)

    def interp_slabs(self, MG, name_ssh_mod, field, fk0, get_record, slab_bytes=256 << 20):
        """K4 over slabs of records                                           (mod2sat.py:86-139)"""
        ctx, lib, n, nrec = self.ctx, self.lib, self.n, self.nrec
        h, ck, P = ctx._h, ctx._ck, self.P
        rmiss = float(config.rmissval)
        if n <= 0:
            return
        first = True
        if field is not None:
            dt = GridContext._field_dtype(field)
            for (ka, kb) in plan_slabs(self.tt, self.tm, max(2, int(1 << 40))):
                if ka < fk0 or kb + 1 - fk0 > field.shape[0]:
                    raise ValueError("records %d..%d are needed, the record window holds %d..%d"
                                     % (ka, kb, fk0, fk0 + field.shape[0] - 1))
                slab = field[ka - fk0:kb + 1 - fk0]
                ck(lib.gzb_interp(h, n, P("t"), P("sm"), P("wb"), nrec, P("tm"), L.ptr(slab), dt, ka, kb + 1 - ka, rmiss,
                                  1 if first else 0, P("vnp"), P("vbl"), L.GZB_DEVICE))
                first = False
            return
        probe = np.ma.getdata(get_record(int(np.searchsorted(self.tm, self.tt[0], side="right") - 1)))
        dt = GridContext._field_dtype(probe)
        per_slab = max(2, int(slab_bytes // probe.nbytes))
        stage = L.pinned_empty((per_slab,) + probe.shape, probe.dtype)
        into = _record_writer(MG)
        try:
            for (ka, kb) in plan_slabs(self.tt, self.tm, per_slab):
                for k in range(ka, kb + 1):
                    if into is None or not into(name_ssh_mod, k, stage[k - ka]):
                        rec, dst = np.ma.getdata(get_record(k)), stage[k - ka]
                        if (isinstance(rec, np.ndarray) and rec.dtype == dst.dtype and rec.shape == dst.shape
                                and rec.flags["C_CONTIGUOUS"] and rec.nbytes >= (4 << 20)):
                            # one record of an ORCA12 field is 53 MB: NumPy's copy is one thread at ~5 GB/s
                            L.check(lib.gzb_copy_swap(L.ptr(dst), L.ptr(rec), rec.itemsize, rec.size, 0))
                        else:
                            dst[...] = rec
                ck(lib.gzb_interp(h, n, P("t"), P("sm"), P("wb"), nrec, P("tm"), L.ptr(stage), dt, ka, kb + 1 - ka, rmiss,
                                  1 if first else 0, P("vnp"), P("vbl"), L.GZB_DEVICE))
                first = False
        finally:
            ctx.sync()
            L.pinned_free(stage)

    # ---- a14: distance and validity mask on the device                       (mod2sat.py:146-148, 157-162)
    def distance_and_mask(self, vssh_s):
        ctx, lib, n = self.ctx, self.lib, self.n
        h, ck = ctx._h, ctx._ck
        if n <= 0:
            return
        if self.prev_point is not None:      # [previous point +] own points: dist0[0] = 0 at the previous point
            ck(lib.gzb_track_distance(h, n + 1, self.P0("y"), self.P0("x"), self.P0("dist"), L.GZB_DEVICE))
        else:
            ck(lib.gzb_track_distance(h, n, self.P("y"), self.P("x"), self.P("dist"), L.GZB_DEVICE))
        self._sat = vssh_s
        ck(lib.gzb_h2d(h, self.P("sat"), L.ptr(vssh_s), n * 8))
        ck(lib.gzb_track_mask(h, n, self.P("vbl"), self.P("sat"), self.P("imask"), self.P("bad")))

    def download_series(self):
        """ONE device -> host transfer for the two series, the distance, the mask and its complement; returns views
        of the host buffer: (ssh_np f64, ssh_bl f64, distance f64, imask i8, bad bool)."""
        n = self.n
        a, b = self.off["vnp"], self.off["np"]
        buf = np.empty(b - a, dtype=np.uint8)
        if n > 0:
            self.ctx._ck(self.lib.gzb_d2h(self.ctx._h, L.ptr(buf), C.c_void_p(self.d_all.ptr + a), buf.nbytes))
        part = lambda k, nbytes, dt, skip=0: buf[self.off[k] - a + skip:self.off[k] - a + skip + nbytes].view(dt)
        return (part("vnp", n * 8, np.float64), part("vbl", n * 8, np.float64), part("dist", n * 8, np.float64, 8),
                part("imask", n, np.int8), part("bad", n, np.bool_))

    def download_mapping(self):
        """NP, SM, WB sit one after the other on the device: one transfer (big enough for the copy threads)."""
        n, off = self.n, self.off
        mbuf = np.empty(off["wb"] + n * 32 - off["np"], dtype=np.uint8)
        part = lambda k, nbytes, dt, shape: mbuf[off[k] - off["np"]:off[k] - off["np"] + nbytes].view(dt).reshape(shape)
        NP, SM, WB = (part("np", n * 16, np.int64, (n, 2)), part("sm", n * 64, np.int64, (n, 4, 2)),
                      part("wb", n * 32, np.float64, (n, 4)))
        if n > 0:
            self.ctx._ck(self.lib.gzb_d2h(self.ctx._h, L.ptr(mbuf), self.P("np"), mbuf.nbytes))
        return MappingView(NP, SM, WB)

    def detach_np(self):
        """A device copy of NP that outlives the job's buffers (for the lazily computed XNPtrack)."""
        d = self.ctx.dev_alloc(max(self.n * 16, 16))
        if self.n:
            self.ctx._ck(self.lib.gzb_d2d(self.ctx._h, C.c_void_p(d.ptr), self.P("np"), self.n * 16))
        return d

    def free(self):
        if self.d_all is not None:
            self.d_all.free()
            self.d_all = None
        if self.d_field is not None:
            self.d_field.free()
            self.d_field = None


class LazyXNP:
    """`XNPtrack` (gonzag/mod2sat.py:177-186: a (Nj, Ni) map nobody may read) is computed on the GPU the first time
    the attribute is read: scatter of the track indices (last writer wins == largest index), land = -100, untouched
    ocean masked.  Holds the nearest points on the device, and the grid context when the constructor created it."""

    def __init__(self, ctx, own_ctx, d_np, nt):
        self.ctx, self.own_ctx, self.d_np, self.nt = ctx, own_ctx, d_np, int(nt)
        self.value = None

    def get(self):
        if self.value is None:
            ctx = self.ctx
            if ctx is None or ctx._h is None:
                raise L.GonzagB200Error(L.GZB_ERR_STATE, "XNPtrack: the grid context of this result has been closed")
            lib, h = L.lib(), ctx._h
            Nj, Ni = ctx.Nj, ctx.Ni
            rmiss = float(config.rmissval)
            d_xnp = ctx.dev_alloc(Nj * Ni * 9)          # the map (f64) followed by its missing-value mask (u8)
            d_xmk = C.c_void_p(d_xnp.ptr + Nj * Ni * 8)
            ctx._ck(lib.gzb_xnp_track_masked(h, self.nt, C.c_void_p(self.d_np.ptr), rmiss, C.c_void_p(d_xnp.ptr), d_xmk,
                                             L.GZB_DEVICE))
            # ONE transfer for the map and its mask (they sit back to back on the device)
            xbuf = np.empty(Nj * Ni * 9, dtype=np.uint8)
            ctx._ck(lib.gzb_d2h(h, L.ptr(xbuf), C.c_void_p(d_xnp.ptr), xbuf.nbytes))
            ctx.sync()
            d_xnp.free()
            xnp = xbuf[:Nj * Ni * 8].view(np.float64).reshape(Nj, Ni)
            xmk = xbuf[Nj * Ni * 8:].view(np.bool_).reshape(Nj, Ni)
            X = xnp.view(np.ma.MaskedArray)              # masked_where(XNPtrack == rmissval, XNPtrack), mod2sat.py:186,
            X._mask = xmk                                # with the comparison done on the GPU
            X._sharedmask = False
            self.value = X
            self.release()
        return self.value

    def release(self):
        if self.d_np is not None:
            self.d_np.free()
            self.d_np = None
        if self.own_ctx and self.ctx is not None:
            self.ctx.close()
        self.ctx = None


def wrap_products(obj, vssh_m_np, vssh_m_bl, vssh_s, imask, bad, vdistance):
    """The reference's output conventions (gonzag/mod2sat.py:157-168) on arrays the constructor owns: no copy of
    the data, one mask array per product."""
    obj.mask = imask
    obj.ssh_mod_np = np.ma.MaskedArray(vssh_m_np, mask=bad.copy(), copy=False)
    obj.ssh_mod = np.ma.MaskedArray(vssh_m_bl, mask=bad.copy(), copy=False)
    obj.ssh_sat = np.ma.MaskedArray(vssh_s, mask=bad, copy=False)
    obj.distance = vdistance


class _LazyXNPMixin:
    # `XNPtrack` exists only when config.l_save_track_on_model_grid was set (gonzag/mod2sat.py:177)
    def __getattr__(self, name):
        if name == "XNPtrack":
            lazy = self.__dict__.get("_xnp")
            if lazy is not None:
                return lazy.get()
        raise AttributeError(name)

    def release(self):
        """Free what a pending `XNPtrack` keeps on the GPU (after this the attribute can no longer be computed)."""
        lazy = self.__dict__.get("_xnp")
        if lazy is not None and lazy.value is None:
            lazy.release()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Model2SatTrack(_LazyXNPMixin):

    def __init__(self, MG, name_ssh_mod, ST, name_ssh_sat, device=0, ctx=None, keep_mapping=True,
                 slab_bytes=256 << 20, resident_bytes=2 << 30, mapping=None, eager_xnp=False):
        """Same four arguments as the reference.  Extras: ``ctx`` (a `GridContext` to reuse),
        ``keep_mapping`` (copy NP / SM / WB back as ``self.BT``), ``mapping`` (an object with
        ``NP, SM, WB`` of this very grid + track pair, e.g. ``previous.BT``: K1-K3 are skipped and
        only the interpolation runs -- another model variable onto the same track), ``eager_xnp`` (compute the
        ``XNPtrack`` map in the constructor, as the reference does, instead of on first access)."""
        (Nj, Ni) = MG.shape
        d_found_km = config.rfactor * MG.HResDeg * config.deg2km                    # mod2sat.py:35
        np_box_radius = SearchBoxSize(MG.HResDeg * config.deg2km, config.search_box_w_km)  # :39
        Nt = ST.size

        tim = {}
        t0 = time.time()
        tc = time.perf_counter()
        if ctx is None:
            ctx = getattr(MG, "ctx", None)          # a gonzag_b200.ModGrid keeps its grid on the GPU
        own_ctx = ctx is None
        if ctx is None:
            ang = MG.xangle if np.shape(getattr(MG, "xangle", ())) == (Nj, Ni) else None
            ctx = GridContext(MG.lat, MG.lon, angle=ang, mask=MG.mask, k_ew_per=MG.EWPer, device=device)
        elif not ctx.has_mask:
            ctx.set_mask(MG.mask)
        tim["context_s"] = time.perf_counter() - tc
        tc = time.perf_counter()

        # ---- one device allocation for the whole track, asynchronous uploads, K1-K4       (mod2sat.py:49-139)
        job = TrackJob(ctx, L.as_f64(ST.lat), L.as_f64(ST.lon), L.as_f64(ST.time), L.as_f64(MG.time))
        self._xnp = None
        try:
            r = int(np_box_radius)
            job.run(MG, name_ssh_mod, d_found_km, r, (r, r), mapping=mapping, slab_bytes=slab_bytes,
                    resident_bytes=resident_bytes)
            # the satellite series is read on the host while the GPU works                 (mod2sat.py:151)
            vssh_s = satellite_ssh(ST, name_ssh_sat)
            if vssh_s.shape != (Nt,):
                raise ValueError("the satellite series has %s values, the track %d points" % (vssh_s.shape, Nt))
            job.distance_and_mask(vssh_s)
            tim["issue_s"] = time.perf_counter() - tc
            vssh_m_np, vssh_m_bl, vdistance, imask, bad = job.download_series()
            if keep_mapping and mapping is None:
                self.BT = job.download_mapping()
            elif keep_mapping:
                self.BT = mapping
            ctx.sync()
            time_bl_mapping = time_bl_interp = time.time() - t0
            tim["mapping_s"] = tim["interp_s"] = 0.5 * (time.perf_counter() - tc)
            tc = time.perf_counter()
            if config.l_save_track_on_model_grid:
                self._xnp = LazyXNP(ctx, own_ctx, job.detach_np(), Nt)
                own_ctx = False                          # the context now lives as long as the pending map
            self.fused_path = job.fused
        finally:
            job.free()
        tim["outputs_s"] = time.perf_counter() - tc
        tc = time.perf_counter()

        # ---- the output conventions (API packaging, host)                       (mod2sat.py:153-168)
        self.time = ST.time
        self.lat = ST.lat
        self.lon = ST.lon
        wrap_products(self, vssh_m_np, vssh_m_bl, vssh_s, imask, bad, vdistance)
        if eager_xnp and self._xnp is not None:
            self.stats = ctx.stats()
            self._xnp.get()
        if ctx._h is not None:
            self.stats = ctx.stats()
        tim["packaging_s"] = time.perf_counter() - tc
        self.timing = tim
        self.time_bl_mapping = time_bl_mapping
        self.time_bl_interp = time_bl_interp
        if config.ivrb > 0:
            print('\n *** Time report:')
            print('     - Construction of the source-target bilinear mapping took: ' + str(round(time_bl_mapping, 3)) + ' s')
            print('     - Interpolation of model data on the ' + str(Nt) + ' satellite points took: '
                  + str(round(time_bl_interp, 3)) + ' s \n')
        if own_ctx:
            tc = time.perf_counter()
            ctx.close()
            tim["close_s"] = time.perf_counter() - tc
