"""Drop-in for the reference's gonzag/mod2sat.py (`Model2SatTrack`, gonzag/mod2sat.py:18-186).

Same constructor, same attributes (`time, lat, lon, mask, ssh_mod_np, ssh_mod, ssh_sat, distance,
XNPtrack`), but the track stays on the GPU from the mapping to the interpolated series: the host
loops only over *slabs of model records*, never over track points.

Where the model records and the satellite SSH come from, in order of preference:
  * ``MG.field``      (nrec, Nj, Ni) array in memory  /  ``MG.get_record(k)`` callable,
    ``ST.ssh``        (Nt,) array in memory;
  * ``MG.source`` / ``ST.source``: the in-memory sources of `gonzag_b200.sources` (what
    `gonzag_b200.ModGrid` / `SatTrack` were built from), whole array or one record at a time;
  * otherwise the reference's own readers ``GetModel2DVar`` / ``GetSatSSH`` from
    ``gonzag.ncio`` / ``gonzag.zarrio`` (gonzag/mod2sat.py:27-31) when that package is importable,
    which is the drop-in situation inside ``alongtrack_sat_vs_nemo.py``.

Record index convention, kept from the reference: record ``k`` is the one whose time is
``MG.time[k]`` *as the reader numbers it*, i.e. the reference asks its reader for ``kt = k`` with
``k`` an index into the (possibly sliced) ``MG.time`` (gonzag/mod2sat.py:94-112).

How the work is issued:
  * a field held as ONE array that the GPU can read (pinned + mapped host memory: gathered in
    place over PCIe; or small enough to be copied to HBM first) goes through `gzb_map_interp`:
    K1-K4 in one call, the chain replays of K1 beside K2-K4;
  * anything else (pageable big arrays, per-record providers) goes through `gzb_mapping` and then
    `gzb_interp` slab by slab, records staged through pinned memory, copies overlapped with K4.
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
    """(whole field array or None, per-record callable or None) for the model variable."""
    field = getattr(MG, "field", None)
    get_record = getattr(MG, "get_record", None)
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
    return field, get_record


def _record_writer(MG):
    """`into(name, k, out) -> bool` of the model source when it can write a record straight into a staging slot
    (`NetCDF3Source.record_into`), else None."""
    if getattr(MG, "field", None) is not None or getattr(MG, "get_record", None) is not None:
        return None
    return getattr(getattr(MG, "source", None), "record_into", None)


class Model2SatTrack:

    def __init__(self, MG, name_ssh_mod, ST, name_ssh_sat, device=0, ctx=None, keep_mapping=True,
                 slab_bytes=256 << 20, resident_bytes=2 << 30, mapping=None):
        """Same four arguments as the reference.  Extras: ``ctx`` (a `GridContext` to reuse),
        ``keep_mapping`` (copy NP / SM / WB back as ``self.BT``), ``mapping`` (an object with
        ``NP, SM, WB`` of this very grid + track pair, e.g. ``previous.BT``: K1-K3 are skipped and
        only the interpolation runs -- another model variable onto the same track)."""
        (Nj, Ni) = MG.shape
        d_found_km = config.rfactor * MG.HResDeg * config.deg2km                    # mod2sat.py:35
        np_box_radius = SearchBoxSize(MG.HResDeg * config.deg2km, config.search_box_w_km)  # :39
        Nt = ST.size
        lib = L.lib()

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
        h = ctx._h
        ck = ctx._ck
        tim["context_s"] = time.perf_counter() - tc
        tc = time.perf_counter()

        # ---- one device allocation for the whole track, asynchronous uploads
        yt = L.as_f64(ST.lat)
        xt = L.as_f64(ST.lon)
        tt = L.as_f64(ST.time)
        tm = L.as_f64(MG.time)
        nrec = tm.shape[0]
        al = lambda n: (int(n) + 255) & ~255
        sizes = [("y", Nt * 8), ("x", Nt * 8), ("t", Nt * 8), ("tm", nrec * 8), ("np", Nt * 16), ("sm", Nt * 64),
                 ("wb", Nt * 32), ("vnp", Nt * 8), ("vbl", Nt * 8), ("dist", Nt * 8)]
        off, total = {}, 0
        for k, n in sizes:
            off[k] = total
            total += al(max(n, 8))
        d_all = ctx.dev_alloc(total)
        P = lambda k: C.c_void_p(d_all.ptr + off[k])
        for k, a in (("y", yt), ("x", xt), ("t", tt), ("tm", tm)):
            if a.nbytes:
                ck(lib.gzb_h2d(h, P(k), L.ptr(a), a.nbytes))

        field, get_record = _providers(MG, name_ssh_mod)
        rmiss = float(config.rmissval)
        r = int(np_box_radius)
        fused = False
        d_field = None
        if mapping is not None:
            mNP, mSM, mWB = L.as_i64(mapping.NP), L.as_i64(mapping.SM), L.as_f64(mapping.WB)
            if mNP.shape != (Nt, 2) or mSM.shape != (Nt, 4, 2) or mWB.shape != (Nt, 4):
                raise ValueError("`mapping` does not belong to this track (Nt = %d)" % Nt)
            for k, a in (("np", mNP), ("sm", mSM), ("wb", mWB)):
                if a.nbytes:
                    ck(lib.gzb_h2d(h, P(k), L.ptr(a), a.nbytes))
        if field is not None and Nt > 0 and mapping is None:
            field = np.ma.getdata(field)
            if not field.flags["C_CONTIGUOUS"]:
                field = np.ascontiguousarray(field)
            dt = GridContext._field_dtype(field)
            # the reference asks its reader for record k = index into the (possibly sliced) MG.time
            # (gonzag/mod2sat.py:94-112): a window starting after record 0 still reads records 0 .. nrec-1
            if field.shape[0] < nrec:
                raise ValueError("MG.field holds %d records, MG.time %d" % (field.shape[0], nrec))
            field = field[:nrec]
            kind = L.pointer_kind(field)
            sparse = Nt * 1024.0 < field.nbytes          # same rule as gzb_interp's zero-copy gather
            if kind >= 1 and (kind == 2 or sparse):
                slab_ptr = L.ptr(field)
                fused = True
            elif field.nbytes <= resident_bytes:
                d_field = ctx.dev_alloc(field.nbytes)    # small field: copy it to HBM once
                ck(lib.gzb_h2d(h, C.c_void_p(d_field.ptr), L.ptr(field), field.nbytes))
                slab_ptr = C.c_void_p(d_field.ptr)
                fused = True

        if fused:
            # ---- K1-K4 in one call                                              (mod2sat.py:49-139)
            ck(lib.gzb_map_interp(h, Nt, P("y"), P("x"), P("t"), float(d_found_km), r, r, r, nrec, P("tm"), slab_ptr,
                                  dt, 0, nrec, rmiss, P("np"), P("sm"), P("wb"), P("vnp"), P("vbl")))
            ctx.sync()
            tim["mapping_s"] = tim["interp_s"] = 0.5 * (time.perf_counter() - tc)
            time_bl_mapping = time_bl_interp = time.time() - t0
        else:
            # ---- the mapping: K1 + K2 + K3                                      (mod2sat.py:49-50)
            if Nt > 0 and mapping is None:
                ck(lib.gzb_mapping(h, Nt, P("y"), P("x"), float(d_found_km), r, r, r, P("np"), P("sm"), P("wb"),
                                   L.GZB_DEVICE))
            ctx.sync()
            time_bl_mapping = time.time() - t0
            tim["mapping_s"] = time.perf_counter() - tc
            tc = time.perf_counter()
            # ---- space-time interpolation: K4 over slabs of records             (mod2sat.py:86-139)
            t0 = time.time()
            first = True
            if Nt > 0:
                if field is not None:
                    field = np.ma.getdata(field)
                    if not field.flags["C_CONTIGUOUS"]:
                        field = np.ascontiguousarray(field)
                    dt = GridContext._field_dtype(field)
                    for (ka, kb) in plan_slabs(tt, tm, max(2, int(1 << 40))):
                        slab = field[ka:kb + 1]
                        ck(lib.gzb_interp(h, Nt, P("t"), P("sm"), P("wb"), nrec, P("tm"), L.ptr(slab), dt,
                                          ka, kb + 1 - ka, rmiss, 1 if first else 0, P("vnp"), P("vbl"), L.GZB_DEVICE))
                        first = False
                else:
                    probe = np.ma.getdata(get_record(int(np.searchsorted(tm, tt[0], side="right") - 1)))
                    dt = GridContext._field_dtype(probe)
                    per_slab = max(2, int(slab_bytes // probe.nbytes))
                    stage = L.pinned_empty((per_slab,) + probe.shape, probe.dtype)
                    into = _record_writer(MG)
                    try:
                        for (ka, kb) in plan_slabs(tt, tm, per_slab):
                            for k in range(ka, kb + 1):
                                if into is None or not into(name_ssh_mod, k, stage[k - ka]):
                                    stage[k - ka] = np.ma.getdata(get_record(k))
                            ck(lib.gzb_interp(h, Nt, P("t"), P("sm"), P("wb"), nrec, P("tm"), L.ptr(stage), dt,
                                              ka, kb + 1 - ka, rmiss, 1 if first else 0, P("vnp"), P("vbl"),
                                              L.GZB_DEVICE))
                            first = False
                    finally:
                        L.pinned_free(stage)
            ctx.sync()
            tim["interp_s"] = time.perf_counter() - tc
            time_bl_interp = time.time() - t0
        tc = time.perf_counter()

        # ---- distance and nearest-point map                                      (mod2sat.py:146-148, 177-184)
        vssh_m_np = np.empty(Nt)
        vssh_m_bl = np.empty(Nt)
        vdistance = np.empty(Nt)
        xnp = None
        if Nt > 0:
            ck(lib.gzb_track_distance(h, Nt, P("y"), P("x"), P("dist"), L.GZB_DEVICE))
            ck(lib.gzb_d2h(h, L.ptr(vssh_m_np), P("vnp"), Nt * 8))
            ck(lib.gzb_d2h(h, L.ptr(vssh_m_bl), P("vbl"), Nt * 8))
            ck(lib.gzb_d2h(h, L.ptr(vdistance), P("dist"), Nt * 8))
        tim["out_series_s"] = time.perf_counter() - tc
        if config.l_save_track_on_model_grid:
            d_xnp = ctx.dev_alloc(Nj * Ni * 9)          # the map (f64) followed by its missing-value mask (u8)
            d_xmk = C.c_void_p(d_xnp.ptr + Nj * Ni * 8)
            ck(lib.gzb_xnp_track_masked(h, Nt, P("np"), rmiss, C.c_void_p(d_xnp.ptr), d_xmk, L.GZB_DEVICE))
            # ONE transfer for the map and its mask (they sit back to back on the device): the mask alone (Nj*Ni bytes)
            # is below the size from which a pageable transfer is cut over the copy threads, and crawled at 4 GB/s
            xbuf = np.empty(Nj * Ni * 9, dtype=np.uint8)
            xnp = xbuf[:Nj * Ni * 8].view(np.float64).reshape(Nj, Ni)
            xmk = xbuf[Nj * Ni * 8:].view(np.bool_).reshape(Nj, Ni)
            tim["out_xnp_kernel_s"] = time.perf_counter() - tc
            ck(lib.gzb_d2h(h, L.ptr(xbuf), C.c_void_p(d_xnp.ptr), xbuf.nbytes))
            tim["out_xnp_d2h_s"] = time.perf_counter() - tc
        if keep_mapping:
            # NP, SM, WB sit one after the other on the device: one transfer (big enough for the copy threads)
            mbuf = np.empty(off["wb"] + Nt * 32 - off["np"], dtype=np.uint8)
            part = lambda k, nbytes, dt, shape: mbuf[off[k] - off["np"]:off[k] - off["np"] + nbytes].view(dt).reshape(shape)
            NP, SM, WB = (part("np", Nt * 16, np.int64, (Nt, 2)), part("sm", Nt * 64, np.int64, (Nt, 4, 2)),
                          part("wb", Nt * 32, np.float64, (Nt, 4)))
            if Nt > 0:
                ck(lib.gzb_d2h(h, L.ptr(mbuf), P("np"), mbuf.nbytes))
            self.BT = MappingView(NP, SM, WB)
        ctx.sync()
        tim["out_sync_s"] = time.perf_counter() - tc
        d_all.free()
        if d_field is not None:
            d_field.free()
        if config.l_save_track_on_model_grid:
            d_xnp.free()
        tim["outputs_s"] = time.perf_counter() - tc
        tc = time.perf_counter()

        # ---- satellite SSH and the output conventions (API packaging, host)      (mod2sat.py:151-168)
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
        # missing value first so that `good` below is a plain ndarray too
        if isinstance(vssh_s, np.ma.MaskedArray):
            vssh_s = np.ma.filled(np.ma.asarray(vssh_s, dtype=np.float64), rmiss)
        self.time = ST.time
        self.lat = ST.lat
        self.lon = ST.lon
        # -100 < x < 100 for both series (mod2sat.py:157-159) written as |x| < 100: same truth table (NaN and inf fail
        # either way), five passes over the track instead of seven
        good = (np.abs(vssh_m_bl) < 100.) & (np.abs(vssh_s) < 100.)
        self.mask = good.astype(np.int8)
        bad = ~good
        # masked_where(imask == 0, x) on arrays this constructor owns: no copy of the data, one mask each
        self.ssh_mod_np = np.ma.MaskedArray(vssh_m_np, mask=bad.copy(), copy=False)
        self.ssh_mod = np.ma.MaskedArray(vssh_m_bl, mask=bad.copy(), copy=False)
        self.ssh_sat = np.ma.MaskedArray(vssh_s, mask=bad, copy=False)
        self.distance = vdistance
        if xnp is not None:
            X = xnp.view(np.ma.MaskedArray)              # masked_where(XNPtrack == rmissval, XNPtrack), mod2sat.py:186,
            X._mask = xmk                                # with the comparison done on the GPU
            X._sharedmask = False
            self.XNPtrack = X
        tim["packaging_s"] = time.perf_counter() - tc
        self.timing = tim
        self.fused_path = fused
        self.time_bl_mapping = time_bl_mapping
        self.time_bl_interp = time_bl_interp
        self.stats = ctx.stats()
        if config.ivrb > 0:
            print('\n *** Time report:')
            print('     - Construction of the source-target bilinear mapping took: ' + str(round(time_bl_mapping, 3)) + ' s')
            print('     - Interpolation of model data on the ' + str(Nt) + ' satellite points took: '
                  + str(round(time_bl_interp, 3)) + ' s \n')
        if own_ctx:
            tc = time.perf_counter()
            ctx.close()
            tim["close_s"] = time.perf_counter() - tc
