"""Drop-in for the reference's gonzag/mod2sat.py (`Model2SatTrack`, gonzag/mod2sat.py:18-186).

Same constructor, same attributes (`time, lat, lon, mask, ssh_mod_np, ssh_mod, ssh_sat, distance,
XNPtrack`), but the track stays on the GPU from the mapping to the interpolated series: the host
loops only over *slabs of model records*, never over track points.

Where the model records and the satellite SSH come from, in order of preference:
  * ``MG.field``      (nrec, Nj, Ni) array in memory  /  ``MG.get_record(k)`` callable,
    ``ST.ssh``        (Nt,) array in memory;
  * otherwise the reference's own readers ``GetModel2DVar`` / ``GetSatSSH`` from
    ``gonzag.ncio`` / ``gonzag.zarrio`` (gonzag/mod2sat.py:27-31) when that package is importable,
    which is the drop-in situation inside ``alongtrack_sat_vs_nemo.py``.
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
        raise RuntimeError("no in-memory field provider (MG.field / MG.get_record, ST.ssh) and the "
                           "reference's I/O readers are not importable: %s" % exc)


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


class Model2SatTrack:

    def __init__(self, MG, name_ssh_mod, ST, name_ssh_sat, device=0, ctx=None, keep_mapping=True,
                 slab_bytes=256 << 20):
        (Nj, Ni) = MG.shape
        d_found_km = config.rfactor * MG.HResDeg * config.deg2km                    # mod2sat.py:35
        np_box_radius = SearchBoxSize(MG.HResDeg * config.deg2km, config.search_box_w_km)  # :39
        Nt = ST.size
        lib = L.lib()

        own_ctx = ctx is None
        tim = {}
        t0 = time.time()
        tc = time.perf_counter()
        if ctx is None:
            ang = MG.xangle if np.shape(getattr(MG, "xangle", ())) == (Nj, Ni) else None
            ctx = GridContext(MG.lat, MG.lon, angle=ang, mask=MG.mask, k_ew_per=MG.EWPer, device=device)
        elif not ctx.has_mask:
            ctx.set_mask(MG.mask)
        h = ctx._h
        ck = ctx._ck
        tim["context_s"] = time.perf_counter() - tc
        tc = time.perf_counter()

        # ---- track to the device, once
        yt = L.as_f64(ST.lat)
        xt = L.as_f64(ST.lon)
        tt = L.as_f64(ST.time)
        tm = L.as_f64(MG.time)
        d_y, d_x, d_t, d_tm = ctx.upload(yt), ctx.upload(xt), ctx.upload(tt), ctx.upload(tm)
        d_np = ctx.dev_alloc(Nt * 16)
        d_sm = ctx.dev_alloc(Nt * 64)
        d_wb = ctx.dev_alloc(Nt * 32)
        d_vnp = ctx.dev_alloc(Nt * 8)
        d_vbl = ctx.dev_alloc(Nt * 8)
        P = lambda b: C.c_void_p(b.ptr)

        # ---- the mapping: K1 + K2 + K3                                          (mod2sat.py:49-50)
        ck(lib.gzb_mapping(h, Nt, P(d_y), P(d_x), float(d_found_km), int(np_box_radius),
                           int(np_box_radius), int(np_box_radius), P(d_np), P(d_sm), P(d_wb), L.GZB_DEVICE))
        ctx.sync()
        time_bl_mapping = time.time() - t0
        tim["mapping_s"] = time.perf_counter() - tc
        tc = time.perf_counter()

        # ---- space-time interpolation: K4 over slabs of records                 (mod2sat.py:86-139)
        t0 = time.time()
        field = getattr(MG, "field", None)
        get_record = getattr(MG, "get_record", None)
        if field is None and get_record is None:
            reader, _ = _reference_readers()
            get_record = lambda k: reader(MG.file, name_ssh_mod, kt=k)
        first = True
        if Nt > 0:
            if field is not None:
                field = np.ma.getdata(field)
                if not field.flags["C_CONTIGUOUS"]:
                    field = np.ascontiguousarray(field)
                dt = GridContext._field_dtype(field)
                rec_bytes = field[0].nbytes
                for (ka, kb) in plan_slabs(tt, tm, max(2, int(1 << 40))):
                    slab = field[ka:kb + 1]
                    ck(lib.gzb_interp(h, Nt, P(d_t), P(d_sm), P(d_wb), tm.shape[0], P(d_tm), L.ptr(slab), dt,
                                      ka, kb + 1 - ka, float(config.rmissval), 1 if first else 0,
                                      P(d_vnp), P(d_vbl), L.GZB_DEVICE))
                    first = False
            else:
                probe = np.ma.getdata(get_record(int(np.searchsorted(tm, tt[0], side="right") - 1)))
                dt = GridContext._field_dtype(probe)
                rec_bytes = probe.nbytes
                per_slab = max(2, int(slab_bytes // rec_bytes))
                stage = L.pinned_empty((per_slab,) + probe.shape, probe.dtype)
                try:
                    for (ka, kb) in plan_slabs(tt, tm, per_slab):
                        for k in range(ka, kb + 1):
                            stage[k - ka] = np.ma.getdata(get_record(k))
                        ck(lib.gzb_interp(h, Nt, P(d_t), P(d_sm), P(d_wb), tm.shape[0], P(d_tm), L.ptr(stage), dt,
                                          ka, kb + 1 - ka, float(config.rmissval), 1 if first else 0,
                                          P(d_vnp), P(d_vbl), L.GZB_DEVICE))
                        first = False
                finally:
                    L.pinned_free(stage)
        if first and Nt > 0:   # no slab at all cannot happen (plan_slabs raises), kept for safety
            raise RuntimeError("no model record brackets the track")

        ctx.sync()
        tim["interp_s"] = time.perf_counter() - tc
        tc = time.perf_counter()
        # ---- distance and nearest-point map                                      (mod2sat.py:146-148, 177-184)
        d_dist = ctx.dev_alloc(Nt * 8)
        ck(lib.gzb_track_distance(h, Nt, P(d_y), P(d_x), P(d_dist), L.GZB_DEVICE))
        xnp = None
        if config.l_save_track_on_model_grid:
            d_xnp = ctx.dev_alloc(Nj * Ni * 8)
            ck(lib.gzb_xnp_track(h, Nt, P(d_np), float(config.rmissval), P(d_xnp), L.GZB_DEVICE))
            xnp = ctx.download(d_xnp, (Nj, Ni), np.float64)
            d_xnp.free()
        vssh_m_np = ctx.download(d_vnp, (Nt,), np.float64)
        vssh_m_bl = ctx.download(d_vbl, (Nt,), np.float64)
        vdistance = ctx.download(d_dist, (Nt,), np.float64)
        time_bl_interp = time.time() - t0
        if keep_mapping:
            self.BT = MappingView(ctx.download(d_np, (Nt, 2), np.int64), ctx.download(d_sm, (Nt, 4, 2), np.int64),
                                  ctx.download(d_wb, (Nt, 4), np.float64))
        for b in (d_y, d_x, d_t, d_tm, d_np, d_sm, d_wb, d_vnp, d_vbl, d_dist):
            b.free()
        tim["outputs_s"] = time.perf_counter() - tc
        tc = time.perf_counter()

        # ---- satellite SSH and the output conventions (API packaging, host)      (mod2sat.py:151-168)
        ssh_attr = getattr(ST, "ssh", None)
        if ssh_attr is not None:
            vssh_s = np.array(ssh_attr, dtype=np.float64, copy=True)
        else:
            _, sat_reader = _reference_readers()
            vssh_s = sat_reader(ST.file, name_ssh_sat, kt1=ST.jt1, kt2=ST.jt2, ikeep=ST.keepit)
        self.time = ST.time
        self.lat = ST.lat
        self.lon = ST.lon
        imask = np.zeros(Nt, dtype=np.int8)
        imask[np.where((vssh_m_bl > -100.) & (vssh_m_bl < 100.) & (vssh_s > -100.) & (vssh_s < 100.))] = 1
        self.mask = imask
        self.ssh_mod_np = np.ma.masked_where(imask == 0, vssh_m_np)
        self.ssh_mod = np.ma.masked_where(imask == 0, vssh_m_bl)
        self.ssh_sat = np.ma.masked_where(imask == 0, vssh_s)
        self.distance = vdistance
        if xnp is not None:
            self.XNPtrack = np.ma.MaskedArray(xnp, mask=(xnp == config.rmissval), copy=False)
        tim["packaging_s"] = time.perf_counter() - tc
        self.timing = tim
        self.time_bl_mapping = time_bl_mapping
        self.time_bl_interp = time_bl_interp
        self.stats = ctx.stats()
        if config.ivrb > 0:
            print('\n *** Time report:')
            print('     - Construction of the source-target bilinear mapping took: ' + str(round(time_bl_mapping, 3)) + ' s')
            print('     - Interpolation of model data on the ' + str(Nt) + ' satellite points took: '
                  + str(round(time_bl_interp, 3)) + ' s \n')
        if own_ctx:
            ctx.close()
