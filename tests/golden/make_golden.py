#!/usr/bin/env python3
"""Generate the golden vectors in this directory by running the UNMODIFIED reference.

Run it only where the reference checkout is mounted (the build container):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [/root/reference]

The reference package is imported as-is from that path.  Two things it needs are not
installed offline and are injected through ``sys.modules`` before the import:

* ``gonzag.ncio`` (netCDF4 readers) -> an in-memory stand-in exposing `GetModel2DVar` and
  `GetSatSSH`, so that `Model2SatTrack` (gonzag/mod2sat.py:20) runs without files;
* ``shapely.geometry`` -> a planar strict-interior point-in-polygon stand-in (only the
  heavy-duty quadrant branch, gonzag/bilin_mapping.py:415-426, reaches it).  That branch
  is therefore "parity unpinned" (see oracle/gz_oracle.py header).

Inputs are stored in the fixtures next to the outputs (fp32-born coordinates are stored
as float32, which is lossless) so the fixtures do not depend on the generator staying
bit-stable across machines.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gonzag_b200 import synthetic as syn          # noqa: E402


# ----------------------------------------------------------------- reference loading
from oracle.ref_loader import load_reference      # noqa: E402  (shapely + ncio stand-ins, then `import gonzag`)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def run_model2sat(gz, store, lat, lon, mask, angle, kew, tmod, field, tsat, ysat, xsat, ssat, res):
    store["fields"]["mod"] = field
    store["sat"]["sat"] = ssat
    MG = types.SimpleNamespace(shape=lat.shape, HResDeg=res, lat=lat, lon=lon, xangle=angle,
                               EWPer=kew, time=tmod, file="mod", mask=mask)
    ST = types.SimpleNamespace(size=len(tsat), lat=ysat, lon=xsat, time=tsat, file="sat",
                               jt1=0, jt2=0, keepit=[])
    with quiet():
        R = gz.Model2SatTrack(MG, "ssh", ST, "ssh")
    return R


def pack_masked(prefix, arr, out):
    out[prefix + "_data"] = np.ma.getdata(arr)
    out[prefix + "_mask"] = np.ma.getmaskarray(arr)


# ----------------------------------------------------------------- cases
def sat_ssh(n, seed):
    rng = np.random.default_rng(seed)
    s = rng.normal(0.0, 0.3, n)
    s[rng.uniform(size=n) < 0.03] = -9999.0
    return s


def case_global(gz, store):
    """C1-shaped global periodic grid, -D angle, land mask, SARAL-like track (2 orbits)."""
    lat, lon = syn.orca_like_grid(292, 362)
    mask = syn.land_mask(lat, lon)
    with quiet():
        angle = gz.GridAngle(lat, lon)
    res = float(gz.GridResolution(lon))
    t, y, x = syn.orbit_track(12070, t0=1000.0, **syn.SARAL)
    tmod = 7000.0 * np.arange(3)
    f32 = syn.ssh_records(lat, lon, 3, dtype=np.float32)
    f64 = f32.astype(np.float64) + lat[None] * 2.0 ** -20
    ssat = sat_ssh(len(t), 11)
    out = dict(lat=lat.astype(np.float32), lon=lon.astype(np.float32), mask=mask.astype(np.int8),
               angle=angle, res=res, kew=2, tmod=tmod, f32=f32, tsat=t, ysat=y, xsat=x, ssat=ssat)
    for tag, fld in (("f32", f32), ("f64", f64)):
        R = run_model2sat(gz, store, lat, lon, mask, angle, 2, tmod, fld, t, y, x, ssat.copy(), res)
        pack_masked("ssh_np_" + tag, R.ssh_mod_np, out)
        pack_masked("ssh_bl_" + tag, R.ssh_mod, out)
        out["imask_" + tag] = R.mask
    pack_masked("ssh_sat", R.ssh_sat, out)
    out["distance"] = R.distance
    pack_masked("xnp", R.XNPtrack, out)
    with quiet():
        BT = gz.BilinTrack(y, x, lat, lon, src_grid_local_angle=angle, k_ew_per=2,
                           rd_found_km=0.75 * res * 111.11, np_box_r=int(0.5 * 500. / (res * 111.11)))
    out.update(NP=BT.NP, SM=BT.SM, WB=BT.WB)
    # same track without -D and without periodicity knowledge
    with quiet():
        BT0 = gz.BilinTrack(y, x, lat, lon, k_ew_per=-1,
                            rd_found_km=0.75 * res * 111.11, np_box_r=int(0.5 * 500. / (res * 111.11)))
    out.update(NP_noper=BT0.NP, SM_noper=BT0.SM, WB_noper=BT0.WB)
    return out


def case_regional(gz, store):
    """1/12-degree sheared regional box (r = 27), SARAL passes over it, bbox slightly larger
    than the grid so some points are outside (-> not found)."""
    lat, lon = syn.regional_grid(300, 360, lat0=-40.0, lon0=150.0, res_deg=1.0 / 12.0, shear=0.02)
    mask = syn.land_mask(lat, lon, seed=5, level=0.9)
    res = float(gz.GridResolution(lon))
    t, y, x = syn.orbit_track(5 * 86400, t0=50.0, drop_land_seed=None,
                              lat_bounds=(lat.min() - 0.4, lat.max() + 0.4),
                              lon_bounds=(lon.min() - 0.4, lon.max() + 0.4), **syn.SARAL)
    rd = 0.75 * res * 111.11
    r = int(0.5 * 500. / (res * 111.11))
    with quiet():
        BT = gz.BilinTrack(y, x, lat, lon, k_ew_per=-1, rd_found_km=rd, np_box_r=r)
    tmod = 2.6 * 86400.0 * np.arange(3)
    f32 = syn.ssh_records(lat, lon, 3, dtype=np.float32)
    ssat = sat_ssh(len(t), 12)
    R = run_model2sat(gz, store, lat, lon, mask, np.zeros(lat.shape), -1, tmod, f32, t, y, x,
                      ssat.copy(), res)
    out = dict(lat=lat.astype(np.float32), lon=lon.astype(np.float32), mask=mask.astype(np.int8),
               res=res, rd=rd, r=r, tmod=tmod, f32=f32, tsat=t, ysat=y, xsat=x, ssat=ssat,
               NP=BT.NP, SM=BT.SM, WB=BT.WB, distance=R.distance, imask=R.mask)
    pack_masked("ssh_np", R.ssh_mod_np, out)
    pack_masked("ssh_bl", R.ssh_mod, out)
    # shuffled targets: no continuity between consecutive points
    perm = np.random.default_rng(4).permutation(len(y))[:400]
    with quiet():
        BTs = gz.BilinTrack(y[perm], x[perm], lat, lon, k_ew_per=-1, rd_found_km=rd, np_box_r=r)
    out.update(perm=perm, NP_perm=BTs.NP, SM_perm=BTs.SM, WB_perm=BTs.WB)
    return out


def case_seam(gz, store):
    """Tracks hugging the E-W seam and the grid edges of the C1-shaped global grid: long
    chains where the box answer differs from the global answer."""
    lat, lon = syn.orca_like_grid(292, 362)
    res = float(gz.GridResolution(lon))
    rd = 0.75 * res * 111.11
    r = int(0.5 * 500. / (res * 111.11))
    rng = np.random.default_rng(21)
    n = 1500
    s = np.linspace(0.0, 1.0, n)
    # slow drift across the seam while moving north, then back south, then a polar leg
    y = np.concatenate([-70.0 + 150.0 * s, 80.0 - 150.0 * s, 60.0 + 29.5 * s])
    x = np.concatenate([357.0 + 5.0 * s, 2.5 - 5.5 * s, 200.0 + 170.0 * s])
    y = y + rng.normal(0, 0.03, y.size)
    x = np.mod(x + rng.normal(0, 0.03, x.size), 360.0)
    with quiet():
        angle = gz.GridAngle(lat, lon)
        BT = gz.BilinTrack(y, x, lat, lon, src_grid_local_angle=angle, k_ew_per=2,
                           rd_found_km=rd, np_box_r=r)
        BTn = gz.BilinTrack(y, x, lat, lon, src_grid_local_angle=angle, k_ew_per=-1,
                            rd_found_km=rd, np_box_r=r)
    return dict(rd=rd, r=r, ysat=y, xsat=x, NP=BT.NP, SM=BT.SM, WB=BT.WB,
                NP_noper=BTn.NP, SM_noper=BTn.SM, WB_noper=BTn.WB)


def case_units(gz, store):
    """Known-answer vectors for the scalar helpers."""
    rng = np.random.default_rng(7)
    out = {}
    # Heading
    n = 2000
    a = np.stack([rng.uniform(-89, 89, n), rng.uniform(0, 360, n),
                  rng.uniform(-89, 89, n), rng.uniform(0, 360, n)], 1)
    a[:50, 2] = a[:50, 0] + rng.normal(0, 0.1, 50)       # nearby pairs
    a[:50, 3] = np.mod(a[:50, 1] + rng.normal(0, 0.1, 50), 360.)
    a[50:60, 3] = a[50:60, 1]                            # same meridian
    a[60:70, 2] = a[60:70, 0]                            # same parallel
    a[70, :] = [90.0, 10.0, 20.0, 30.0]                  # pole (epsilon clamp)
    out["heading_in"] = a
    out["heading_out"] = np.array([gz.Heading(*row) for row in a])
    # degE_to_degWE
    v = np.concatenate([rng.uniform(0, 360, 200), [0.0, 180.0, 360.0, 179.999, 180.001]])
    out["pm180_in"] = v
    out["pm180_out"] = np.array([gz.degE_to_degWE(float(q)) for q in v])
    out["pm180_out_vec"] = gz.degE_to_degWE(v.copy())
    # AlfaBeta on random quads
    m = 3000
    cy = rng.uniform(-80, 80, m)
    cx = rng.uniform(0, 360, m)
    vy = np.zeros((m, 5))
    vx = np.zeros((m, 5))
    h = rng.uniform(0.02, 1.0, m)
    for k, (dy, dx) in enumerate([(-.5, -.5), (-.5, .5), (.5, .5), (.5, -.5)]):
        vy[:, k + 1] = cy + h * (dy + rng.normal(0, 0.08, m))
        vx[:, k + 1] = cx + h * (dx + rng.normal(0, 0.08, m))
    vy[:, 0] = cy + h * rng.uniform(-.6, .6, m)
    vx[:, 0] = cx + h * rng.uniform(-.6, .6, m)
    vx[:300] = np.mod(vx[:300] - cx[:300, None], 360.)   # quads straddling Greenwich
    vy[300:320, 1:] = vy[300:320, 1:2]                   # four equal latitudes
    vx[320:330, 1:] = vx[320:330, 1:2]
    vy[320:330, 1:] = vy[320:330, 1:2]                   # fully degenerate quad (det = 0)
    ab = np.zeros((m, 2))
    for q in range(m):
        ab[q] = gz.AlfaBeta(vy[q].copy(), vx[q].copy())
    out.update(ab_vy=vy, ab_vx=vx, ab_out=ab)
    # Haversine
    glat = rng.uniform(-80, 80, (20, 30))
    glon = rng.uniform(0, 360, (20, 30))
    out["hav_glat"], out["hav_glon"] = glat, glon
    out["hav_pt"] = np.array([12.3, 321.0])
    out["hav_out"] = gz.Haversine(np.float64(12.3), np.float64(321.0), glat, glon)
    # NearestPoint quirks on a small regular grid
    lon1 = np.arange(0.0, 72.0, 1.0)
    lat1 = np.arange(-20.0, 21.0, 1.0)
    X, Y = np.meshgrid(lon1, lat1)
    out["np_Y"], out["np_X"] = Y, X
    rd = 0.75 * 111.11
    calls = []
    res = []
    for (yt, xt, jp, ip, rr) in [
            (0.3, 30.4, 20, 30, 2), (0.3, 33.0, 20, 30, 2), (0.3, 30.4, 0, 30, 2),
            (0.3, 30.4, 20, 0, 2), (0.3, 30.4, -1, -1, 2), (0.3, 30.4, None, None, 2),
            (0.3, 30.4, 5, 5, 2), (21.0, 30.2, 38, 30, 2), (20.0 + 1.332 * 0.75, 30.0, 38, 30, 2),
            (20.0 + 1.465 * 0.75, 30.0, 38, 30, 2), (20.0 + 1.2 * 0.75, 30.0, None, None, 2),
            (-20.4, 0.2, 2, 2, 2), (0.5, 30.5, 20, 30, 2), (0.0, 30.5, 20, 30, 2),
            (0.3, 71.4, 20, 69, 3), (0.3, 359.8, 20, 69, 3), (5.2, 10.7, 25, 12, 0)]:
        with quiet():
            got = gz.NearestPoint((yt, xt), Y, X, rd_found_km=rd, j_prv=jp, i_prv=ip, np_box_r=rr)
        calls.append([yt, xt, -99 if jp is None else jp, -99 if ip is None else ip, rr])
        res.append(got)
    out["np_calls"] = np.array(calls)
    out["np_out"] = np.array(res, dtype=np.int64)
    out["np_rd"] = rd
    # GridAngle on a small rotated grid
    glat, glon = syn.orca_like_grid(60, 82, pole_shift_deg=25.0)
    with quiet():
        out["ga_out"] = gz.GridAngle(glat, glon)
    out["ga_lat"], out["ga_lon"] = glat, glon
    # Haversine_sclr / RadiusEarth
    pr = np.stack([rng.uniform(-89, 89, 300), rng.uniform(0, 360, 300)], 1)
    pr2 = pr + rng.normal(0, 0.1, pr.shape)
    out["hs_in"] = np.concatenate([pr, pr2], 1)
    out["hs_out"] = np.array([gz.Haversine_sclr(a0, a1, b0, b1) for a0, a1, b0, b1 in out["hs_in"]])
    out["re_out"] = np.array([gz.RadiusEarth(a0) for a0 in pr[:, 0]])
    # IDSourceMesh / WeightBL single calls incl. forced heavy-duty (angle > 45) and a target
    # exactly due north of its nearest point (light test fails -> heavy-duty)
    sm_calls, sm_out, wb_out = [], [], []
    for (yt, xt, jP, iP, kew, ang) in [
            (0.3, 30.4, 20, 30, -1, 0.0), (0.3, 30.4, 20, 30, -1, 60.0), (0.4, 30.0, 20, 30, -1, 0.0),
            (-0.3, 29.6, 20, 30, -1, -50.0), (0.3, 29.6, 20, 30, 2, 0.0), (-0.2, 71.3, 20, 71, 2, 0.0),
            (0.2, 0.3, 20, 0, 2, 0.0), (0.0, 30.4, 20, 30, -1, 0.0)]:
        with quiet():
            sm = gz.IDSourceMesh((yt, xt), Y, X, jP, iP, k_ew_per=kew, grid_s_angle=ang)
            wb = gz.WeightBL((yt, xt), Y, X, sm)
        sm_calls.append([yt, xt, jP, iP, kew, ang])
        sm_out.append(sm)
        wb_out.append(wb)
    out["sm_calls"] = np.array(sm_calls)
    out["sm_out"] = np.array(sm_out)
    out["wb_out"] = np.array(wb_out)
    return out


def case_grid(gz, store):
    """f1 rows: the grid heuristics on the reference's own test arrays (tests/test_functions.py) and
    the UNMODIFIED `ModGrid` -> `SatTrack` -> `Model2SatTrack` sequence of alongtrack_sat_vs_nemo.py
    on three in-memory data sets (global periodic with -D; regional across Greenwich, 2-D coordinates;
    regional with 1-D coordinates and a mask from _FillValue)."""
    import tempfile
    from gonzag_b200.sources import ModelArrays, TrackArrays
    out = {}
    # ---- known-answer arrays of the reference's tests/test_functions.py
    kat = [np.array([[0., 60., 120., 180., 240., 300.]] * 3),
           np.array([[0., 60., 120., 180., 240., 300., 360.]] * 3),
           np.array([[0., 60., 120., 180., 240., 300., 360., 60.]] * 3),
           np.array([[0.5, 180., 359.5], [0.5, 180., 359.5]]),
           np.array([[340., 358., 5.], [341., 359., 6.]]),
           np.array([[170., 178., 201.], [169., 181., 200.]]),
           np.array([[299., 324., 350.], [300., 325., 352.]]),
           np.array([[30., 40., 60.], [29., 39., 59.]])]
    for k, x in enumerate(kat):
        out["kat%d_lon" % k] = x
        res = float(gz.GridResolution(x)) if k < 3 else 1.0
        g = gz.IsGlobalLongitudeWise(x, resd=res)
        ew = gz.IsEastWestPeriodic(x) if k < 3 else -99     # the reference indexes 5 columns: 3-column arrays raise
        out["kat%d_out" % k] = np.array([res, float(g[0]), float(g[1]), g[2], g[3], ew])
    # ---- scan_idx
    rng = np.random.default_rng(31)
    vt = np.cumsum(rng.uniform(0.5, 1.5, 4000)) + 1000.0
    qs = np.array([[vt[0], vt[-1]], [vt[10] + 0.2, vt[3000] + 0.1], [vt[1], vt[2]], [0.0, vt[50]],
                   [vt[100], 1e9], [vt[3999] - 0.01, vt[3999]], [vt[7], vt[7]], [-5.0, -1.0]])
    out["scan_vt"] = vt
    out["scan_q"] = qs
    out["scan_out"] = np.array([gz.scan_idx(vt, a, b) for a, b in qs], dtype=np.int64)
    # ---- the three end-to-end data sets
    tmp = tempfile.mkdtemp(prefix="gzgold")

    def touch(name):
        path = os.path.join(tmp, name)
        open(path, "w").close()
        return path

    def run(tag, model, track, varlsm, lsm_model, distorded, Np):
        fm, fs = touch(tag + "_mod.nc"), touch(tag + "_sat.nc")
        fl = fm if lsm_model is None else touch(tag + "_lsm.nc")
        store["src"][fm] = model
        store["src"][fs] = track
        if lsm_model is not None:
            store["src"][fl] = lsm_model
        with quiet():
            (it1, it2), (Nts, Ntm) = gz.GetEpochTimeOverlap(fs, fm)
            MG = gz.ModGrid(fm, it1, it2, fl, varlsm, distorded_grid=distorded)
            ST = gz.SatTrack(fs, it1, it2, Np=Np, domain_bounds=MG.domain_bounds, l_0_360=MG.l360)
            R = gz.Model2SatTrack(MG, "ssh", ST, "ssh")
        o = {"it": np.array([it1, it2]), "nts_ntm": np.array([Nts, Ntm]),
             "mg_time": MG.time, "mg_size": MG.size, "mg_lat": MG.lat, "mg_lon": MG.lon, "mg_mask": MG.mask,
             "mg_scal": np.array([MG.HResDeg, MG.HResKM, float(MG.IsLonGlobal), float(MG.l360), MG.EWPer]),
             "mg_bounds": np.array(MG.domain_bounds), "mg_xangle": MG.xangle,
             "st_time": ST.time, "st_lat": ST.lat, "st_lon": ST.lon, "st_j": np.array([ST.jt1, ST.jt2, ST.size]),
             "st_keepit": ST.keepit[0], "r_mask": R.mask, "r_dist": R.distance}
        for nm in ("ssh_mod_np", "ssh_mod", "ssh_sat", "XNPtrack"):
            pack_masked("r_" + nm, getattr(R, nm), o)
        for k, v in o.items():
            out[tag + "_" + k] = v

    # (a) global periodic, -D, mask in a separate "file"; long track file (two-pass time scan)
    lat, lon = syn.orca_like_grid(146, 182, pole_lon_deg=0.0)      # cap bent along the first meridian: column 0 keeps lon 0
    # overlap columns at the END (cols ni-2, ni-1 repeat cols 0, 1), the layout of the reference's own
    # periodic test array: the longitude heuristics then say "global, 2 overlapping points"
    lat = np.ascontiguousarray(np.concatenate([lat[:, 1:-1], lat[:, 1:3]], axis=1))
    lon = np.ascontiguousarray(np.concatenate([lon[:, 1:-1], lon[:, 1:3]], axis=1))
    mask = syn.land_mask(lat, lon)
    tm = 1.0e9 + 7000.0 * np.arange(6)
    f32 = syn.ssh_records(lat, lon, 6, dtype=np.float32)
    # (the reference needs the track window strictly inside the model records: its bracket search
    # runs off the end of MG.time otherwise, gonzag/mod2sat.py:95)
    t, y, x = syn.orbit_track(30000, t0=1000.0, **syn.SARAL)
    t = t + 1.0e9
    ssat = sat_ssh(len(t), 13)
    out.update(a_lat=lat.astype(np.float32), a_lon=lon.astype(np.float32), a_maskin=mask.astype(np.int8), a_tm=tm,
               a_f32=f32, a_t=t, a_y=y, a_x=x, a_ssat=ssat)
    run("a", ModelArrays(tm, lat.astype(np.float32).astype(np.float64), lon.astype(np.float32).astype(np.float64),
                         fields={"ssh": f32}),
        TrackArrays(t, y, x, {"ssh": ssat}), "tmask",
        ModelArrays(tm, None, None, masks={"tmask": mask}), True, len(t))
    # (b) regional across Greenwich (2-D coordinates in [-180,180] input frame), no -D
    latb, lonb = syn.regional_grid(120, 160, lat0=30.0, lon0=352.0, res_deg=0.1, shear=0.05)
    lonb_in = np.where(lonb > 180.0, lonb - 360.0, lonb)
    maskb = syn.land_mask(latb, lonb, seed=6, level=0.95)
    tmb = 2.0e9 + 3600.0 * np.arange(5)
    fb = syn.ssh_records(latb, lonb, 5, dtype=np.float64)
    tb, yb, xb = syn.orbit_track(43200, t0=0.0, drop_land_seed=None, **syn.SARAL)
    tb = tb * (3.5 * 3600.0 / 43200.0) + 2.0e9 + 300.0     # squeeze a day of passes into the record span
    xb_in = np.where(xb > 180.0, xb - 360.0, xb)
    ssb = sat_ssh(len(tb), 14)
    out.update(b_lat=latb, b_lon=lonb_in, b_maskin=maskb.astype(np.int8), b_tm=tmb, b_f64=fb, b_t=tb, b_y=yb,
               b_x=xb_in, b_ssat=ssb)
    run("b", ModelArrays(tmb, latb, lonb_in, fields={"ssh": fb}, masks={"tmask": maskb}),
        TrackArrays(tb, yb, xb_in, {"ssh": ssb}), "tmask", None, False, 100)
    # (c) regular regional grid given by 1-D coordinates, mask from _FillValue of the field
    lat1 = np.linspace(-30.0, -10.0, 81)
    lon1 = np.linspace(100.0, 130.0, 121)
    lon2, lat2 = np.meshgrid(lon1, lat1)
    maskc = syn.land_mask(lat2, lon2, seed=7, level=0.9)
    tmc = 3.0e9 + 1800.0 * np.arange(7)
    fc = syn.ssh_records(lat2, lon2, 7, dtype=np.float32)
    fcm = np.ma.masked_array(fc, mask=np.broadcast_to(maskc == 0, fc.shape))
    tc_, yc, xc = syn.orbit_track(43200, t0=0.0, drop_land_seed=None, **syn.JASON)
    tc_ = tc_ * (2.8 * 3600.0 / 43200.0) + 3.0e9 + 30.0
    ssc = sat_ssh(len(tc_), 15)
    out.update(c_lat1=lat1, c_lon1=lon1, c_maskin=maskc.astype(np.int8), c_tm=tmc, c_f32=fc, c_t=tc_, c_y=yc, c_x=xc,
               c_ssat=ssc)
    run("c", ModelArrays(tmc, lat1, lon1, fields={"ssh": fcm}),
        TrackArrays(tc_, yc, xc, {"ssh": ssc}), "_FillValue@ssh", None, False, len(tc_))
    return out


def case_window(gz, store):
    """A model time window that does NOT start at record 0: the track begins more than two model
    records into the model file, so ModGrid keeps MG.time = time[jt1:jt2+1] with jt1 >= 2 while Model2SatTrack goes on
    asking its reader for record k = index into that sliced vector (gonzag/mod2sat.py:94-112) -- i.e. records
    0 .. nrec-1 of the file.  Unmodified GetEpochTimeOverlap -> ModGrid -> SatTrack -> Model2SatTrack."""
    import tempfile
    from gonzag_b200.sources import ModelArrays, TrackArrays
    tmp = tempfile.mkdtemp(prefix="gzgoldw")
    lat, lon = syn.regional_grid(90, 110, lat0=-42.0, lon0=12.0, res_deg=0.125, shear=0.03)
    mask = syn.land_mask(lat, lon, seed=9, level=0.9)
    tm = 4.0e9 + 1800.0 * np.arange(9)
    f64 = syn.ssh_records(lat, lon, 9, dtype=np.float64)
    t, y, x = syn.orbit_track(43200, t0=0.0, drop_land_seed=None, **syn.SARAL)
    t = t * (2.6 * 1800.0 / 43200.0) + 4.0e9 + 2.0 * 1800.0 + 600.0      # inside records 2 .. 5
    ssat = sat_ssh(len(t), 16)
    out = dict(w_lat=lat, w_lon=lon, w_maskin=mask.astype(np.int8), w_tm=tm, w_f64=f64, w_t=t, w_y=y, w_x=x, w_ssat=ssat)
    fm, fs = os.path.join(tmp, "w_mod.nc"), os.path.join(tmp, "w_sat.nc")
    for f in (fm, fs):
        open(f, "w").close()
    store["src"][fm] = ModelArrays(tm, lat, lon, fields={"ssh": f64}, masks={"tmask": mask})
    store["src"][fs] = TrackArrays(t, y, x, {"ssh": ssat})
    with quiet():
        (it1, it2), (Nts, Ntm) = gz.GetEpochTimeOverlap(fs, fm)
        MG = gz.ModGrid(fm, it1, it2, fm, "tmask", distorded_grid=False)
        ST = gz.SatTrack(fs, it1, it2, Np=100, domain_bounds=MG.domain_bounds, l_0_360=MG.l360)
        R = gz.Model2SatTrack(MG, "ssh", ST, "ssh")
    assert MG.time[0] == tm[2], "the window must start two records in"
    o = {"it": np.array([it1, it2]), "nts_ntm": np.array([Nts, Ntm]),
         "mg_time": MG.time, "mg_size": MG.size, "mg_lat": MG.lat, "mg_lon": MG.lon, "mg_mask": MG.mask,
         "mg_scal": np.array([MG.HResDeg, MG.HResKM, float(MG.IsLonGlobal), float(MG.l360), MG.EWPer]),
         "mg_bounds": np.array(MG.domain_bounds), "mg_xangle": MG.xangle,
         "st_time": ST.time, "st_lat": ST.lat, "st_lon": ST.lon, "st_j": np.array([ST.jt1, ST.jt2, ST.size]),
         "st_keepit": ST.keepit[0], "r_mask": R.mask, "r_dist": R.distance}
    for nm in ("ssh_mod_np", "ssh_mod", "ssh_sat", "XNPtrack"):
        pack_masked("r_" + nm, getattr(R, nm), o)
    for k, v in o.items():
        out["w_" + k] = v
    return out


def main():
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    gz, store = load_reference(ref_root)
    only = sys.argv[2:] if len(sys.argv) > 2 else None
    for name, fn in (("units", case_units), ("global", case_global), ("regional", case_regional),
                     ("seam", case_seam), ("grid", case_grid), ("window", case_window)):
        if only and name not in only:
            continue
        data = fn(gz, store)
        path = os.path.join(HERE, "golden_%s.npz" % name)
        np.savez_compressed(path, **data)
        print("%-9s -> %s  (%.2f MB)" % (name, os.path.basename(path), os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    main()
