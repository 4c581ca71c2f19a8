"""Edge cases of the hot path on the GPU against the CPU oracle: empty and one-point tracks, the
smallest grids, degenerate search parameters (no box, a box larger than the grid, a found-radius that
accepts nothing), non-finite and out-of-domain targets, targets exactly on grid nodes (systematic
ties with the duplicated E-W columns), longitudes given outside [0,360), repeated points, every
East-West overlap convention, and tracks without any continuity."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import gonzag_b200 as gzb
from gonzag_b200 import synthetic as syn
from oracle import gz_oracle as orc

RTOL, ATOL = 1e-6, 1e-9


def _against_oracle(lat, lon, y, x, kew, rd, r, angle=None, what=""):
    with gzb.GridContext(lat, lon, angle=angle, k_ew_per=kew) as ctx:
        NP, SM, WB = ctx.mapping(y, x, rd, r)
        NP1 = ctx.nearest(y, x, rd, r)
    np.testing.assert_array_equal(NP, NP1)
    want = orc.nearest_chain(y, x, lat, lon, kew, rd, r)
    bad = np.where((NP != want).any(axis=1))[0]
    assert bad.size == 0, "%s NP: %d rows differ, first %s got %s want %s" % (what, bad.size, bad[:5], NP[bad[:3]], want[bad[:3]])
    sm = orc.source_meshes(y, x, lat, lon, want, kew, angle=angle)
    bad = np.where((SM != sm).reshape(len(y), -1).any(axis=1))[0]
    assert bad.size == 0, "%s SM: %d rows differ, first %s" % (what, bad.size, bad[:5])
    np.testing.assert_allclose(WB, orc.weights(y, x, lat, lon, sm), rtol=RTOL, atol=ATOL)
    return NP, SM, WB


def test_empty_and_single_point():
    lat, lon = syn.orca_like_grid(60, 82)
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    r = max(r, 3)
    bt = gzb.BilinTrack(np.zeros(0), np.zeros(0), lat, lon, k_ew_per=2, rd_found_km=rd, np_box_r=r)
    assert bt.NP.shape == (0, 2) and bt.SM.shape == (0, 4, 2) and bt.WB.shape == (0, 4)
    bt.close()
    _against_oracle(lat, lon, np.array([12.3]), np.array([200.1]), 2, rd, r, what="single")
    mask = syn.land_mask(lat, lon)
    fld = syn.ssh_records(lat, lon, 2)
    MG = types.SimpleNamespace(shape=lat.shape, HResDeg=res, lat=lat, lon=lon, xangle=np.zeros(lat.shape), EWPer=2,
                               time=np.array([0.0, 10.0]), file="mem", mask=mask, field=fld)
    ST = types.SimpleNamespace(size=0, lat=np.zeros(0), lon=np.zeros(0), time=np.zeros(0), file="mem", jt1=0, jt2=0,
                               keepit=[], ssh=np.zeros(0))
    R = gzb.Model2SatTrack(MG, "ssh", ST, "ssh")
    assert R.ssh_mod.shape == (0,) and R.distance.shape == (0,) and R.mask.shape == (0,)
    assert R.XNPtrack.shape == lat.shape


@pytest.mark.parametrize("shape", [(3, 3), (3, 40), (40, 3), (5, 9)])
def test_smallest_grids(shape):
    nj, ni = shape
    lat, lon = syn.regional_grid(nj, ni, lat0=-2.0, lon0=10.0, res_deg=1.0, shear=0.1)
    rng = np.random.default_rng(nj * 100 + ni)
    y = rng.uniform(lat.min() - 1.0, lat.max() + 1.0, 300)
    x = rng.uniform(lon.min() - 1.0, lon.max() + 1.0, 300)
    # (np_box_r = 0 is left out here: with a not-found predecessor the reference slices an empty box and
    # raises, gonzag/bilin_mapping.py:159-166; see test_degenerate_search_parameters for r = 0)
    for r in (1, 2, 50):
        _against_oracle(lat, lon, y, x, -1, 0.75 * 111.11, r, what="grid %s r=%d" % (shape, r))


def test_degenerate_search_parameters():
    lat, lon = syn.orca_like_grid(72, 92)
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    r = max(r, 3)
    t, y, x = syn.orbit_track(6000, t0=3.0, **syn.SARAL)
    y, x = y[::3], x[::3]
    # np_box_r = 0: the "box" is the previous cell alone, accepted while the target stays within r0 of it.
    # Only defined in the reference while every point is found (else it slices an empty box and raises).
    latr, lonr = syn.regional_grid(60, 80, lat0=-20.0, lon0=100.0, res_deg=0.5, shear=0.2)
    s_ = np.linspace(0, 1, 700)
    yi = -15.0 + 20.0 * s_ + 0.8 * np.sin(40 * s_)
    xi = 105.0 + 28.0 * s_
    NP0, _, _ = _against_oracle(latr, lonr, yi, xi, -1, 0.75 * 0.5 * 111.11, 0, what="r=0")
    assert (NP0 >= 0).all()
    _against_oracle(lat, lon, y, x, 2, rd, 500, what="box larger than the grid")
    NP, SM, WB = _against_oracle(lat, lon, y, x, 2, 1.0e-3, r, what="found radius of one metre")
    assert (NP == -1).all() and (SM == 0).all() and np.allclose(WB, 0.25)
    _against_oracle(lat, lon, y, x, 2, 1.0e5, r, what="found radius larger than the Earth")


def test_nonfinite_out_of_domain_and_on_node_targets():
    lat, lon = syn.orca_like_grid(72, 92)
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    r = max(r, 3)
    rng = np.random.default_rng(5)
    # targets exactly on grid nodes, duplicated E-W columns included (exact distance ties), then
    # nodes nudged by one ulp-ish amount
    jj = rng.integers(1, 71, 400)
    ii = np.concatenate([rng.integers(0, 92, 300), np.repeat([0, 1, 90, 91], 25)])
    y = lat[jj, ii].copy()
    x = lon[jj, ii].copy()
    y[200:300] += 1e-9
    _against_oracle(lat, lon, y, x, 2, rd, r, what="targets on nodes")
    # NaN / inf targets in the middle of a track; longitudes outside [0,360); the poles
    t, y, x = syn.orbit_track(3000, t0=3.0, **syn.SARAL)
    y, x = y.copy(), x.copy()
    y[100] = np.nan
    x[200] = np.nan
    x[300] = np.inf          # (an infinite LATITUDE makes the reference raise: math.cos(inf), gonzag/utils.py:312)
    x[400:500] -= 360.0
    x[500:600] += 720.0
    y[700], y[701] = 90.0, -90.0
    NP, _, _ = _against_oracle(lat, lon, y, x, 2, rd, r, what="non-finite / unwrapped longitudes")
    assert (NP[[100, 200, 300]] == -1).all()
    # a regional grid and a track that never enters it
    latr, lonr = syn.regional_grid(50, 60, lat0=10.0, lon0=20.0, res_deg=0.25)
    yo = rng.uniform(-60.0, -40.0, 500)
    xo = rng.uniform(200.0, 300.0, 500)
    NP, _, _ = _against_oracle(latr, lonr, yo, xo, -1, 0.75 * 0.25 * 111.11, 9, what="track outside the domain")
    assert (NP == -1).all()


def test_repeated_points_and_no_continuity():
    lat, lon = syn.orca_like_grid(72, 92)
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    r = max(r, 3)
    rng = np.random.default_rng(9)
    y = np.repeat(rng.uniform(-70, 80, 40), 25)
    x = np.repeat(rng.uniform(0, 360, 40), 25)
    _against_oracle(lat, lon, y, x, 2, rd, r, what="repeated points")
    y = np.degrees(np.arcsin(rng.uniform(-1, 1, 1500)))
    x = rng.uniform(0, 360, 1500)
    _against_oracle(lat, lon, y, x, 2, rd, r, what="random points, every one a break")


@pytest.mark.parametrize("kew", [-1, 0, 1, 2])
def test_every_overlap_convention(kew):
    """k_ew_per = -1 (none), 0, 1, 2 overlapping columns: grids built with exactly that overlap,
    tracks hugging the seam."""
    nj, nint = 48, 60
    dx = 360.0 / nint
    ni = nint + max(kew, 0)
    lon1 = np.mod(dx * np.arange(ni), 360.0)         # columns nint.. repeat columns 0..
    lat1 = np.linspace(-60.0, 60.0, nj)
    lon, lat = np.meshgrid(lon1, lat1)
    lat = np.ascontiguousarray(lat + 0.3 * np.sin(np.radians(lon)))      # a little curvature
    lon = np.ascontiguousarray(lon)
    rd, r = 0.75 * dx * 111.11, 4
    rng = np.random.default_rng(40 + kew)
    s = np.linspace(0, 1, 600)
    y = -55.0 + 110.0 * s + rng.normal(0, 0.05, 600)
    x = np.mod(352.0 + 16.0 * s + rng.normal(0, 0.05, 600), 360.0)
    _against_oracle(lat, lon, y, x, kew, rd, r, what="k_ew_per=%d" % kew)
    _against_oracle(lat, lon, y[::-1].copy(), x[::-1].copy(), kew, rd, r, what="k_ew_per=%d, westward" % kew)


def test_arbitrary_int64_indices_in_caller_arrays():
    """K2, K3 and K4 take NP / SM arrays from the caller: entries outside the grid -- also far outside 32 bits -- must give
    what the 64-bit rules give (no mesh; Python-wrapped then clamped corner), in the 32-bit index variant the kernels use
    on every grid of fewer than 2^31 cells exactly as in the 64-bit one (option "ix64")."""
    lat, lon = syn.orca_like_grid(72, 92)
    mask = syn.land_mask(lat, lon)
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    t, y, x = syn.orbit_track(4000, t0=3.0, **syn.JASON)
    tm = np.array([0.0, 3000.0, 6000.0])
    fld = syn.ssh_records(lat, lon, 3).astype(np.float32)
    rng = np.random.default_rng(77)
    odd = np.array([-1, -2, 0, 71, 72, 91, 92, 2 ** 31 - 1, 2 ** 31, -2 ** 31, -2 ** 31 - 1, 2 ** 40, -2 ** 40, 2 ** 62, -2 ** 62,
                    2 ** 32 + 5, -2 ** 32 + 5], dtype=np.int64)
    out = {}
    for ix64 in (0, 1):
        with gzb.GridContext(lat, lon, mask=mask, k_ew_per=2) as ctx:
            ctx.set_option("ix64", ix64)
            NP, SM, WB = ctx.mapping(y, x, rd, r)
            NP2, SM2 = NP.copy(), SM.copy()
            k = rng.integers(0, len(y), 600)
            NP2[k[:300], rng.integers(0, 2, 300)] = rng.choice(odd, 300)
            NP2[k[300:400]] = rng.choice(odd, (100, 2))
            SM2[k, rng.integers(0, 4, 600), rng.integers(0, 2, 600)] = rng.choice(odd, 600)
            sm_from_np2 = ctx.source_mesh(y, x, NP2)
            wb_from_sm2 = ctx.weights(y, x, SM2)
            v = ctx.interp(t, SM2, WB, tm, fld)
            out[ix64] = (NP, SM, WB, sm_from_np2, wb_from_sm2, v[0], v[1])
        rng = np.random.default_rng(77)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a.view(np.int64), b.view(np.int64))
    # an out-of-grid nearest point gives "no mesh" (-1 everywhere), never a wrapped cell
    sm = out[0][3]
    assert ((sm == -1).all(axis=(1, 2)) | (sm == 0).all(axis=(1, 2)) | (sm >= 0).all(axis=(1, 2))).all()


def test_heavy_duty_quadrant_everywhere():
    """|angle| > 45 everywhere forces the polygon search for every point (option -D on a grid
    declared fully rotated), including targets due north of their nearest point."""
    lat, lon = syn.orca_like_grid(72, 92)
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    r = max(r, 3)
    t, y, x = syn.orbit_track(5000, t0=3.0, **syn.JASON)
    ang = np.full(lat.shape, 60.0)
    _against_oracle(lat, lon, y, x, 2, rd, r, angle=ang, what="heavy-duty")
    jj = np.arange(5, 65)
    yn = lat[jj, 40] + 0.3 * (lat[jj + 1, 40] - lat[jj, 40])
    xn = lon[jj, 40].copy()
    _against_oracle(lat, lon, yn, xn, 2, rd, r, what="due north (light test finds no quadrant)")


def test_device_pool_and_parallel_copies():
    """Device buffers are recycled through the process-wide pool; big pageable arrays travel
    through the parallel staged copy: same bytes either way."""
    import ctypes as C
    from gonzag_b200 import _lib as L
    lib = L.lib()
    lat, lon = syn.orca_like_grid(700, 5000)          # 28 MB per array: above the parallel-copy threshold
    with gzb.GridContext(lat, lon, k_ew_per=2) as ctx:
        a1 = ctx.grid_angle()                          # big D2H
    cb, nb, live = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.check(lib.gzb_pool_stats(C.byref(cb), C.byref(nb), C.byref(live)))
    assert cb.value > 3 * lat.nbytes and nb.value >= 8
    L.check(lib.gzb_pool_config(0, 0))                 # pool off: plain cudaMalloc / cudaFree
    L.check(lib.gzb_pool_stats(C.byref(cb), C.byref(nb), C.byref(live)))
    assert cb.value == 0 and nb.value == 0
    with gzb.GridContext(lat, lon, k_ew_per=2) as ctx:
        a2 = ctx.grid_angle()
        buf = ctx.dev_alloc(lat.nbytes)
        ctx._ck(lib.gzb_h2d(ctx._h, C.c_void_p(buf.ptr), L.ptr(lat), lat.nbytes))       # big H2D ...
        back = np.empty_like(lat)
        ctx._ck(lib.gzb_d2h(ctx._h, L.ptr(back), C.c_void_p(buf.ptr), lat.nbytes))      # ... and back
        ctx.sync()
        buf.free()
    L.check(lib.gzb_pool_config(1, 16 << 30))
    np.testing.assert_array_equal(a1, a2)
    np.testing.assert_array_equal(back, lat)
    np.testing.assert_allclose(a1[:, 100:103], orc.grid_angle(lat[:, 100:103], lon[:, 100:103]), rtol=0, atol=1e-9)
