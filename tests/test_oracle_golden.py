"""Pin the CPU oracle (oracle/gz_oracle.py) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  Integer outputs must be identical;
float outputs are expected to be bit-identical on the machine that generated the
fixtures and may differ in the last place elsewhere (NumPy dispatches sin/cos/arcsin to
CPU-specific SIMD code), so they are held to <= 4 ulp."""
import numpy as np
import pytest

from oracle import gz_oracle as orc
from tests.conftest import ulp_diff

ULP = 4


def _grid(g):
    lat = g["lat"].astype(np.float64)
    lon = g["lon"].astype(np.float64)
    return lat, lon


def test_heading(golden_units):
    g = golden_units
    got = np.array([orc.heading(*row) for row in g["heading_in"]])
    assert ulp_diff(got, g["heading_out"]) <= ULP


def test_frame_pm180(golden_units):
    g = golden_units
    got = np.array([orc.frame_pm180(float(v)) for v in g["pm180_in"]])
    np.testing.assert_array_equal(got, g["pm180_out"])
    np.testing.assert_array_equal(orc.frame_pm180(g["pm180_in"].copy()), g["pm180_out_vec"])


def test_local_coords(golden_units):
    g = golden_units
    got = np.array([orc.local_coords(vy.copy(), vx.copy()) for vy, vx in zip(g["ab_vy"], g["ab_vx"])])
    np.testing.assert_allclose(got, g["ab_out"], rtol=1e-12, atol=1e-14)
    # degenerate quads fall back to the cell centre
    np.testing.assert_array_equal(got[300:330], 0.5)


def test_haversine(golden_units):
    g = golden_units
    got = orc.haversine_km(np.float64(g["hav_pt"][0]), np.float64(g["hav_pt"][1]), g["hav_glat"], g["hav_glon"])
    assert ulp_diff(got, g["hav_out"]) <= ULP
    got = np.array([orc.haversine_scalar_km(*r) for r in g["hs_in"]])
    assert ulp_diff(got, g["hs_out"]) <= ULP
    got = np.array([orc.earth_radius_km(r[0]) for r in g["hs_in"]])
    assert ulp_diff(got, g["re_out"]) <= ULP


def test_nearest_point_quirks(golden_units):
    g = golden_units
    Y, X = g["np_Y"], g["np_X"]
    for call, want in zip(g["np_calls"], g["np_out"]):
        yt, xt, jp, ip, rr = call
        jp = None if jp == -99 else int(jp)
        ip = None if ip == -99 else int(ip)
        got = orc.nearest_point((yt, xt), Y, X, rd_found_km=float(g["np_rd"]), j_prv=jp, i_prv=ip,
                                np_box_r=int(rr))
        assert list(got) == list(want), (call, got, want)


def test_grid_angle(golden_units):
    g = golden_units
    got = orc.grid_angle(g["ga_lat"], g["ga_lon"])
    np.testing.assert_allclose(got, g["ga_out"], rtol=0, atol=1e-11)
    assert (np.abs(got) > 45.0).mean() > 0.02      # the fixture does exercise -D


def test_single_mesh_calls(golden_units):
    g = golden_units
    Y, X = g["np_Y"], g["np_X"]
    for call, sm_want, wb_want in zip(g["sm_calls"], g["sm_out"], g["wb_out"]):
        yt, xt, jP, iP, kew, ang = call
        sm = orc.source_mesh((yt, xt), Y, X, int(jP), int(iP), k_ew_per=int(kew), grid_s_angle=ang)
        np.testing.assert_array_equal(sm, sm_want)
        wb = orc.bilinear_weights((yt, xt), Y, X, sm)
        np.testing.assert_allclose(wb, wb_want, rtol=1e-12, atol=1e-14)


def _check_mapping(g, lat, lon, kew, angle, rd, r, suffix=""):
    bt = orc.TrackMapping(g["ysat"], g["xsat"], lat, lon, src_grid_local_angle=angle,
                          k_ew_per=kew, rd_found_km=rd, np_box_r=r)
    np.testing.assert_array_equal(bt.NP, g["NP" + suffix])
    np.testing.assert_array_equal(bt.SM, g["SM" + suffix])
    np.testing.assert_allclose(bt.WB, g["WB" + suffix], rtol=1e-12, atol=1e-14)
    return bt


def test_global_case(golden_global):
    g = golden_global
    lat, lon = _grid(g)
    res = float(g["res"])
    rd = 0.75 * res * 111.11
    r = int(0.5 * 500. / (res * 111.11))
    assert r == 2
    bt = _check_mapping(g, lat, lon, 2, g["angle"], rd, r)
    _check_mapping(g, lat, lon, -1, (), rd, r, suffix="_noper")
    # quirk coverage inside the fixture: not-found points, failed quadrants, heavy-duty cells
    assert (bt.NP[:, 0] == -1).any()
    assert (np.abs(g["angle"][bt.NP[:, 0], bt.NP[:, 1]]) > 45.0).any()
    mask = g["mask"].astype(np.int64)
    f32 = g["f32"]
    f64 = f32.astype(np.float64) + lat[None] * 2.0 ** -20
    for tag, fld in (("f32", f32), ("f64", f64)):
        for faithful in (False, True):
            v_np, v_bl = orc.interp_track(g["tsat"], g["tmod"], lambda k: fld[k], mask, bt.SM, bt.WB,
                                          faithful_cost=faithful)
            im = orc.output_mask(v_bl, g["ssat"])
            np.testing.assert_array_equal(im, g["imask_" + tag])
            ok = im == 1
            np.testing.assert_allclose(v_np[ok], g["ssh_np_%s_data" % tag][ok], rtol=1e-12)
            np.testing.assert_allclose(v_bl[ok], g["ssh_bl_%s_data" % tag][ok], rtol=1e-12, atol=1e-15)
            np.testing.assert_array_equal(~ok, g["ssh_bl_%s_mask" % tag])
            # unmasked raw values too (the reference keeps them under the mask)
            np.testing.assert_allclose(v_bl, g["ssh_bl_%s_data" % tag], rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(v_np, g["ssh_np_%s_data" % tag], rtol=1e-12)
    d = orc.along_track_distance(g["ysat"], g["xsat"])
    np.testing.assert_allclose(d, g["distance"], rtol=1e-13)
    x = orc.xnp_track(bt.NP, mask)
    np.testing.assert_array_equal(np.ma.getmaskarray(x), g["xnp_mask"])
    np.testing.assert_array_equal(np.ma.getdata(x), g["xnp_data"])


def test_regional_case(golden_regional):
    g = golden_regional
    lat, lon = _grid(g)
    rd, r = float(g["rd"]), int(g["r"])
    assert r == 26          # res = 0.08335 after the shear -> int(250 / 9.261)
    bt = _check_mapping(g, lat, lon, -1, (), rd, r)
    assert (bt.NP[:, 0] == -1).sum() > 10           # points outside the domain
    mask = g["mask"].astype(np.int64)
    v_np, v_bl = orc.interp_track(g["tsat"], g["tmod"], lambda k: g["f32"][k], mask, bt.SM, bt.WB)
    np.testing.assert_allclose(v_bl, g["ssh_bl_data"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(v_np, g["ssh_np_data"], rtol=1e-12)
    np.testing.assert_array_equal(orc.output_mask(v_bl, g["ssat"]), g["imask"])
    # shuffled targets (no continuity)
    p = g["perm"]
    bts = orc.TrackMapping(g["ysat"][p], g["xsat"][p], lat, lon, k_ew_per=-1, rd_found_km=rd, np_box_r=r)
    np.testing.assert_array_equal(bts.NP, g["NP_perm"])
    np.testing.assert_array_equal(bts.SM, g["SM_perm"])
    np.testing.assert_allclose(bts.WB, g["WB_perm"], rtol=1e-12, atol=1e-14)


def test_seam_case(golden_seam, golden_global):
    g = golden_seam
    lat, lon = _grid(golden_global)
    rd, r = float(g["rd"]), int(g["r"])
    bt = _check_mapping(g, lat, lon, 2, golden_global["angle"], rd, r)
    _check_mapping(g, lat, lon, -1, golden_global["angle"], rd, r, suffix="_noper")
    # the fixture must contain box answers that differ from the plain global answer
    ni = lat.shape[1]
    assert (bt.NP[:, 1] == ni - 2).any() and (bt.NP[:, 1] == ni - 1).any()
