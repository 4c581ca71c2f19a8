"""Full-size checks on the B200 (BASELINE configs 3 and 4 shapes), where running the oracle on
every point would take hours.  Two kinds of evidence:

* the oracle itself on a contiguous prefix (the seeding chain) and on random subsets (per-point
  stages);
* size-independent properties checked on EVERY point with plain torch fp64 on the GPU:
  - every link of the nearest-point chain, R(t) == F_t(R(t-1)), re-derived by brute force over the
    reference's (2r+1)^2 search box (and over the whole grid where the box does not accept);
  - chunk invariance: mapping the track in two calls with the seed handed over == one call;
  - a constant field interpolates to the constant wherever the reference interpolates at all;
  - K0 against a vectorised torch restatement of GridAngle.
"""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import gonzag_b200 as gzb
from gonzag_b200 import synthetic as syn
from oracle import gz_oracle as orc

D2R = np.pi / 180.0


def hav_torch(plat, plon, glat, glon):
    """gonzag/utils.py:296-316 in torch fp64 (same operation order)."""
    a1 = torch.sin(0.5 * ((glat - plat) * D2R))
    a2 = torch.sin(0.5 * ((glon - plon) * D2R))
    a3 = torch.cos(glat * D2R) * torch.cos(plat * D2R)
    return 2.0 * 6360.0 * torch.arcsin(torch.sqrt(a1 * a1 + a3 * a2 * a2))


def edge(j, i, Nj, Ni, kew):
    bad = (i == 0) | ((i == Ni - 1) & (kew == -1))
    j = torch.where(bad, -1, j)
    i = torch.where(bad, -1, i)
    bad = (j == 0) | (j == Nj - 1)
    return torch.where(bad, -1, j), torch.where(bad, -1, i)


def verify_chain_links(lat, lon, y, x, NP, kew, rd, r, batch=2048):
    """Brute-force check of R(t) == F_t(R(t-1)) for every t (gonzag/bilin_mapping.py:144-191,
    566-593).  Returns the number of links that needed the whole-grid fallback."""
    dev = torch.device("cuda")
    glat = torch.as_tensor(lat, device=dev)
    glon = torch.as_tensor(lon, device=dev)
    Nj, Ni = lat.shape
    yt = torch.as_tensor(y, device=dev)
    xt = torch.as_tensor(x, device=dev)
    R = torch.as_tensor(NP, device=dev)
    nt = len(y)
    prev = torch.cat([torch.tensor([[r, r]], device=dev, dtype=torch.int64), R[:-1]], 0)
    thr2 = 1.2 * (1.2 * rd)
    w = 2 * r + 1
    oj = torch.arange(w, device=dev).view(1, w, 1)
    oi = torch.arange(w, device=dev).view(1, 1, w)
    n_global = 0
    need_global = []
    for a in range(0, nt, batch):
        b = min(a + batch, nt)
        pj, pi = prev[a:b, 0], prev[a:b, 1]
        use_box = (pj != 0) & (pi != 0)
        j0 = torch.clamp(pj - r, min=0)
        j1 = torch.clamp(pj + r + 1, max=Nj)
        i0 = torch.clamp(pi - r, min=0)
        i1 = torch.clamp(pi + r + 1, max=Ni)
        jj = j0.view(-1, 1, 1) + oj
        ii = i0.view(-1, 1, 1) + oi
        inb = (jj < j1.view(-1, 1, 1)) & (ii < i1.view(-1, 1, 1))
        jc = torch.clamp(jj, max=Nj - 1).expand(-1, w, w)
        ic = torch.clamp(ii, max=Ni - 1).expand(-1, w, w)
        flat = jc * Ni + ic
        d = hav_torch(yt[a:b].view(-1, 1, 1), xt[a:b].view(-1, 1, 1), glat.view(-1)[flat], glon.view(-1)[flat])
        d = torch.where(inb.expand(-1, w, w), d, torch.full_like(d, float("inf")))
        # first occurrence in row-major order of the box
        dmin, k = torch.min(d.view(b - a, -1), dim=1)
        ties = (d.view(b - a, -1) == dmin.view(-1, 1))
        k = torch.argmax(ties.to(torch.int8), dim=1)
        bj = j0 + k // w
        bi = i0 + k % w
        acc = use_box & (dmin < rd)
        ej, ei = edge(bj, bi, Nj, Ni, kew)
        got = R[a:b]
        ok_box = acc & (got[:, 0] == ej) & (got[:, 1] == ei)
        rest = ~acc
        assert bool((ok_box | rest).all()), "box-accepted link differs at %s" % (a + torch.nonzero(~(ok_box | rest))[:5].flatten().tolist())
        need_global.append(a + torch.nonzero(rest).flatten())
    need_global = torch.cat(need_global)
    n_global = int(need_global.numel())
    gl, go = glat.view(1, -1), glon.view(1, -1)
    for a in range(0, n_global, 16):
        idx = need_global[a:a + 16]
        d = hav_torch(yt[idx].view(-1, 1), xt[idx].view(-1, 1), gl, go)
        dmin, _ = torch.min(d, dim=1)
        k = torch.argmax((d == dmin.view(-1, 1)).to(torch.int8), dim=1)
        gj, gi = k // Ni, k % Ni
        found = dmin < thr2
        gj = torch.where(found, gj, -1)
        gi = torch.where(found, gi, -1)
        ej, ei = edge(gj, gi, Nj, Ni, kew)
        got = R[idx]
        bad = (got[:, 0] != ej) | (got[:, 1] != ei)
        assert not bool(bad.any()), "global-fallback link differs at %s" % idx[bad][:5].tolist()
    return n_global


def grid_angle_torch(lat, lon):
    """Vectorised restatement of gonzag/utils.py:226-262."""
    la = torch.as_tensor(lat, device="cuda")
    lo = torch.as_tensor(lon, device="cuda")
    t = torch.tan(np.pi / 4.0 - D2R * la / 2.0)
    c, s = torch.cos(D2R * lo), torch.sin(D2R * lo)
    px = 0.0 - 2.0 * c[1:-1] * t[1:-1]
    py = 0.0 - 2.0 * s[1:-1] * t[1:-1]
    pn = px * px + py * py
    vx = 2.0 * (c[2:] * t[2:] - c[:-2] * t[:-2])
    vy = 2.0 * (s[2:] * t[2:] - s[:-2] * t[:-2])
    nv = torch.clamp(torch.sqrt(pn * (vx * vx + vy * vy)), min=1.0e-12)
    ang = torch.atan2((px * vy - py * vx) / nv, (px * vx + py * vy) / nv) * 180.0 / np.pi
    same = torch.abs(torch.remainder(lo[2:], 360.0) - torch.remainder(lo[:-2], 360.0)) < 1.0e-8
    ang = torch.where(same, torch.zeros_like(ang), ang)
    out = torch.zeros_like(la)
    out[1:-1] = ang
    return out.cpu().numpy()


@pytest.fixture(scope="module")
def c3():
    lat, lon = syn.orca_like_grid(3059, 4322)
    mask = syn.land_mask(lat, lon)
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    t, y, x = syn.orbit_track(2 * 86400, t0=100.0, **syn.JASON)
    ctx = gzb.GridContext(lat, lon, mask=mask, k_ew_per=2)
    ang = ctx.grid_angle()
    yield dict(lat=lat, lon=lon, mask=mask, rd=rd, r=r, t=t, y=y, x=x, ctx=ctx, ang=ang)
    ctx.close()


def test_c3_grid_angle(c3):
    want = grid_angle_torch(c3["lat"], c3["lon"])
    np.testing.assert_allclose(c3["ang"], want, rtol=0, atol=1e-8)
    assert 0.02 < (np.abs(c3["ang"]) > 45.0).mean() < 0.3
    band = orc.grid_angle(c3["lat"][:, 1000:1003], c3["lon"][:, 1000:1003])
    np.testing.assert_allclose(c3["ang"][:, 1000:1003], band, rtol=0, atol=1e-9)


def test_c3_narrowed_upload_is_exact(c3, monkeypatch):
    """A big grid whose float64 coordinates are float32 values (written in single precision by the model, widened by the
    reader) goes up as float32 and is widened on the device; a grid with a value that does not survive the narrowing goes up
    in full precision.  Both must give the context the very same float64 arrays as the plain upload: compared through K0
    (every cell's angle depends on its coordinates and its neighbours') and K1 on a day of track."""
    import os
    g = c3
    n = 20000
    want_ang, want_np = g["ang"], g["ctx"].nearest(g["y"][:n], g["x"][:n], g["rd"], g["r"])
    monkeypatch.setenv("GZB_NO_NARROW_UPLOAD", "1")
    with gzb.GridContext(g["lat"], g["lon"], k_ew_per=2) as c_full:
        assert c_full.stats()["h2d_bytes"] >= 2 * g["lat"].nbytes
        assert np.array_equal(c_full.grid_angle().view(np.int64), want_ang.view(np.int64))
        np.testing.assert_array_equal(c_full.nearest(g["y"][:n], g["x"][:n], g["rd"], g["r"]), want_np)
    monkeypatch.delenv("GZB_NO_NARROW_UPLOAD")
    with gzb.GridContext(g["lat"], g["lon"], k_ew_per=2) as c_nar:
        assert c_nar.stats()["h2d_bytes"] < 1.5 * g["lat"].nbytes          # the fixture's grid is float32-exact: half the bytes
        assert np.array_equal(c_nar.grid_angle().view(np.int64), want_ang.view(np.int64))
        for q, src in ((0, g["lat"]), (1, g["lon"])):
            back = c_nar.download(types.SimpleNamespace(ptr=c_nar.grid_arrays()[q]), src.shape, np.float64)
            assert np.array_equal(back.view(np.int64), src.view(np.int64))
    # one value that float32 cannot hold, somewhere in the middle: full precision, and the value arrives intact
    lat2 = g["lat"].copy()
    j, i = 1500, 2100
    lat2[j, i] = lat2[j, i] + 1.0e-11
    assert float(np.float32(lat2[j, i])) != lat2[j, i]
    with gzb.GridContext(lat2, g["lon"], k_ew_per=2) as c2:
        assert c2.stats()["h2d_bytes"] >= 2 * g["lat"].nbytes
        a2 = c2.grid_angle()
        back = c2.download(types.SimpleNamespace(ptr=c2.grid_arrays()[0]), lat2.shape, np.float64)
        assert np.array_equal(back.view(np.int64), lat2.view(np.int64))
    monkeypatch.setenv("GZB_NO_NARROW_UPLOAD", "1")
    with gzb.GridContext(lat2, g["lon"], k_ew_per=2) as c3_:
        a3 = c3_.grid_angle()
    assert np.array_equal(a2.view(np.int64), a3.view(np.int64))


def test_c3_chain_links_and_oracle_prefix(c3):
    g = c3
    NP, SM, WB = g["ctx"].mapping(g["y"], g["x"], g["rd"], g["r"])
    n_glob = verify_chain_links(g["lat"], g["lon"], g["y"], g["x"], NP, 2, g["rd"], g["r"])
    assert n_glob < 0.2 * len(g["y"])
    # the sequential oracle on a prefix
    n = 2500
    want = orc.nearest_chain(g["y"][:n], g["x"][:n], g["lat"], g["lon"], 2, g["rd"], g["r"])
    np.testing.assert_array_equal(NP[:n], want)
    # chunk invariance with the seed handed over
    cut = len(g["y"]) // 2 + 17
    a = g["ctx"].nearest(g["y"][:cut], g["x"][:cut], g["rd"], g["r"])
    b = g["ctx"].nearest(g["y"][cut:], g["x"][cut:], g["rd"], g["r"], seed=tuple(int(v) for v in a[-1]))
    np.testing.assert_array_equal(np.concatenate([a, b]), NP)
    # the scheduling of the overlapped path (bulk kernel released at once / behind the warp search, persistent bulk blocks)
    # changes nothing in the results; the default (chain_first = 6) only applies to tracks this long
    for opts in ({"chain_first": 0}, {"chain_first": 12, "bulk_resident": 3}, {"chain_first": 6, "bulk_resident": 0}):
        for k_, v_ in opts.items():
            g["ctx"].set_option(k_, v_)
        NP2, SM2, WB2 = g["ctx"].mapping(g["y"], g["x"], g["rd"], g["r"])
        assert np.array_equal(NP2, NP) and np.array_equal(SM2, SM) and np.array_equal(WB2.view(np.int64), WB.view(np.int64)), opts
    # per-point stages against the oracle on a random subset (includes cells with |angle| > 45)
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(len(g["y"]), 3000, replace=False))
    hot = np.where((NP[:, 0] >= 0) & (np.abs(g["ang"][NP[:, 0], NP[:, 1]]) > 45.0))[0]
    if hot.size:
        idx = np.unique(np.concatenate([idx, hot[:1500]]))
    sm = orc.source_meshes(g["y"], g["x"], g["lat"], g["lon"], NP, 2, angle=g["ang"], idx=idx)
    bad = np.where((SM[idx] != sm).reshape(len(idx), -1).any(1))[0]
    assert bad.size == 0, (idx[bad][:5], SM[idx][bad][:2], sm[bad][:2])
    wb = orc.weights(g["y"], g["x"], g["lat"], g["lon"], sm, idx=idx)
    np.testing.assert_allclose(WB[idx], wb, rtol=1e-6, atol=1e-9)
    # interpolation: oracle on the subset + constant-field property on every point
    nrec = 3
    tmod = np.array([0.0, g["t"][-1] / 2, g["t"][-1] + 10.0])
    fld = syn.ssh_records(g["lat"], g["lon"], nrec)
    v_np, v_bl = g["ctx"].interp(g["t"], SM, WB, tmod, fld)
    o_np, o_bl = orc.interp_track(g["t"][idx], tmod, lambda k: fld[k], g["mask"], SM[idx], WB[idx])
    np.testing.assert_allclose(v_bl[idx], o_bl, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(v_np[idx], o_np, rtol=1e-6, atol=1e-9)
    const = np.full((2,) + g["lat"].shape, 0.375, dtype=np.float32)
    c_np, c_bl = g["ctx"].interp(g["t"], SM, WB, np.array([0.0, g["t"][-1] + 10.0]), const)
    done = c_bl != -9999.0
    assert done.mean() > 0.5
    np.testing.assert_allclose(c_bl[done], 0.375, rtol=1e-9)
    assert set(np.unique(c_np)) <= {-9999.0, 0.375}
    np.testing.assert_array_equal(c_np != -9999.0, v_np != -9999.0)


def test_c4_shape_chain_links():
    """ORCA36-class grid (BASELINE config 4 shape), half a day of track: K0-K4 against the oracle on a subset, every chain
    link of K1 verified by brute force."""
    lat, lon = syn.orca_like_grid(9000, 13000)
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    assert r >= 80
    t, y, x = syn.orbit_track(43200, t0=100.0, **syn.JASON)
    mask = syn.land_mask(lat, lon)
    with gzb.GridContext(lat, lon, mask=mask, k_ew_per=2) as ctx:
        NP = ctx.nearest(y, x, rd, r)
        assert ctx.stats()["pyramid_levels"] >= 5
        # K0 + K2 + K3 + K4 at this size: the fused mapping call equals the staged nearest points, then the per-point
        # stages against the oracle on a random subset plus the points whose cell is rotated by more than 45 degrees
        ang = ctx.grid_angle()
        NP2, SM, WB = ctx.mapping(y, x, rd, r)
        np.testing.assert_array_equal(NP2, NP)
        rng = np.random.default_rng(4)
        idx = np.sort(rng.choice(len(y), 2000, replace=False))
        hot = np.where((NP[:, 0] >= 0) & (np.abs(ang[NP[:, 0], NP[:, 1]]) > 45.0))[0]
        if hot.size:
            idx = np.unique(np.concatenate([idx, hot[:1000]]))
        sm = orc.source_meshes(y, x, lat, lon, NP, 2, angle=ang, idx=idx)
        bad = np.where((SM[idx] != sm).reshape(len(idx), -1).any(1))[0]
        assert bad.size == 0, (idx[bad][:5], SM[idx][bad][:2], sm[bad][:2])
        np.testing.assert_allclose(WB[idx], orc.weights(y, x, lat, lon, sm, idx=idx), rtol=1e-6, atol=1e-9)
        tmod = np.array([0.0, t[-1] + 10.0])
        fld = syn.ssh_records(lat, lon, 2)
        v_np, v_bl = ctx.interp(t, SM, WB, tmod, fld)
        o_np, o_bl = orc.interp_track(t[idx], tmod, lambda k: fld[k], mask, SM[idx], WB[idx])
        np.testing.assert_allclose(v_bl[idx], o_bl, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(v_np[idx], o_np, rtol=1e-6, atol=1e-9)
        assert (v_bl != -9999.0).mean() > 0.5
    assert (NP[:, 0] >= 0).mean() > 0.9
    verify_chain_links(lat, lon, y, x, NP, 2, rd, r, batch=128)


def test_c3_polar_track_all_paths_agree(c3):
    """A SARAL-like (81.5 deg) day on the global ORCA12-shaped grid: it spends a third of its time over
    the bent northern cap, where certificates fail most often and the chain logic works hardest.  Every
    link re-derived by brute force; per-thread / warp search and twin / plain speculation agree."""
    g = c3
    t, y, x = syn.orbit_track(86400, t0=33.0, **syn.SARAL)
    ctx = g["ctx"]
    NP = ctx.nearest(y, x, g["rd"], g["r"])
    st = ctx.stats()
    assert st["last_unresolved"] < 0.25 * len(y)      # the bent cap defeats the per-cell certificate often; the warp search takes those
    verify_chain_links(g["lat"], g["lon"], y, x, NP, 2, g["rd"], g["r"])
    for opt in ("twin_chain", "nearest_path"):
        ctx.set_option(opt, 0)
        other = ctx.nearest(y, x, g["rd"], g["r"])
        ctx.set_option(opt, 1)
        np.testing.assert_array_equal(other, NP)
    n = 1500
    want = orc.nearest_chain(y[:n], x[:n], g["lat"], g["lon"], 2, g["rd"], g["r"])
    np.testing.assert_array_equal(NP[:n], want)
