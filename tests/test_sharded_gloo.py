"""World-size-2 / 3 tests on CPU (gloo) of the sharded public call `gonzag_b200.ShardedModel2SatTrack`: grid
replication from one rank, time segments, seed hand-off (incl. a boundary INSIDE a chain deviation, which forces the
re-seed round), distance offsets, gather to root / all-gather, record windows per rank, XNPtrack over the whole track.
The C ABI is emulated by the oracle (tests/emulib.py, test infrastructure); expected values are the golden outputs of
the unmodified reference (tests/golden/golden_global.npz) and the oracle on a seam-hugging track."""
import os
import socket
import types

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _install_emulator():
    from gonzag_b200 import _lib as L
    from tests import emulib
    lib = emulib.EmuLib()
    L.lib = lambda: lib


def _global_case():
    g = load_golden("global")
    lat, lon = g["lat"].astype(np.float64), g["lon"].astype(np.float64)
    return g, lat, lon


def _worker_global(rank, world, port, gather, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _install_emulator()
        import gonzag_b200 as gzb
        g, lat, lon = _global_case()
        nt = len(g["tsat"])
        # only rank 0 holds the grid arrays; every rank holds the record WINDOW of its own time segment
        from gonzag_b200.multi import segment_bounds
        b = segment_bounds(nt, world)
        tm = g["tmod"]
        k_lo = int(np.searchsorted(tm, g["tsat"][b[rank]], side="right") - 1)
        k_hi = int(np.searchsorted(tm, g["tsat"][b[rank + 1] - 1], side="right"))
        window = gzb.RecordWindow(k_lo, np.ascontiguousarray(g["f32"][k_lo:k_hi + 1]))
        have_grid = rank == 0
        MG = types.SimpleNamespace(shape=lat.shape, HResDeg=float(g["res"]), lat=lat if have_grid else None,
                                   lon=lon if have_grid else None, xangle=g["angle"] if have_grid else None, EWPer=2, time=tm,
                                   file="mem", mask=g["mask"].astype(np.int64) if have_grid else None, field=window)
        ST = types.SimpleNamespace(size=nt, lat=g["ysat"], lon=g["xsat"], time=g["tsat"], file="mem", jt1=0, jt2=0,
                                   keepit=[], ssh=g["ssat"])
        R = gzb.ShardedModel2SatTrack(MG, "ssh", ST, "ssh", gather=gather, keep_mapping=True)
        out = {"rank": rank, "complete": R.complete, "segment": R.segment, "reseeds": R.reseeds}
        if R.complete:
            out.update(NP=R.BT.NP.copy(), SM=R.BT.SM.copy(), WB=R.BT.WB.copy(), mask=R.mask.copy(),
                       bl=np.ma.getdata(R.ssh_mod).copy(), blm=np.ma.getmaskarray(R.ssh_mod).copy(),
                       np_=np.ma.getdata(R.ssh_mod_np).copy(), dist=R.distance.copy(),
                       xnp=np.ma.getdata(R.XNPtrack).copy(), xnpm=np.ma.getmaskarray(R.XNPtrack).copy())
        q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,gather", [(2, "root"), (3, "all")])
def test_sharded_call_matches_reference(world, gather):
    g, lat, lon = _global_case()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_global, args=(rk, world, port, gather, q)) for rk in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs.sort(key=lambda o: o["rank"])
    assert [o["complete"] for o in outs] == ([True] + [False] * (world - 1) if gather == "root" else [True] * world)
    assert outs[0]["segment"][0] == 0 and outs[-1]["segment"][1] == len(g["tsat"])
    for o in outs:
        if not o["complete"]:
            continue
        np.testing.assert_array_equal(o["NP"], g["NP"])
        np.testing.assert_array_equal(o["SM"], g["SM"])
        np.testing.assert_allclose(o["WB"], g["WB"], rtol=1e-6, atol=1e-9)
        np.testing.assert_array_equal(o["mask"], g["imask_f32"])
        np.testing.assert_array_equal(o["blm"], g["ssh_bl_f32_mask"])
        np.testing.assert_allclose(o["bl"], g["ssh_bl_f32_data"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(o["np_"], g["ssh_np_f32_data"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(o["dist"], g["distance"], rtol=1e-9)
        np.testing.assert_array_equal(o["xnpm"], g["xnp_mask"])
        np.testing.assert_array_equal(o["xnp"], g["xnp_data"])


def _seam_case():
    from gonzag_b200 import synthetic as syn
    lat, lon = syn.orca_like_grid(146, 182)
    res = syn.grid_resolution_deg(lon)
    rng = np.random.default_rng(21)
    n = 240
    s = np.linspace(0.0, 1.0, n)
    # a seam-hugging leg: chain deviations are long, so a segment boundary falls inside one
    y = -60.0 + 120.0 * s + rng.normal(0, 0.02, n)
    x = np.mod(356.5 + 6.0 * s + rng.normal(0, 0.02, n), 360.0)
    t = 10.0 + np.arange(n, dtype=np.float64)
    mask = syn.land_mask(lat, lon)
    fld = syn.ssh_records(lat, lon, 2)
    return lat, lon, res, mask, fld, np.array([0.0, 1000.0]), t, y, x


def _worker_seam(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _install_emulator()
        import gonzag_b200 as gzb
        lat, lon, res, mask, fld, tm, t, y, x = _seam_case()
        MG = types.SimpleNamespace(shape=lat.shape, HResDeg=res, lat=lat, lon=lon, xangle=np.zeros(lat.shape), EWPer=2, time=tm,
                                   file="mem", mask=mask, field=fld)
        ST = types.SimpleNamespace(size=len(t), lat=y, lon=x, time=t, file="mem", jt1=0, jt2=0, keepit=[], ssh=np.zeros(len(t)))
        R = gzb.ShardedModel2SatTrack(MG, "ssh", ST, "ssh", gather="all", keep_mapping=True)
        q.put({"rank": rank, "NP": R.BT.NP.copy(), "bl": np.ma.getdata(R.ssh_mod).copy(), "dist": R.distance.copy(),
               "reseeds": R.reseeds})
    finally:
        dist.destroy_process_group()


def test_boundary_inside_a_chain_deviation_is_reseeded():
    from oracle import gz_oracle as orc
    lat, lon, res, mask, fld, tm, t, y, x = _seam_case()
    want = orc.track_products(lat, lon, mask, tm, lambda k: fld[k], t, y, x, np.zeros(len(t)), res,
                              angle=np.zeros(lat.shape), k_ew_per=2)
    spec = want["NP"]
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_seam, args=(rk, world, port, q)) for rk in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for o in outs:
        np.testing.assert_array_equal(o["NP"], spec)
        np.testing.assert_allclose(o["bl"], want["ssh_bl_raw"], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(o["dist"], want["distance"], rtol=1e-9)
    assert sum(o["reseeds"] for o in outs) >= 1, "the case is meant to put a boundary inside a chain deviation"
