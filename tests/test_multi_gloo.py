"""World-size-2 test of the multi-GPU driver on CPU (gloo): segmentation, the halo / seed
hand-off between segments and the final gather.  The compute engine is the CPU oracle plugged
into the driver's engine interface -- test infrastructure only; the product engine is CUDA."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gonzag_b200 import synthetic as syn
from gonzag_b200.multi import ShardedMapper, segment_bounds
from oracle import gz_oracle as orc


class OracleEngine:
    """Same interface as gonzag_b200.multi.CudaEngine, backed by the oracle (tests only)."""

    def __init__(self, lat, lon, kew):
        self.lat, self.lon, self.kew = lat, lon, kew
        self.torch = torch

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def nearest(self, y, x, rd, r, seed, out):
        res = orc.nearest_chain(y.numpy(), x.numpy(), self.lat, self.lon, self.kew, rd, r, seed=tuple(seed))
        out.copy_(torch.from_numpy(res))

    # ---- the rest of the engine interface (whole path), used by the mission-batch test
    mask = None

    def tensor(self, arr, dtype=None):
        t = torch.as_tensor(np.ascontiguousarray(arr))
        return t if dtype is None else t.to(dtype)

    def mesh_weights(self, y, x, NP, SM, WB):
        sm = orc.source_meshes(y.numpy(), x.numpy(), self.lat, self.lon, NP.numpy(), self.kew)
        SM.copy_(torch.from_numpy(sm))
        WB.copy_(torch.from_numpy(orc.weights(y.numpy(), x.numpy(), self.lat, self.lon, sm)))

    def interp(self, t, SM, WB, tmod, slab, k0, v_np, v_bl):
        recs = slab.numpy()
        a, b = orc.interp_track(t.numpy(), tmod.numpy(), lambda k: recs[k - k0], self.mask, SM.numpy(), WB.numpy())
        v_np.copy_(torch.from_numpy(a))
        v_bl.copy_(torch.from_numpy(b))

    def map_interp(self, y, x, t, rd, r, seed, tmod, slab, k0, NP, SM, WB, v_np, v_bl):
        self.nearest(y, x, rd, r, seed, NP)
        self.mesh_weights(y, x, NP, SM, WB)
        self.interp(t, SM, WB, tmod, slab, k0, v_np, v_bl)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    lat, lon = syn.orca_like_grid(146, 182)
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    rng = np.random.default_rng(21)
    n = 240
    s = np.linspace(0.0, 1.0, n)
    # a seam-hugging leg: chain deviations are long, so a segment boundary falls inside one
    y = -60.0 + 120.0 * s + rng.normal(0, 0.02, n)
    x = np.mod(356.5 + 6.0 * s + rng.normal(0, 0.02, n), 360.0)
    return lat, lon, rd, r, y, x


def _worker(rank, world, port, cut, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lat, lon, rd, r, y, x = _case()
        bounds = [0, cut, len(y)] if cut is not None else segment_bounds(len(y), world)
        s0, s1 = bounds[rank], bounds[rank + 1]
        halo = rank > 0
        h0 = s0 - 1 if halo else s0
        m = ShardedMapper(OracleEngine(lat, lon, 2), dist, torch).make_buffers(bounds)
        yh, xh = torch.from_numpy(y[h0:s1].copy()), torch.from_numpy(x[h0:s1].copy())
        off = 1 if halo else 0
        m.nearest_speculative(yh, xh, rd, r)
        m.v_np.copy_(torch.arange(s0, s1, dtype=torch.float64))        # stand-ins for the K4 products
        m.v_bl.copy_(-torch.arange(s0, s1, dtype=torch.float64))
        g = m.gather()
        m.check_boundaries_async(g)                # device-side check, no collective ...
        assert m.boundaries_verdict() == m.boundaries_ok(g)     # ... same verdict as the all-reduce version
        if not m.boundaries_ok(g):
            m.reseed(g, yh[off:], xh[off:], rd, r)
            g = m.gather()
        v_np, v_bl, full = m.unpack(g)
        assert torch.equal(v_np, torch.arange(len(y), dtype=torch.float64))
        assert torch.equal(v_bl, -torch.arange(len(y), dtype=torch.float64))
        q.put((rank, full.numpy(), m.reseeds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cut", [None, "inside_chain"])
def test_two_ranks_match_single_chain(cut):
    lat, lon, rd, r, y, x = _case()
    want = orc.nearest_chain(y, x, lat, lon, 2, rd, r)
    cut_at = None
    if cut == "inside_chain":
        # put the boundary where the sequential answer differs from the plain global answer
        dev = np.where((want != np.array([orc.edge_exclusion(*orc.nearest_point((y[t], x[t]), lat, lon, rd), *lat.shape, 2)
                                          for t in range(len(y))])).any(axis=1))[0]
        assert dev.size > 3, "the test track must contain a chain deviation"
        cut_at = int(dev[len(dev) // 2]) + 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, cut_at, q)) for rk in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, full, reseeds in got:
        np.testing.assert_array_equal(full, want)
    if cut == "inside_chain":
        assert max(rs for _, _, rs in got) >= 1      # the hand-off really had to re-seed rank 1


# ---------------------------------------------------------------------- BASELINE config 5 in small
def _missions():
    lat, lon = syn.orca_like_grid(146, 182)
    mask = syn.land_mask(lat, lon)
    res = syn.grid_resolution_deg(lon)
    rd, r = syn.search_params(res)
    tmod = 1800.0 * np.arange(4)
    fld = syn.ssh_records(lat, lon, 4)
    missions = []
    for k, (orb, n) in enumerate(((syn.SARAL, 700), (syn.JASON, 500), (syn.SARAL, 301))):
        t, y, x = syn.orbit_track(n, t0=40.0 + 900.0 * k, drop_land_seed=None, **orb)
        missions.append((t[::3].copy(), y[::3].copy(), x[::3].copy()))
    lat_, lon_, rd_, r_, ys, xs = _case()                       # the seam-hugging leg: a boundary inside a deviation
    missions.append((np.linspace(100.0, 5000.0, len(ys)), ys, xs))
    return lat, lon, mask, rd, r, tmod, fld, missions


def _mission_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gonzag_b200.multi import map_mission_batch
        lat, lon, mask, rd, r, tmod, fld, missions = _missions()
        eng = OracleEngine(lat, lon, 2)
        eng.mask = mask
        out = map_mission_batch(eng, dist, missions, rd, r, tmod=torch.from_numpy(tmod), slab=torch.from_numpy(fld), k0=0)
        q.put((rank, [(a.numpy(), b.numpy(), c.numpy()) for a, b, c in out]))
    except BaseException as exc:          # the parent must not wait for a result that will never come
        q.put((rank, repr(exc)))
        raise
    finally:
        dist.destroy_process_group()


def test_mission_batch_two_ranks_match_one_chain_per_mission():
    """Every rank maps its half of EVERY mission; the gathered products equal what one sequential pass
    over each mission alone gives (each mission its own chain, seeded (r, r))."""
    lat, lon, mask, rd, r, tmod, fld, missions = _missions()
    want = []
    for (t, y, x) in missions:
        NP = orc.nearest_chain(y, x, lat, lon, 2, rd, r)
        SM = orc.source_meshes(y, x, lat, lon, NP, 2)
        WB = orc.weights(y, x, lat, lon, SM)
        a, b = orc.interp_track(t, tmod, lambda k: fld[k], mask, SM, WB)
        want.append((a, b, NP))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mission_worker, args=(rk, 2, port, q)) for rk in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in got:
        assert not isinstance(res, str), res
        assert len(res) == len(want)
        for (a, b, NP), (wa, wb, wNP) in zip(res, want):
            np.testing.assert_array_equal(NP, wNP)
            np.testing.assert_array_equal(a, wa)
            np.testing.assert_array_equal(b, wb)
