"""Track sharding across the GPUs of one box.

The model grid is replicated (one `GridContext` per process / GPU); the track is cut into
contiguous time segments, one per rank.  The only dependency between segments is the
nearest-point seed of a segment's first point (gonzag/bilin_mapping.py:579-580): every rank
maps its segment preceded by a one-point halo, assuming that halo point obeys the
speculative chain; the ranks then exchange 2 x int64 per boundary, and a rank whose assumption
turns out wrong re-maps its segment from the true predecessor (rare: only when a chain
deviation straddles the boundary).  The single data-path collective is the final all-gather
of the per-segment products.

The driver is written against a small `engine` interface so that its control flow can be
tested on CPU (gloo, world_size 2) with the oracle plugged in; the product engine is
`CudaEngine` and has no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from . import config


def segment_bounds(nt: int, world: int):
    """Contiguous, equal-count segments: rank g owns points [b[g], b[g+1])."""
    base, rem = divmod(int(nt), int(world))
    b = [0]
    for g in range(world):
        b.append(b[-1] + base + (1 if g < rem else 0))
    return b


class CudaEngine:
    """Segment-level compute on one GPU; arrays are torch CUDA tensors (so that NCCL can move
    them) whose storage is handed to the C ABI as raw device pointers."""

    def __init__(self, ctx, torch_mod):
        self.ctx = ctx
        self.torch = torch_mod
        self.dev = torch_mod.device("cuda", ctx.device)
        ctx.set_stream(torch_mod.cuda.current_stream(self.dev).cuda_stream)

    def tensor(self, arr, dtype=None):
        t = self.torch.as_tensor(np.ascontiguousarray(arr))
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.dev, non_blocking=False)

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, device=self.dev)

    def _p(self, t):
        return C.c_void_p(t.data_ptr())

    def nearest(self, y, x, rd, r, seed, out):
        self.ctx._ck(L.lib().gzb_nearest(self.ctx._h, y.numel(), self._p(y), self._p(x), float(rd), int(r),
                                         int(seed[0]), int(seed[1]), self._p(out), L.GZB_DEVICE))

    def mapping(self, y, x, rd, r, seed, NP, SM, WB):
        """K1 + K2 + K3 in one call: the chain replays of K1 run beside K2 + K3."""
        self.ctx._ck(L.lib().gzb_mapping(self.ctx._h, y.numel(), self._p(y), self._p(x), float(rd), int(r),
                                         int(seed[0]), int(seed[1]), self._p(NP), self._p(SM), self._p(WB),
                                         L.GZB_DEVICE))

    def map_interp(self, y, x, t, rd, r, seed, tmod, slab, k0, NP, SM, WB, v_np, v_bl):
        """K1 + K2 + K3 + K4 in one call (the chain replays of K1 run beside K2-K4)."""
        dt = L.GZB_F32 if slab.dtype == self.torch.float32 else L.GZB_F64
        self.ctx._ck(L.lib().gzb_map_interp(self.ctx._h, y.numel(), self._p(y), self._p(x), self._p(t), float(rd), int(r),
                                            int(seed[0]), int(seed[1]), tmod.numel(), self._p(tmod), self._p(slab), dt,
                                            int(k0), int(slab.shape[0]), float(config.rmissval), self._p(NP),
                                            self._p(SM), self._p(WB), self._p(v_np), self._p(v_bl)))

    def map_interp_multi(self, seg_start, seg_seed, y, x, t, rd, r, tmod, slab, k0, NP, SM, WB, v_np, v_bl):
        """K1-K4 of several independent chains in one call: sub-track s starts at point seg_start[s] (seg_start[0] == 0) and is
        seeded with seg_seed[s] (host int64 arrays) -- several missions inside one window of records."""
        dt = L.GZB_F32 if slab.dtype == self.torch.float32 else L.GZB_F64
        st = np.ascontiguousarray(seg_start, dtype=np.int64)
        sd = np.ascontiguousarray(seg_seed, dtype=np.int64).reshape(len(st), 2)
        self.ctx._ck(L.lib().gzb_map_interp_multi(self.ctx._h, len(st), L.ptr(st), L.ptr(sd), y.numel(), self._p(y), self._p(x),
                                                  self._p(t), float(rd), int(r), tmod.numel(), self._p(tmod), self._p(slab), dt,
                                                  int(k0), int(slab.shape[0]), float(config.rmissval), self._p(NP), self._p(SM),
                                                  self._p(WB), self._p(v_np), self._p(v_bl)))

    def mesh_weights_interp(self, y, x, t, NP, tmod, slab, k0, SM, WB, v_np, v_bl):
        """K2 + K3 + K4 as one kernel from known nearest points (the bulk kernel of `map_interp`)."""
        dt = L.GZB_F32 if slab.dtype == self.torch.float32 else L.GZB_F64
        self.ctx._ck(L.lib().gzb_mesh_weights_interp(self.ctx._h, y.numel(), self._p(y), self._p(x), self._p(t), self._p(NP),
                                                     tmod.numel(), self._p(tmod), self._p(slab), dt, int(k0),
                                                     int(slab.shape[0]), float(config.rmissval), self._p(SM), self._p(WB),
                                                     self._p(v_np), self._p(v_bl)))

    def mesh_weights(self, y, x, NP, SM, WB):
        n = y.numel()
        self.ctx._ck(L.lib().gzb_mesh_weights(self.ctx._h, n, self._p(y), self._p(x), self._p(NP), self._p(SM),
                                              self._p(WB), L.GZB_DEVICE))

    def interp(self, t, SM, WB, tmod, slab, k0, v_np, v_bl):
        dt = L.GZB_F32 if slab.dtype == self.torch.float32 else L.GZB_F64
        self.ctx._ck(L.lib().gzb_interp(self.ctx._h, t.numel(), self._p(t), self._p(SM), self._p(WB), tmod.numel(),
                                        self._p(tmod), self._p(slab), dt, int(k0), int(slab.shape[0]),
                                        float(config.rmissval), 1, self._p(v_np), self._p(v_bl), L.GZB_DEVICE))


class ShardedMapper:
    """Maps one rank's segment and gathers the products of all ranks.

    Product buffer of a rank (one allocation, so that ONE `all_gather_into_tensor` moves it):
    int64 words ``[(nmax+1)*4]`` = ``ssh_np`` (f64, nmax+1) | ``ssh_bl`` (f64, nmax+1) | ``NP``
    (i64, (nmax+1)*2).  Row 0 of every section is the slot of the halo point (the last point of the
    previous segment, mapped again here only to seed the chain), rows 1..n are the rank's own
    points: the kernels write straight into views of it, halo included, and nothing is copied."""

    def __init__(self, engine, dist=None, torch_mod=None):
        self.e = engine
        self.dist = dist
        self.torch = torch_mod if torch_mod is not None else engine.torch
        self.rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.reseeds = 0

    # ---- buffers -----------------------------------------------------------------------------
    def make_buffers(self, bounds):
        tc = self.torch
        self.bounds = list(bounds)
        self.nmax = max(bounds[g + 1] - bounds[g] for g in range(self.world))
        n = bounds[self.rank + 1] - bounds[self.rank]
        self.n = n
        self.halo = self.rank > 0
        m1 = self.nmax + 1
        self.m1 = m1
        self.pack = self.e.empty((4 * m1,), tc.int64)
        self.pack.zero_()
        self.v_np_h = self.pack[:m1].view(tc.float64)[:n + 1]              # halo slot + own rows
        self.v_bl_h = self.pack[m1:2 * m1].view(tc.float64)[:n + 1]
        self.NPh = self.pack[2 * m1:].view(m1, 2)[:n + 1]
        self.v_np, self.v_bl, self.NP = self.v_np_h[1:], self.v_bl_h[1:], self.NPh[1:]
        self.SMh = self.e.empty((n + 1, 4, 2), tc.int64)
        self.WBh = self.e.empty((n + 1, 4), tc.float64)
        self.SM, self.WB = self.SMh[1:], self.WBh[1:]
        self.gathered = self.e.empty((self.world, 4 * m1), tc.int64) if self.world > 1 else None
        self.flag = self.e.empty((1,), tc.int32)
        return self

    def _rows(self, t):
        """The rows an engine call on [halo +] own points writes: from the halo slot, or from row 1."""
        return t if self.halo else t[1:]

    # ---- K1 with the cross-segment seed hand-off -----------------------------------------------
    def _seed(self, r):
        # (0,0) is "no previous hit" in the reference (gonzag/bilin_mapping.py:157), so the halo point
        # gets its plain global answer, i.e. the speculative value of the chain there
        return (0, 0) if self.halo else (r, r)

    def nearest_speculative(self, y_halo, x_halo, rd, r):
        """K1 on [halo point +] own points.  The result lands in ``self.NP`` (halo in ``self.NPh[0]``)."""
        self.e.nearest(y_halo, x_halo, rd, r, self._seed(r), self._rows(self.NPh))

    def map_speculative(self, y_halo, x_halo, rd, r):
        """K1 + K2 + K3 on [halo point +] own points in one engine call: ``self.NP/SM/WB``."""
        self.e.mapping(y_halo, x_halo, rd, r, self._seed(r), self._rows(self.NPh), self._rows(self.SMh),
                       self._rows(self.WBh))

    def path_speculative(self, y_halo, x_halo, t_halo, rd, r, tmod, slab, k0):
        """K1 + K2 + K3 + K4 on [halo point +] own points in one engine call: ``self.NP/SM/WB``
        and the interpolated series ``self.v_np / self.v_bl``."""
        self.e.map_interp(y_halo, x_halo, t_halo, rd, r, self._seed(r), tmod, slab, k0, self._rows(self.NPh),
                          self._rows(self.SMh), self._rows(self.WBh), self._rows(self.v_np_h), self._rows(self.v_bl_h))

    def gather(self):
        """The one data-path collective: all-gather of the packed products."""
        if self.world == 1:
            return self.pack.view(1, -1)
        self.dist.all_gather_into_tensor(self.gathered.view(-1), self.pack)
        return self.gathered

    def _last_np_of(self, gathered, g):
        last = self.bounds[g + 1] - self.bounds[g]          # own rows start at 1
        return gathered[g, 2 * self.m1:].view(self.m1, 2)[last]

    def boundaries_ok(self, gathered):
        """Does every rank's halo assumption equal the true last point of the previous rank?
        One tiny all-reduce + one host read per step."""
        if self.world == 1:
            return True
        tc = self.torch
        if self.halo:
            self.flag[0] = (self._last_np_of(gathered, self.rank - 1) != self.NPh[0]).any().to(tc.int32)
        else:
            self.flag.zero_()
        self.dist.all_reduce(self.flag, op=self.dist.ReduceOp.MAX)
        return int(self.flag.item()) == 0

    def redo_own_points(self, y_own, x_own, t_own, tmod, slab, k0):
        """K2 + K3 + K4 on this rank's own points (after `reseed` changed their nearest points)."""
        self.e.mesh_weights(y_own, x_own, self.NP, self.SM, self.WB)
        self.e.interp(t_own, self.SM, self.WB, tmod, slab, k0, self.v_np, self.v_bl)

    def check_boundaries_async(self, gathered):
        """Device-side version of `boundaries_ok`, no collective and no host read: every rank's halo
        assumption travels in row 0 of its NP section, so after the all-gather each rank can compare
        all boundaries itself.  The verdict is OR-ed into ``self.flag_acc`` (a device int32) and read
        later with `boundaries_verdict()` -- a pipeline checks once per batch of steps, like an
        asynchronous error flag, and re-seeds (`reseed`) only if it is ever set."""
        tc = self.torch
        if getattr(self, "flag_acc", None) is None:
            self.flag_acc = self.e.empty((1,), tc.int32)
            self.flag_acc.zero_()
            self._last_rows = tc.as_tensor([self.bounds[g + 1] - self.bounds[g] for g in range(self.world - 1)],
                                           device=self.flag_acc.device) if self.world > 1 else None
        if self.world == 1:
            return
        NPs = gathered[:, 2 * self.m1:].view(self.world, self.m1, 2)
        prev_last = NPs[tc.arange(self.world - 1, device=NPs.device), self._last_rows]     # (world-1, 2)
        halo = NPs[1:, 0]                                                                  # (world-1, 2)
        self.flag_acc |= (prev_last != halo).any().to(tc.int32)

    def boundaries_verdict(self):
        """Host read of the accumulated flag (one synchronisation); True = every step was valid."""
        if getattr(self, "flag_acc", None) is None:
            return True
        if getattr(self, "root_only", False) and self.world > 1:      # only rank 0 saw every edge: share its verdict
            self.dist.all_reduce(self.flag_acc, op=self.dist.ReduceOp.MAX)
        bad = int(self.flag_acc.item())
        self.flag_acc.zero_()
        return bad == 0

    def reseed(self, gathered, y_own, x_own, rd, r):
        """Rare path: re-map the segments whose predecessor assumption was wrong, in rank order,
        until every boundary agrees (at most world-1 rounds).  Returns True if THIS rank re-mapped
        (its mesh / weights / interpolation must then be redone by the caller)."""
        changed = False
        for _ in range(self.world):
            if self.halo:
                prev_last = self._last_np_of(gathered, self.rank - 1).clone()
                if bool((prev_last != self.NPh[0]).any()):
                    self.reseeds += 1
                    changed = True
                    pl = prev_last.cpu()
                    self.e.nearest(y_own, x_own, rd, r, (int(pl[0]), int(pl[1])), self.NP)
                    self.NPh[0] = prev_last
            gathered = self.gather()
            if self.boundaries_ok(gathered):
                break
        return changed

    # ---- views of the gathered products -------------------------------------------------------
    def unpack(self, gathered):
        """Full-track (ssh_np, ssh_bl, NP) tensors in track order (concatenation of the shards)."""
        tc = self.torch
        m1 = self.m1
        parts_np, parts_bl, parts_NP = [], [], []
        for g in range(self.world):
            n = self.bounds[g + 1] - self.bounds[g]
            row = gathered[g]
            parts_np.append(row[:m1].view(tc.float64)[1:n + 1])
            parts_bl.append(row[m1:2 * m1].view(tc.float64)[1:n + 1])
            parts_NP.append(row[2 * m1:].view(m1, 2)[1:n + 1])
        return tc.cat(parts_np), tc.cat(parts_bl), tc.cat(parts_NP)


def map_mission_batch(engine, dist, missions, rd, r, tmod=None, slab=None, k0=0, torch_mod=None):
    """BASELINE config 5 (several altimeter missions onto one model grid, strong scaling): every mission
    is cut into `world` contiguous time segments and rank g maps segment g of EVERY mission, one sharded
    step per mission -- so each GPU gets 1/world of each mission and the grid context,
    the model records and the kernels' workspace stay resident across missions.  A mission is a separate
    chain: its first point is seeded with (r, r) exactly as a `BilinTrack` of its own would be
    (gonzag/bilin_mapping.py:570); within a mission the segments hand their seed over as in
    `ShardedMapper`.

    missions: list of ``(t, y, x)`` host arrays (``t`` may be None for mapping only);
    tmod / slab / k0: model record times (tensor), records ``k0 ..`` (tensor) -- None: K1 only.
    Returns one ``(ssh_np, ssh_bl, NP)`` per mission, full track order, on every rank (tensors; the
    two series are None without records)."""
    tc = torch_mod if torch_mod is not None else engine.torch
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    results = []
    for (t, y, x) in missions:
        n = len(y)
        if n < 2 * world:
            raise ValueError("a mission of %d points cannot be cut into %d segments" % (n, world))
        bounds = segment_bounds(n, world)
        s0, s1 = bounds[rank], bounds[rank + 1]
        h0 = s0 - 1 if rank > 0 else s0
        off = s0 - h0
        m = ShardedMapper(engine, dist if world > 1 else None, tc).make_buffers(bounds)
        yh, xh = engine.tensor(y[h0:s1]), engine.tensor(x[h0:s1])
        with_k4 = slab is not None and t is not None
        if with_k4:
            th = engine.tensor(t[h0:s1])
            m.path_speculative(yh, xh, th, rd, r, tmod, slab, k0)
        else:
            m.nearest_speculative(yh, xh, rd, r)
        g = m.gather()
        if not m.boundaries_ok(g):                      # a chain deviation straddles a segment boundary
            if m.reseed(g, yh[off:], xh[off:], rd, r) and with_k4:
                m.redo_own_points(yh[off:], xh[off:], th[off:], tmod, slab, k0)
            g = m.gather()
        v_np, v_bl, NP = m.unpack(g)
        results.append((v_np.clone() if with_k4 else None, v_bl.clone() if with_k4 else None, NP.clone()))
    return results


class P2PUnavailable(RuntimeError):
    """Raised by `P2PShardedMapper.make_buffers` ON EVERY RANK when some rank cannot map a peer's buffer;
    the caller falls back to `ShardedMapper` (NCCL all-gather)."""


class _DevView:
    """`__cuda_array_interface__` wrapper so that torch can view raw device memory owned by a context."""

    def __init__(self, ptr, nwords):
        self.__cuda_array_interface__ = {"shape": (int(nwords),), "typestr": "<i8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class P2PShardedMapper(ShardedMapper):
    """`ShardedMapper` whose gather is fused into K4: every rank maps the product buffers of the
    other ranks (CUDA IPC) and K4 stores the two interpolated series into all of them, over NVLink,
    as it computes them; the collective left is a one-word all-reduce that orders the ranks.

    Buffer of a rank, two of them used alternately (a rank may start writing step k+1 into a peer
    while that peer still reads step k): ``[world][2*(nmax+1) + 4]`` int64 words -- per source rank
    ``ssh_np`` (f64, nmax+1, row 0 = halo slot) | ``ssh_bl`` | 4 edge words (halo NP assumption,
    last NP of the segment).  NP, SM and WB stay on the rank that computed them."""

    def make_buffers(self, bounds):
        tc = self.torch
        ctx = self.e.ctx
        lib = L.lib()
        self.bounds = list(bounds)
        self.nmax = max(bounds[g + 1] - bounds[g] for g in range(self.world))
        n = bounds[self.rank + 1] - bounds[self.rank]
        self.n = n
        self.halo = self.rank > 0
        m1 = self.nmax + 1
        self.m1 = m1
        self.slot = 2 * m1 + 4
        self.NPh = self.e.empty((n + 1, 2), tc.int64)
        self.NPh.zero_()
        self.NP = self.NPh[1:]
        self.SMh = self.e.empty((n + 1, 4, 2), tc.int64)
        self.WBh = self.e.empty((n + 1, 4), tc.float64)
        self.SM, self.WB = self.SMh[1:], self.WBh[1:]
        self.flag = self.e.empty((1,), tc.int32)
        self.flag.zero_()
        nbytes = (self.world * self.slot + self.world) * 8        # + the epoch counters of the peer barrier
        self._bufs = [ctx.dev_alloc(nbytes) for _ in range(2)]
        for b in self._bufs:
            ctx._ck(lib.gzb_memset(ctx._h, C.c_void_p(b.ptr), 0, nbytes))
        ctx.sync()
        self.gath = [tc.as_tensor(_DevView(b.ptr, self.world * self.slot), device=self.e.dev).view(self.world, self.slot)
                     for b in self._bufs]
        # exchange the IPC handles and map the peers' buffers
        mine = []
        for b in self._bufs:
            hbuf = (C.c_ubyte * 64)()
            ctx._ck(lib.gzb_ipc_export(ctx._h, C.c_void_p(b.ptr), hbuf))
            mine.append(bytes(hbuf))
        allh = [None] * self.world
        if self.world > 1:
            self.dist.all_gather_object(allh, mine)
        else:
            allh[0] = mine
        self._peer_base = []          # [buffer][peer index] -> mapped base address on this device
        # where this rank's series go besides its own buffer: every other rank (all-gather: everybody ends up
        # with the whole track), or rank 0 only (gather: ``root_only``, what writing one result file needs)
        if getattr(self, "root_only", False):
            self._peers = [0] if self.rank != 0 else []
        else:
            self._peers = [g for g in range(self.world) if g != self.rank]
        ok = 1
        try:
            for bi in range(2):
                row = []
                self._peer_base.append(row)
                for g in self._peers:
                    p = C.c_void_p()
                    ctx._ck(lib.gzb_ipc_open(ctx._h, (C.c_ubyte * 64).from_buffer_copy(allh[g][bi]), C.byref(p)))
                    row.append(p.value)
        except L.GonzagB200Error as exc:
            ok = 0
            why = str(exc)
        if self.world > 1:        # every rank must take the same decision (this all-reduce is also the barrier)
            okt = tc.tensor([ok], dtype=tc.int32, device=self.e.dev)
            self.dist.all_reduce(okt, op=self.dist.ReduceOp.MIN)
            if int(okt.item()) == 0:
                self.close()
                raise P2PUnavailable("peer memory (CUDA IPC) is not available between the GPUs of this job" +
                                     (": " + why if not ok else ""))
        # barrier: this rank's entry in every peer's counters (kept in buffer 0), its own counters
        self._bar_slots = (C.c_void_p * max(len(self._peers), 1))()
        for k in range(len(self._peers)):
            self._bar_slots[k] = self._peer_base[0][k] + (self.world * self.slot + self.rank) * 8
        self._bar_local = C.c_void_p(self._bufs[0].ptr + self.world * self.slot * 8)
        self._epoch = 0
        self.step_no = 0
        self._select(0)
        return self

    def _selection(self, bi, off):
        """Views and pointer tables for buffer `bi`, kernels writing from row `off` of this rank's slot."""
        tc = self.torch
        m1, slot, r = self.m1, self.slot, self.rank
        mine = self.gath[bi][r]
        v_np_h = mine[:m1].view(tc.float64)[:self.n + 1]
        v_bl_h = mine[m1:2 * m1].view(tc.float64)[:self.n + 1]
        npk = len(self._peers)
        arr_np = (C.c_void_p * max(npk, 1))()
        arr_bl = (C.c_void_p * max(npk, 1))()
        edge_dst = (C.c_void_p * (npk + 1))()
        for k in range(npk):
            base = self._peer_base[bi][k] + r * slot * 8
            arr_np[k] = base + off * 8
            arr_bl[k] = base + (m1 + off) * 8
            edge_dst[k] = base + 2 * m1 * 8
        edge_dst[npk] = self._bufs[bi].ptr + (r * slot + 2 * m1) * 8
        return dict(v_np_h=v_np_h, v_bl_h=v_bl_h, arr_np=arr_np, arr_bl=arr_bl, edge_dst=edge_dst, npk=npk,
                    loc_np=C.c_void_p(v_np_h[off:].data_ptr()), loc_bl=C.c_void_p(v_bl_h[off:].data_ptr()))

    def _select(self, bi, first_row=None):
        """Point the kernels at buffer `bi`: local views, and the peers' slots for this rank starting
        at `first_row` (0: the call also covers the halo point, 1: own points only)."""
        off = (0 if self.halo else 1) if first_row is None else first_row
        cache = self.__dict__.setdefault("_sel_cache", {})
        sel = cache.get((bi, off))
        if sel is None:
            sel = cache[(bi, off)] = self._selection(bi, off)
        self.v_np_h, self.v_bl_h = sel["v_np_h"], sel["v_bl_h"]
        self.v_np, self.v_bl = self.v_np_h[1:], self.v_bl_h[1:]
        self.cur = bi
        self._edge_dst = sel["edge_dst"]
        ctx = self.e.ctx
        ctx._ck(L.lib().gzb_set_peer_outputs(ctx._h, sel["npk"], sel["arr_np"], sel["arr_bl"], sel["loc_np"], sel["loc_bl"]))

    def _publish(self):
        ctx = self.e.ctx
        ctx._ck(L.lib().gzb_publish_edges(ctx._h, C.c_void_p(self.NPh.data_ptr()), self.n + 1, len(self._peers) + 1,
                                          self._edge_dst))

    def _fused_finish(self):
        """barrier "p2p": edge publication, barrier and boundary check are ONE kernel (gzb_peer_finish)"""
        return getattr(self, "barrier", "nccl") == "p2p" and self.world > 1 and not getattr(self, "root_only", False)

    def path_speculative(self, y_halo, x_halo, t_halo, rd, r, tmod, slab, k0):
        """K1-K4 on [halo +] own points; the two series go here and, chunk by chunk behind the bulk kernel, into
        every peer (k_push on its own stream)."""
        self._select(self.step_no & 1)
        self.step_no += 1
        super().path_speculative(y_halo, x_halo, t_halo, rd, r, tmod, slab, k0)
        if not self._fused_finish():
            self._publish()

    def gather(self):
        """No data moves here: a barrier orders the ranks (``barrier="p2p"``: one stream-ordered kernel over peer
        memory that also publishes the edge words and checks the boundaries; ``"nccl"``: NCCL's one-word
        all-reduce), after it every rank's buffer holds the series of all segments."""
        if self.world > 1:
            if self._fused_finish():
                ctx = self.e.ctx
                if getattr(self, "flag_acc", None) is None:
                    self.flag_acc = self.e.empty((1,), self.torch.int32)
                    self.flag_acc.zero_()
                self._epoch += 1
                g = self.gath[self.cur]
                ctx._ck(L.lib().gzb_peer_finish(ctx._h, C.c_void_p(self.NPh.data_ptr()), self.n + 1, len(self._peers) + 1,
                                                self._edge_dst, len(self._peers), self._bar_slots, self._bar_local, self.world,
                                                self.rank, self._epoch, C.c_void_p(g.data_ptr()), self.slot, 2 * self.m1,
                                                C.c_void_p(self.flag_acc.data_ptr())))
                self._checked = True
            elif getattr(self, "barrier", "nccl") == "nccl":
                self.dist.all_reduce(self.flag, op=self.dist.ReduceOp.MAX)
            else:
                ctx = self.e.ctx
                self._epoch += 1
                ctx._ck(L.lib().gzb_peer_barrier(ctx._h, len(self._peers), self._bar_slots, self._bar_local, self.world,
                                                 self.rank, self._epoch))
        return self.gath[self.cur]

    def _edges(self, gathered):
        return gathered[:, 2 * self.m1:2 * self.m1 + 4]          # (world, 4): halo j,i, last j,i

    def check_boundaries_async(self, gathered):
        tc = self.torch
        if getattr(self, "flag_acc", None) is None:
            self.flag_acc = self.e.empty((1,), tc.int32)
            self.flag_acc.zero_()
        if self.world == 1 or (getattr(self, "root_only", False) and self.rank != 0):
            return                          # gather to root: only rank 0 holds every segment's edge words
        if getattr(self, "_checked", False):      # the fused finish kernel of `gather` has already compared the edges
            self._checked = False
            return
        ctx = self.e.ctx
        ctx._ck(L.lib().gzb_check_edges(ctx._h, C.c_void_p(gathered.data_ptr()), self.slot, 2 * self.m1, self.world,
                                        C.c_void_p(self.flag_acc.data_ptr())))

    def boundaries_ok(self, gathered):
        if self.world == 1:
            return True
        if getattr(self, "root_only", False):
            self.check_boundaries_async(gathered)
            return self.boundaries_verdict()
        e = self._edges(gathered)
        return not bool((e[1:, 0:2] != e[:-1, 2:4]).any().item())

    def _last_np_of(self, gathered, g):
        return self._edges(gathered)[g, 2:4]

    def reseed(self, gathered, y_own, x_own, rd, r):
        """Rare path, as in the base class; the edge words are published again after a re-map.  The
        caller then redoes K2-K4 on its own points with `redo_own_points`."""
        if getattr(self, "root_only", False):
            raise RuntimeError("a chain deviation straddles a segment boundary: re-seeding needs every rank to see its "
                               "predecessor's last point, i.e. the all-gather mode (root_only = False)")
        changed = False
        for _ in range(self.world):
            if self.halo:
                prev_last = self._last_np_of(gathered, self.rank - 1).clone()
                if bool((prev_last != self.NPh[0]).any()):
                    self.reseeds += 1
                    changed = True
                    pl = prev_last.cpu()
                    self.e.nearest(y_own, x_own, rd, r, (int(pl[0]), int(pl[1])), self.NP)
                    self.NPh[0] = prev_last
            self._publish()
            gathered = self.gather()
            if self.boundaries_ok(gathered):
                break
        if getattr(self, "flag_acc", None) is not None:      # the rounds above may have flagged boundaries that are settled now
            self.flag_acc.zero_()
        self._checked = False
        return changed

    def redo_own_points(self, y_own, x_own, t_own, tmod, slab, k0):
        """K2 + K3 + K4 on this rank's own points (after `reseed` changed their nearest points); K4
        writes into the peers again."""
        self._select(self.cur, first_row=1)
        self.e.mesh_weights(y_own, x_own, self.NP, self.SM, self.WB)
        self.e.interp(t_own, self.SM, self.WB, tmod, slab, k0, self.v_np, self.v_bl)

    def unpack(self, gathered):
        """Full-track (ssh_np, ssh_bl) in track order; NP stays sharded (None)."""
        tc = self.torch
        m1 = self.m1
        parts_np, parts_bl = [], []
        for g in range(self.world):
            n = self.bounds[g + 1] - self.bounds[g]
            row = gathered[g]
            parts_np.append(row[:m1].view(tc.float64)[1:n + 1])
            parts_bl.append(row[m1:2 * m1].view(tc.float64)[1:n + 1])
        return tc.cat(parts_np), tc.cat(parts_bl), None

    def detach(self):
        """Stop replicating K4's outputs into the peers (local work only from here on)."""
        ctx = self.e.ctx
        if ctx._h is not None:
            ctx._ck(L.lib().gzb_set_peer_outputs(ctx._h, 0, None, None, None, None))

    def close(self):
        ctx = self.e.ctx
        if ctx._h is None:
            return
        L.lib().gzb_set_peer_outputs(ctx._h, 0, None, None, None, None)
        for row in getattr(self, "_peer_base", []):
            for p in row:
                L.lib().gzb_ipc_close(ctx._h, C.c_void_p(p))
        self._peer_base = []
        if self.world > 1:
            self.dist.barrier()
        for b in getattr(self, "_bufs", []):
            b.free()
        self._bufs = []
