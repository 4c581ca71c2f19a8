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

    def mesh_weights(self, y, x, NP, SM, WB):
        n = y.numel()
        self.ctx._ck(L.lib().gzb_source_mesh(self.ctx._h, n, self._p(y), self._p(x), self._p(NP), self._p(SM), L.GZB_DEVICE))
        self.ctx._ck(L.lib().gzb_weights(self.ctx._h, n, self._p(y), self._p(x), self._p(SM), self._p(WB), L.GZB_DEVICE))

    def interp(self, t, SM, WB, tmod, slab, k0, v_np, v_bl):
        dt = L.GZB_F32 if slab.dtype == self.torch.float32 else L.GZB_F64
        self.ctx._ck(L.lib().gzb_interp(self.ctx._h, t.numel(), self._p(t), self._p(SM), self._p(WB), tmod.numel(),
                                        self._p(tmod), self._p(slab), dt, int(k0), int(slab.shape[0]),
                                        float(config.rmissval), 1, self._p(v_np), self._p(v_bl), L.GZB_DEVICE))


class ShardedMapper:
    """Maps + interpolates one rank's segment and gathers the products of all ranks."""

    def __init__(self, engine, dist=None, torch_mod=None):
        self.e = engine
        self.dist = dist
        self.torch = torch_mod if torch_mod is not None else engine.torch
        self.rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.reseeds = 0

    # ---- K1 with the cross-segment seed hand-off -------------------------------------------
    def nearest_segment(self, y_halo, x_halo, has_halo, rd, r):
        """``y_halo/x_halo``: the rank's points, preceded by the previous rank's last point when
        ``has_halo``.  Returns NP of the rank's own points, (n,2) int64."""
        tc = self.torch
        n_all = y_halo.numel()
        NPh = self.e.empty((n_all, 2), tc.int64)
        # (0,0) is "no previous hit" in the reference (gonzag/bilin_mapping.py:157), so the halo
        # point gets its plain global answer, i.e. the speculative value of the chain there
        seed = (0, 0) if has_halo else (r, r)
        self.e.nearest(y_halo, x_halo, rd, r, seed, NPh)
        off = 1 if has_halo else 0
        NP = NPh[off:]
        if self.world == 1:
            return NP
        for _ in range(self.world):
            mine = tc.stack([NPh[0] if has_halo else tc.full((2,), -7, dtype=tc.int64, device=NPh.device),
                             NP[-1] if NP.shape[0] else tc.full((2,), -7, dtype=tc.int64, device=NPh.device)])
            allb = [tc.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(allb, mine)
            allb = [b.cpu() for b in allb]
            wrong = []
            for g in range(1, self.world):
                prev_last = allb[g - 1][1]
                if not bool((allb[g][0] == prev_last).all()):
                    wrong.append(g)
            if not wrong:
                break
            if self.rank in wrong:
                prev_last = allb[self.rank - 1][1]
                self.reseeds += 1
                self.e.nearest(y_halo[off:], x_halo[off:], rd, r, (int(prev_last[0]), int(prev_last[1])), NP)
                NPh[0] = prev_last.to(NPh.device)      # the halo now carries the true predecessor
        return NP

    # ---- final gather ------------------------------------------------------------------------
    def gather(self, tensors, bounds):
        """All-gather per-segment tensors (first dim = segment length) into full-track tensors."""
        if self.world == 1:
            return list(tensors)
        tc = self.torch
        nmax = max(bounds[g + 1] - bounds[g] for g in range(self.world))
        out = []
        for t in tensors:
            pad = tc.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[:t.shape[0]] = t
            parts = [tc.empty_like(pad) for _ in range(self.world)]
            self.dist.all_gather(parts, pad)
            out.append(tc.cat([parts[g][:bounds[g + 1] - bounds[g]] for g in range(self.world)], dim=0))
        return out
