"""`Model2SatTrack` over the GPUs of one box: one process per GPU (torchrun), every rank makes THE SAME call

    R = gonzag_b200.ShardedModel2SatTrack(MG, name_ssh_mod, ST, name_ssh_sat)          # inside an initialised process group

and rank ``root`` gets a result object with the reference's attributes over the WHOLE track (gonzag/mod2sat.py:153-186):
``time, lat, lon, mask, ssh_mod_np, ssh_mod, ssh_sat, distance, XNPtrack`` (+ ``BT.NP / SM / WB`` with
``keep_mapping=True``).  ``gather="all"`` gives every rank that object.

How the path shards:
  * the grid is REPLICATED -- rank ``grid_root`` uploads it once, the other ranks receive lat / lon / angle / mask over
    NVLink (one broadcast each) and build their search structures from the device copies (`replicate_grid`); a rank
    that already holds the grid (``ctx=`` or a `gonzag_b200.ModGrid`) skips this;
  * the track is cut into ``world`` contiguous time segments; rank g maps and interpolates segment g (`TrackJob`: K1-K4
    in one call) from the model records that bracket ITS segment (``MG.field`` may be a `RecordWindow`);
  * the only dependency between segments is the seed of a segment's first nearest-point search
    (gonzag/bilin_mapping.py:579-580): a rank first maps the last point of its predecessor alone, unseeded, which gives
    that point's value on the speculative chain, and seeds its segment with it; the ranks then exchange (assumed
    predecessor, own last point) -- 4 integers -- and a rank whose assumption was wrong maps its segment again from the
    true predecessor (a chain deviation straddling a boundary: rare), until every boundary agrees;
  * the along-track distance restarts at every segment from the predecessor's last point; the offsets (sum of the
    previous segments) are added after the gather;
  * one gather (to ``root``, or all-gather) of the per-segment products over NVLink: the two series, distance, mask and
    its complement, NP (16 B/point; + SM and WB on request: 128 B/point in all);
  * `XNPtrack` (last writer wins over the WHOLE track, gonzag/mod2sat.py:177-186) is computed from the gathered NP
    with GLOBAL track indices, on first access, on the rank that holds them.

torch.distributed (NCCL over NVLink; gloo in the CPU tests) is the plumbing: tensors are views of the library's device
buffers.  The kernels are the library's own; nothing here has a CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from . import _lib as L
from . import config
from .context import GridContext
from .mod2sat import LazyXNP, MappingView, TrackJob, _LazyXNPMixin, satellite_ssh, wrap_products
from .multi import segment_bounds
from .utils import SearchBoxSize


class _CAI:
    """`__cuda_array_interface__` wrapper so that torch can view raw device memory owned by a context."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def tensor_view(torch, tdev, ptr, nbytes):
    """uint8 tensor over `nbytes` of memory at `ptr` (device memory; host memory when the C ABI is emulated on CPU)."""
    if tdev.type == "cuda":
        return torch.as_tensor(_CAI(ptr, nbytes), device=tdev)
    buf = (C.c_char * int(nbytes)).from_address(int(ptr))
    return torch.from_numpy(np.frombuffer(buf, dtype=np.uint8))


def _dist_env(group):
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("ShardedModel2SatTrack needs an initialised torch.distributed process group (torchrun)")
    return torch, dist, dist.get_rank(group), dist.get_world_size(group)


def cap_host_threads():
    """One process per GPU: every rank gets its share of the cores for the library's host-side helpers."""
    if os.environ.get("GZB_HOST_THREADS"):
        return
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    if local_world > 1:
        L.lib().gzb_set_host_threads(max(2, (os.cpu_count() or 2) // local_world))


def replicate_grid(MG, device, group=None, root=0):
    """A `GridContext` of MG's grid on this rank's GPU.  Rank ``root`` uploads lat / lon / angle / mask from its host
    arrays (the only rank whose ``MG.lat, MG.lon, MG.xangle, MG.mask`` are read); the others receive the device copies
    over NVLink and build from them.  Returns (context, bytes received over the interconnect)."""
    torch, dist, rank, world = _dist_env(group)
    tdev = torch.device("cuda", device) if torch.cuda.is_available() else torch.device("cpu")
    (Nj, Ni) = MG.shape
    n = int(Nj) * int(Ni)
    ctx = None
    meta = torch.zeros(2, dtype=torch.int64, device=tdev)
    if rank == root:
        ang = MG.xangle if np.shape(getattr(MG, "xangle", ())) == (Nj, Ni) else None
        ctx = GridContext(MG.lat, MG.lon, angle=ang, mask=MG.mask, k_ew_per=MG.EWPer, device=device)
        meta[0] = 1 if ctx.has_angle else 0
        meta[1] = 1 if ctx.has_mask else 0
    if world == 1:
        return ctx, 0
    dist.broadcast(meta, src=root, group=group)
    has_angle, has_mask = bool(int(meta[0].item())), bool(int(meta[1].item()))
    sizes = [n * 8, n * 8, n * 8 if has_angle else 0, n if has_mask else 0]
    if rank == root:
        ptrs = ctx.grid_arrays()
        bufs = [tensor_view(torch, tdev, p, nb) if nb else None for p, nb in zip(ptrs, sizes)]
    else:
        bufs = [torch.empty(nb, dtype=torch.uint8, device=tdev) if nb else None for nb in sizes]
    moved = 0
    for b in bufs:
        if b is not None:
            dist.broadcast(b, src=root, group=group)
            moved += b.numel()
    if rank != root:
        if tdev.type == "cuda":
            torch.cuda.current_stream(tdev).synchronize()
        p = [b.data_ptr() if b is not None else None for b in bufs]
        ctx = GridContext.from_device(Nj, Ni, p[0], p[1], p[2], p[3], k_ew_per=MG.EWPer, device=device)
        del bufs
    return ctx, (moved if rank != root else 0)


class _TensorBuf:
    """DevBuf-like holder of a torch tensor (for `LazyXNP`)."""

    def __init__(self, t):
        self.t = t
        self.ptr = t.data_ptr()

    def free(self):
        self.t = None
        self.ptr = None


class ShardedModel2SatTrack(_LazyXNPMixin):

    def __init__(self, MG, name_ssh_mod, ST, name_ssh_sat, group=None, device=None, gather="root", root=0,
                 keep_mapping=False, ctx=None, grid_root=0, slab_bytes=256 << 20, resident_bytes=2 << 30):
        """Collective: every rank of ``group`` calls it with the same ``ST`` (the whole track) and the same grid
        description; only rank ``grid_root`` needs the grid ARRAYS of ``MG`` (unless ``ctx`` / ``MG.ctx`` gives every
        rank its resident grid).  ``MG.time`` is the whole record axis; ``MG.field`` / ``MG.get_record`` need to serve
        only the records that bracket this rank's time segment (see `RecordWindow`).  gather: "root" | "all"."""
        torch, dist, rank, world = _dist_env(group)
        if gather not in ("root", "all"):
            raise ValueError('gather must be "root" or "all"')
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if torch.cuda.is_available() else 0
        tdev = torch.device("cuda", device) if torch.cuda.is_available() else torch.device("cpu")
        self.rank, self.world, self.root = rank, world, root
        self.is_root = rank == root
        self.complete = self.is_root or gather == "all"
        lib = L.lib()
        cap_host_threads()
        (Nj, Ni) = MG.shape
        d_found_km = config.rfactor * MG.HResDeg * config.deg2km                    # mod2sat.py:35
        r = int(SearchBoxSize(MG.HResDeg * config.deg2km, config.search_box_w_km))  # :39
        Nt = int(ST.size)
        tim = {}
        tc = time.perf_counter()

        # ---- the grid, replicated
        if ctx is None:
            ctx = getattr(MG, "ctx", None)
        own_ctx = ctx is None
        self.interconnect_bytes = 0
        if ctx is None:
            ctx, self.interconnect_bytes = replicate_grid(MG, device, group, grid_root)
        elif not ctx.has_mask:
            ctx.set_mask(MG.mask)
        if tdev.type == "cuda":
            ctx.set_stream(torch.cuda.current_stream(tdev).cuda_stream)      # one stream for the kernels and the collectives
        h, ck = ctx._h, ctx._ck
        tim["context_s"] = time.perf_counter() - tc
        tc = time.perf_counter()

        # ---- this rank's time segment
        yt, xt, tt, tm = L.as_f64(ST.lat), L.as_f64(ST.lon), L.as_f64(ST.time), L.as_f64(MG.time)
        bounds = segment_bounds(Nt, world)
        s0, s1 = bounds[rank], bounds[rank + 1]
        n = s1 - s0
        nmax = max(bounds[g + 1] - bounds[g] for g in range(world))
        self.segment = (s0, s1)
        prev_point = (yt[s0 - 1], xt[s0 - 1]) if (s0 > 0 and n > 0) else None
        job = TrackJob(ctx, yt[s0:s1], xt[s0:s1], tt[s0:s1], tm, prev_point=prev_point, n_alloc=nmax)
        self._xnp = None
        self.reseeds = 0
        try:
            # the predecessor's last point, unseeded: its value on the speculative chain (E(t) of DESIGN.md K1)
            tim["job_alloc_h2d_s"] = time.perf_counter() - tc
            tq = time.perf_counter()
            if prev_point is not None:
                halo = ctx.nearest(yt[s0 - 1:s0], xt[s0 - 1:s0], d_found_km, r, seed=(0, 0))[0]
                seed = (int(halo[0]), int(halo[1]))
            else:
                seed = (r, r)                                                      # bilin_mapping.py:570
            tim["halo_s"] = time.perf_counter() - tq
            tq = time.perf_counter()
            kw = dict(slab_bytes=slab_bytes, resident_bytes=resident_bytes)
            if n > 0:
                job.run(MG, name_ssh_mod, d_found_km, r, seed, **kw)
            tim["run_issue_s"] = time.perf_counter() - tq
            tq = time.perf_counter()
            edges = torch.zeros((world, 4), dtype=torch.int64, device=tdev)
            mine = torch.zeros(4, dtype=torch.int64, device=tdev)
            last = np.zeros(2, dtype=np.int64)

            def my_edges():
                if n > 0:
                    ck(lib.gzb_d2h(h, L.ptr(last), C.c_void_p(job.P("np").value + (n - 1) * 16), 16))
                    ctx.sync()
                    v = [seed[0], seed[1], int(last[0]), int(last[1])]
                else:
                    v = [0, 0, 0, 0]
                mine.copy_(torch.tensor(v, dtype=torch.int64))

            # ---- seed hand-off: agree on every boundary (empty segments are transparent)
            for _ in range(world):
                my_edges()
                if world > 1:
                    dist.all_gather_into_tensor(edges.view(-1), mine, group=group)
                else:
                    edges[0] = mine
                e = edges.cpu().numpy()
                wrong = []
                prev_last = None
                for g in range(world):
                    if bounds[g + 1] - bounds[g] == 0:
                        continue
                    if prev_last is not None and (e[g, 0] != prev_last[0] or e[g, 1] != prev_last[1]):
                        wrong.append((g, (int(prev_last[0]), int(prev_last[1]))))
                    prev_last = e[g, 2:4]
                if not wrong:
                    break
                for g, true_seed in wrong:
                    if g == rank:                 # a chain deviation straddles my boundary: map again from the true predecessor
                        seed = true_seed
                        self.reseeds += 1
                        job.run(MG, name_ssh_mod, d_found_km, r, seed, **kw)
            else:
                raise RuntimeError("segment boundaries did not settle in %d rounds" % world)
            tim["run_wait_and_edges_s"] = time.perf_counter() - tq
            tq = time.perf_counter()

            # ---- a14 on the segment, then the gather
            ssh_attr = getattr(ST, "ssh", None)
            if ssh_attr is not None and not self.complete and np.shape(ssh_attr) == (Nt,):
                vssh_s = None                                   # this rank only needs its own slice: no whole-track copy
                my_sat = np.ascontiguousarray(np.ma.filled(np.ma.asarray(ssh_attr[s0:s1], dtype=np.float64), float(config.rmissval))
                                              if isinstance(ssh_attr, np.ma.MaskedArray) else ssh_attr[s0:s1], dtype=np.float64)
            else:
                vssh_s = satellite_ssh(ST, name_ssh_sat)
                if vssh_s.shape != (Nt,):
                    raise ValueError("the satellite series has %s values, the track %d points" % (vssh_s.shape, Nt))
                my_sat = np.ascontiguousarray(vssh_s[s0:s1])
            job.distance_and_mask(my_sat)
            tim["sat_distmask_s"] = time.perf_counter() - tq
            tim["segment_issue_s"] = time.perf_counter() - tc
            tc = time.perf_counter()
            off = job.off
            a, b = off["vnp"], (off["end"] if keep_mapping else off["sm"])
            mine_b = tensor_view(torch, tdev, job.d_all.ptr + a, b - a)
            parts = None
            if world == 1:
                parts = [mine_b]
            elif gather == "all":
                allb = torch.empty((world, b - a), dtype=torch.uint8, device=tdev)
                dist.all_gather_into_tensor(allb.view(-1), mine_b, group=group)
                parts = [allb[g] for g in range(world)]
            else:
                lst = [torch.empty(b - a, dtype=torch.uint8, device=tdev) for _ in range(world)] if self.is_root else None
                dist.gather(mine_b, lst, dst=root, group=group)
                parts = lst
            self.gathered_bytes = (b - a) * (world - 1) if (self.complete and world > 1) else 0
            tim["gather_s"] = time.perf_counter() - tc
            tc = time.perf_counter()
            self.time = ST.time
            self.lat = ST.lat
            self.lon = ST.lon
            if self.complete:
                def whole(key, item, skip=0):
                    """track-order concatenation of one product over the segments (a device tensor of bytes)"""
                    o = off[key] - a + skip
                    return torch.cat([parts[g][o:o + (bounds[g + 1] - bounds[g]) * item] for g in range(world)])

                d_np, d_bl, d_dist = whole("vnp", 8), whole("vbl", 8), whole("dist", 8, 8).view(torch.float64)
                # distance: every segment counted from its predecessor's last point -> add the sum of the segments before
                carry = 0.0
                tails = []
                for g in range(world):
                    ng = bounds[g + 1] - bounds[g]
                    if ng:
                        tails.append(d_dist[bounds[g + 1] - 1:bounds[g + 1]])
                tails = torch.cat(tails).cpu().numpy() if tails else np.zeros(0)
                k = 0
                for g in range(world):
                    ng = bounds[g + 1] - bounds[g]
                    if not ng:
                        continue
                    if carry != 0.0:
                        d_dist[bounds[g]:bounds[g + 1]] += carry
                    carry += float(tails[k])
                    k += 1
                d_im, d_bad, d_NP = whole("imask", 1), whole("bad", 1), whole("np", 16)
                series = torch.cat([d_np, d_bl, d_dist.view(torch.uint8), d_im, d_bad])
                hbuf = np.empty(series.numel(), dtype=np.uint8)
                if series.numel():
                    ck(lib.gzb_d2h(h, L.ptr(hbuf), C.c_void_p(series.data_ptr()), hbuf.nbytes))
                if keep_mapping:
                    d_SM, d_WB = whole("sm", 64), whole("wb", 32)
                    mcat = torch.cat([d_NP, d_SM, d_WB])
                    mbuf = np.empty(mcat.numel(), dtype=np.uint8)
                    if mcat.numel():
                        ck(lib.gzb_d2h(h, L.ptr(mbuf), C.c_void_p(mcat.data_ptr()), mbuf.nbytes))
                    self.BT = MappingView(mbuf[:Nt * 16].view(np.int64).reshape(Nt, 2),
                                          mbuf[Nt * 16:Nt * 80].view(np.int64).reshape(Nt, 4, 2),
                                          mbuf[Nt * 80:].view(np.float64).reshape(Nt, 4))
                ctx.sync()
                o1, o2, o3, o4 = Nt * 8, Nt * 16, Nt * 24, Nt * 25
                wrap_products(self, hbuf[:o1].view(np.float64), hbuf[o1:o2].view(np.float64), vssh_s,
                              hbuf[o3:o4].view(np.int8), hbuf[o4:].view(np.bool_), hbuf[o2:o3].view(np.float64))
                if config.l_save_track_on_model_grid:
                    self._xnp = LazyXNP(ctx, own_ctx, _TensorBuf(d_NP.contiguous()), Nt)
                    own_ctx = False
            else:
                ctx.sync()
            self.fused_path = job.fused
        finally:
            job.free()
        self.stats = ctx.stats()
        tim["outputs_s"] = time.perf_counter() - tc
        self.timing = tim
        if own_ctx:
            ctx.close()
