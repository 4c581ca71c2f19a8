"""Drop-in for the reference's gonzag/bilin_mapping.py: same class, signature and attributes,
the three per-point Python loops replaced by CUDA kernels reached through the C ABI.

    BilinTrack(Yt, Xt, Ys, Xs, src_grid_local_angle=[], k_ew_per=-1, rd_found_km=100.,
               np_box_r=4, freq_talk=0)                    (gonzag/bilin_mapping.py:533-564)
      .NP (Nt,2) int64   nearest point [j,i], (-1,-1) = none           (:566-593)
      .SM (Nt,4,2) int64 source mesh, counter-clockwise from NP        (:595-608)
      .WB (Nt,4) float64 bilinear weights                              (:610-617)

Differences a caller can observe: coordinates are converted to float64 on entry (the reference
computes in whatever dtype the arrays have); nothing is printed unless ``config.ivrb > 0``.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import config
from .context import GridContext

# ---------------------------------------------------------------------------------------------
# The free functions the reference package re-exports (gonzag/__init__.py:13).  Each is a
# single-point call into the same CUDA kernels (no Python restatement of the arithmetic); the
# grid context of the last (Ys, Xs) pair is kept so that loops over points do not re-upload it.
_CTX_CACHE = {"key": None, "ctx": None}


def _fingerprint(A):
    """Cheap content fingerprint of a grid array: its corners plus a strided sample of ~4096 values.  The key of the
    context cache must change when a caller edits Ys / Xs in place, or when a freed array's id and address are reused by
    another grid of the same shape (the reference recomputes from the arrays on every call)."""
    a = np.asarray(A)
    flat = a.reshape(-1)
    step = max(1, flat.size // 4096)
    sample = np.ascontiguousarray(flat[::step])
    edge = np.ascontiguousarray(np.concatenate([a[0, ::max(1, a.shape[1] // 64)], a[-1, ::max(1, a.shape[1] // 64)],
                                                a[::max(1, a.shape[0] // 64), 0], a[::max(1, a.shape[0] // 64), -1]])) \
        if a.ndim == 2 and a.size else sample
    return hash((sample.tobytes(), edge.tobytes()))


def _grid_ctx(Ys, Xs, device=0):
    key = (id(Ys), id(Xs), np.shape(Ys), getattr(Ys, "ctypes", None) and Ys.ctypes.data,
           getattr(Xs, "ctypes", None) and Xs.ctypes.data, device, _fingerprint(Ys), _fingerprint(Xs))
    if _CTX_CACHE["key"] != key:
        if _CTX_CACHE["ctx"] is not None:
            _CTX_CACHE["ctx"].close()
        _CTX_CACHE["ctx"] = GridContext(Ys, Xs, device=device)
        _CTX_CACHE["key"] = key
    return _CTX_CACHE["ctx"]


def release_cached_context():
    if _CTX_CACHE["ctx"] is not None:
        _CTX_CACHE["ctx"].close()
    _CTX_CACHE["ctx"] = None
    _CTX_CACHE["key"] = None


def Heading(plata, plona, platb, plonb):
    """True heading from point a to point b, degrees clockwise from north
    (gonzag/bilin_mapping.py:16-47)."""
    a = [np.array([float(v)]) for v in (plata, plona, platb, plonb)]
    out = np.empty(1)
    L.check(L.lib().gzb_heading(0, 1, *[L.ptr(v) for v in a], L.ptr(out)))
    return float(out[0])


def AlfaBeta(vy, vx):
    """Local coordinates (alpha, beta) of point 0 in the quad 1-2-3-4 (gonzag/bilin_mapping.py:50-141)."""
    y = np.ascontiguousarray(vy, dtype=np.float64).reshape(5)
    x = np.ascontiguousarray(vx, dtype=np.float64).reshape(5)
    ab = np.empty(2)
    L.check(L.lib().gzb_alfabeta(0, 1, L.ptr(y), L.ptr(x), L.ptr(ab)))
    return float(ab[0]), float(ab[1])


def NearestPoint(pcoor_trg, Ys, Xs, rd_found_km=10., resolkm=[], j_prv=None, i_prv=None, np_box_r=10, max_itr=5):
    """[j, i] of the nearest source point, [-1,-1] if none within the acceptance radius
    (gonzag/bilin_mapping.py:144-191).  The unused ``resolkm`` branch and ``max_itr != 5`` are not
    supported."""
    if np.shape(resolkm) == np.shape(Ys):
        raise NotImplementedError("resolkm (2-D adaptive radius) is not supported")
    if max_itr != 5:
        raise NotImplementedError("only max_itr=5 (the reference's only use) is supported")
    (yT, xT) = pcoor_trg
    ctx = _grid_ctx(Ys, Xs)
    ctx.set_option("edge_exclusion", 0)
    try:
        seed = (0 if j_prv is None else int(j_prv), 0 if i_prv is None else int(i_prv))
        NP = ctx.nearest(np.array([yT], dtype=np.float64), np.array([xT], dtype=np.float64),
                         rd_found_km, np_box_r, seed=seed)
    finally:
        ctx.set_option("edge_exclusion", 1)
    return [int(NP[0, 0]), int(NP[0, 1])]


def surrounding_j_i(jP, iP, Ny, Nx, k_ew_per=-1):
    """(jP-1, jP+1, iP-1, iP+1) with the East-West wrap, -1 where the index does not exist
    (gonzag/bilin_mapping.py:194-212).  Integer bookkeeping, done on the host like in the reference; the
    kernels carry their own copy (gzb_mesh.cu)."""
    jm, jp, im, ip = jP - 1, jP + 1, iP - 1, iP + 1
    if jp == Ny:
        jp = -1
    if im == -1 and k_ew_per >= 0:
        im = Nx - k_ew_per - 1
    if ip == Nx:
        ip = k_ew_per if k_ew_per >= 0 else -1
    return jm, jp, im, ip


def _neighbours(jP, iP, Ny, Nx, k_ew_per):
    return surrounding_j_i(jP, iP, Ny, Nx, k_ew_per=k_ew_per)


_QUADRANT_CORNERS = {   # iquadran -> (P2, P3, P4) as offsets picked from (jm, jp, im, ip); gonzag/bilin_mapping.py:240-262
    1: (("j", "ip"), ("jp", "ip"), ("jp", "i")),
    2: (("jm", "i"), ("jm", "ip"), ("j", "ip")),
    3: (("j", "im"), ("jm", "im"), ("jm", "i")),
    4: (("jp", "i"), ("jp", "im"), ("j", "im")),
}


def Iquadran2SrcMesh(jP, iP, Ny, Nx, iqd, k_ew_per=1):
    """The (j,i) pairs of the 3 points that close the source mesh around the nearest point for quadrant `iqd`
    (gonzag/bilin_mapping.py:215-262).  Exits the reference's way on a bad `iqd` or an incomplete
    neighbourhood."""
    from .utils import MsgExit
    if iqd not in (1, 2, 3, 4):
        MsgExit('[Iquadran2SrcMesh()] non-valid "iquadran" value')
    jm, jp, im, ip = surrounding_j_i(jP, iP, Ny, Nx, k_ew_per=k_ew_per)
    if -1 in (jm, jp, im, ip):
        MsgExit('you should not see this! [ Iquadran2SrcMesh() ]')
    v = {"j": jP, "i": iP, "jm": jm, "jp": jp, "im": im, "ip": ip}
    return tuple((v[a], v[b]) for a, b in _QUADRANT_CORNERS[iqd])


def IQH(yA, xA, yB, xB, vY, vX):
    """Through which side of the quadrangle (vY, vX: bottom-left, bottom-right, upper-right, upper-left) the
    segment from A (inside) to B (outside) leaves: (1,'down'), (2,'right'), (3,'up'), (4,'left')
    (gonzag/bilin_mapping.py:267-312).  The five headings are one batched `gzb_heading` call; like the
    reference, four independent tests on un-wrapped headings with the last true one winning, and an exit when
    none is (as written there only 'up' can hold for a quad whose corners lie in their usual directions: the
    other three tests would need a heading interval across 0/360; kept as is, the golden vectors show it)."""
    lonA = float(xA) % 360.
    la = np.full(5, float(yA))
    lo = np.full(5, lonA)
    lb = np.array([float(yB)] + [float(vY[k]) for k in range(4)])
    ob = np.array([float(xB) % 360.] + [float(vX[k]) % 360. for k in range(4)])
    h = np.empty(5)
    L.check(L.lib().gzb_heading(0, 5, L.ptr(la), L.ptr(lo), L.ptr(lb), L.ptr(ob), L.ptr(h)))
    ht, hBL, hBR, hUR, hUL = (float(v) for v in h)
    idir, cdir = -1, 'nowhere'
    if ht > hBL and ht < hBR:
        idir, cdir = 1, 'down'
    if ht > hBR and ht < hUR:
        idir, cdir = 2, 'right'
    if ht > hUR and ht < hUL:
        idir, cdir = 3, 'up'
    if ht > hUL and ht < hBL:
        idir, cdir = 4, 'left'
    if idir not in (1, 2, 3, 4):
        from .utils import GonzagExit
        print('ERROR [IQH]: no side of the quadrangle found!')
        raise GonzagExit(0)
    return idir, cdir


def _mesh_one(pcoor_trg, Ys, Xs, jP, iP, k_ew_per, grid_s_angle, lforceHD):
    (yT, xT) = pcoor_trg
    ctx = _grid_ctx(Ys, Xs)
    ctx.set_ew_per(k_ew_per)
    ctx.set_option("force_heavy", 2 if lforceHD else (1 if abs(grid_s_angle) > 45. else 0))
    try:
        SM = ctx.source_mesh(np.array([yT], dtype=np.float64), np.array([xT], dtype=np.float64),
                             np.array([[jP, iP]], dtype=np.int64))
    finally:
        ctx.set_option("force_heavy", 0)
    return SM[0]


def Iquadran(pcoor_trg, Ys, Xs, jP, iP, k_ew_per=1, grid_s_angle=0., lforceHD=False):
    """Which of the 4 cells around the nearest point holds the target: 1..4, -1 if undecided, None
    when a neighbour is missing (gonzag/bilin_mapping.py:329-449)."""
    (Ny, Nx) = np.shape(Ys)
    jm, jp, im, ip = _neighbours(int(jP), int(iP), Ny, Nx, k_ew_per)
    if -1 in (jm, jp, im, ip) or im < 0 or jm < 0 or ip > Nx - 1:
        return None
    sm = _mesh_one(pcoor_trg, Ys, Xs, int(jP), int(iP), k_ew_per, grid_s_angle, lforceHD)
    if sm[0, 0] < 0:
        return -1
    p2 = (int(sm[1, 0]), int(sm[1, 1]))
    return {(int(jP), ip): 1, (jm, int(iP)): 2, (int(jP), im): 3, (jp, int(iP)): 4}[p2]


def IDSourceMesh(pcoor_trg, Ys, Xs, jP, iP, iquadran=-1, k_ew_per=-1, grid_s_angle=0., lforceHD=False):
    """(4,2) int64 source mesh [P1(nearest), P2, P3, P4], all -1 if not found
    (gonzag/bilin_mapping.py:453-486)."""
    (Ny, Nx) = np.shape(Ys)
    jP, iP = int(jP), int(iP)
    if iquadran > 0:
        if iquadran not in (1, 2, 3, 4):
            raise ValueError('[Iquadran2SrcMesh()] non-valid "iquadran" value')
        jm, jp, im, ip = _neighbours(jP, iP, Ny, Nx, k_ew_per)
        if -1 in (jm, jp, im, ip):
            raise ValueError("the nearest point has no complete neighbourhood")
        table = {1: [(jP, ip), (jp, ip), (jp, iP)], 2: [(jm, iP), (jm, ip), (jP, ip)],
                 3: [(jP, im), (jm, im), (jm, iP)], 4: [(jp, iP), (jp, im), (jP, im)]}
        return np.array([(jP, iP)] + table[iquadran], dtype=np.int64)
    jm, jp, im, ip = _neighbours(jP, iP, Ny, Nx, k_ew_per)
    if -1 in (jm, jp, im, ip) or im < 0 or jm < 0 or ip > Nx - 1:
        return np.full((4, 2), -1, dtype=np.int64)
    return _mesh_one(pcoor_trg, Ys, Xs, jP, iP, k_ew_per, grid_s_angle, lforceHD)


def WeightBL(pcoor_trg, Ys, Xs, isrc_msh):
    """The 4 bilinear weights of a target inside its source mesh (gonzag/bilin_mapping.py:490-527)."""
    (yT, xT) = pcoor_trg
    ctx = _grid_ctx(Ys, Xs)
    sm = np.ascontiguousarray(isrc_msh, dtype=np.int64).reshape(1, 4, 2)
    WB = ctx.weights(np.array([yT], dtype=np.float64), np.array([xT], dtype=np.float64), sm)
    return WB[0]


class BilinTrack:
    def __init__(self, Yt, Xt, Ys, Xs, src_grid_local_angle=[], k_ew_per=-1, rd_found_km=100.,
                 np_box_r=4, freq_talk=0, ctx=None, device=0, seed=None):
        (self.Nt,) = np.shape(Yt)
        self.Yt = Yt
        self.Xt = Xt
        self.Ys = Ys
        self.Xs = Xs
        self.sangle = src_grid_local_angle
        self.kewp = k_ew_per
        self.rfound = rd_found_km
        self.nprad = np_box_r
        (self.Nj, self.Ni) = np.shape(Ys)
        self.ftalk = 2 * self.Nt
        if freq_talk > 0:
            self.ftalk = int(freq_talk)
        self._seed = seed
        self._own_ctx = ctx is None
        if ctx is None:
            ang = src_grid_local_angle if np.shape(src_grid_local_angle) == np.shape(Ys) else None
            ctx = GridContext(Ys, Xs, angle=ang, k_ew_per=k_ew_per, device=device)
        else:
            if (ctx.Nj, ctx.Ni) != (self.Nj, self.Ni):
                raise ValueError("ctx was built for a grid of another shape")
            ctx.set_ew_per(k_ew_per)
            if np.shape(src_grid_local_angle) == np.shape(Ys):
                if not ctx.has_angle:
                    ctx.set_angle(src_grid_local_angle)
            else:
                ctx.set_angle(None)
        self.ctx = ctx
        if config.ivrb > 0:
            print('\n *** [gonzag_b200] mapping %d track points on a %dx%d grid (rd_found_km, np_box_r = %s, %s)'
                  % (self.Nt, self.Nj, self.Ni, rd_found_km, np_box_r))
        # K1 + K2 + K3 in one pass over the device-resident track
        self.NP, self.SM, self.WB = ctx.mapping(Yt, Xt, rd_found_km, np_box_r, seed=seed)

    # the reference exposes the three stages as methods returning fresh arrays
    def nrpt(self):
        return self.ctx.nearest(self.Yt, self.Xt, self.rfound, self.nprad, seed=self._seed)

    def srcm(self):
        return self.ctx.source_mesh(self.Yt, self.Xt, self.NP)

    def wght(self):
        return self.ctx.weights(self.Yt, self.Xt, self.SM)

    def close(self):
        if self._own_ctx and self.ctx is not None:
            self.ctx.close()
        self.ctx = None
