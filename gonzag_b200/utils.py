"""Drop-in for the part of the reference's gonzag/utils.py that parametrises the hot path:
the grid heuristics (`GridResolution`, `IsEastWestPeriodic`, `IsGlobalLongitudeWise`), the time
window (`scan_idx`, `GetEpochTimeOverlap`) and the two classes `alongtrack_sat_vs_nemo.py` builds
before calling `Model2SatTrack`: `ModGrid` (gonzag/utils.py:369-466) and `SatTrack` (:477-549).

What runs where:
  * `GridAngle` is kernel K0; `ModGrid` keeps the grid resident on the GPU (``MG.ctx``, a
    `GridContext` that `Model2SatTrack` reuses instead of uploading the grid again) and takes the
    whole-grid reductions of `IsGlobalLongitudeWise` and the latitude range from the device;
  * everything that looks at one row of the grid or at the track's time vector is host NumPy,
    vectorised where the reference loops in Python (`scan_idx`, the 1-D -> 2-D broadcast).

Where the reference takes a file name these classes also take an in-memory source
(`gonzag_b200.sources.ModelArrays` / `TrackArrays`).
"""
from __future__ import annotations

import ctypes as C
from math import copysign

import numpy as np

from . import _lib as L
from . import config
from .config import deg2km
from .context import GridContext
from .sources import as_source


class GonzagExit(SystemExit):
    """What the reference's `MsgExit` amounts to (gonzag/utils.py:15-17: print, then ``exit(0)``)."""


def MsgExit(cmsg):
    print('\n ERROR: ' + cmsg + ' !\n')
    raise GonzagExit(0)


def chck4f(ncfile):
    """Exit the reference's way when a file is missing (gonzag/utils.py:19-22)."""
    from os.path import exists
    if not exists(ncfile):
        MsgExit('File ' + ncfile + ' does not exist')


def find_j_i_min(x):
    """(j, i) of the first minimum of a 2-D array in row-major order (gonzag/utils.py:48-55).  Index
    bookkeeping on a caller's array: host NumPy, like the reference (on the mapping path the argmin is
    part of kernel K1 and never comes through here)."""
    k = x.argmin()
    nx = x.shape[1]
    return k // nx, k % nx


def find_j_i_max(x):
    """(j, i) of the first maximum of a 2-D array (gonzag/utils.py:57-64)."""
    k = x.argmax()
    nx = x.shape[1]
    return k // nx, k % nx


def Haversine(plat, plon, xlat, xlon, device=0):
    """Distance in km (R = 6360 km) from the point (plat, plon) to every point of the arrays `xlat`, `xlon`
    (gonzag/utils.py:296-316): the formula K1 evaluates, batched on the GPU (`gzb_haversine`).  Returns
    float64 with the shape of `xlat`."""
    la, lo = L.as_f64(xlat), L.as_f64(xlon)
    if la.shape != lo.shape:
        raise ValueError("xlat and xlon must have the same shape")
    out = np.empty(la.shape, dtype=np.float64)
    L.check(L.lib().gzb_haversine(int(device), la.size, float(plat), float(plon), L.ptr(la), L.ptr(lo), L.ptr(out)))
    return out


def Haversine_sclr(lat1, lon1, lat2, lon2, device=0):
    """Distance in km between two points with the latitude-dependent Earth radius (gonzag/utils.py:281-293),
    the per-step length of the along-track distance (`gzb_haversine_sclr`).  Scalars in, float out; arrays of
    equal shape are accepted too (one distance per element)."""
    a = [L.as_f64(np.atleast_1d(v)) for v in (lat1, lon1, lat2, lon2)]
    if any(v.shape != a[0].shape for v in a):
        raise ValueError("the four arguments must have the same shape")
    out = np.empty(a[0].shape, dtype=np.float64)
    L.check(L.lib().gzb_haversine_sclr(int(device), a[0].size, *[L.ptr(v) for v in a], L.ptr(out)))
    return float(out[0]) if np.shape(lat1) == () else out


def RadiusEarth(lat, device=0):
    """Radius of the Earth in km at latitude `lat` (degrees North), gonzag/utils.py:267-278 (`gzb_radius_earth`)."""
    a = L.as_f64(np.atleast_1d(lat))
    out = np.empty(a.shape, dtype=np.float64)
    L.check(L.lib().gzb_radius_earth(int(device), a.size, L.ptr(a), L.ptr(out)))
    return float(out[0]) if np.shape(lat) == () else out


def SearchBoxSize(res_mod, width_box):
    """Half-width, in grid points, of the first-attempt search box (gonzag/utils.py:97-110)."""
    return int(0.5 * width_box / res_mod)


def degE_to_degWE(X):
    """Longitude from the 0..360 frame to the -180..180 frame (gonzag/utils.py:24-33)."""
    if np.shape(X) == ():
        return copysign(1., 180. - X) * min(X, abs(X - 360.))
    return np.copysign(1., 180. - X) * np.minimum(X, np.abs(X - 360.))


def EpochT2Str(time):
    """UNIX epoch time -> readable UTC date (gonzag/utils.py:36-45)."""
    from datetime import datetime as dtm, timezone
    return dtm.fromtimestamp(int(round(time, 0)), tz=timezone.utc).strftime('%c')


def GridAngle(xlat, xlon, device=0):
    """Local rotation of the grid in degrees, `GridAngle` of gonzag/utils.py:226-262, on the GPU."""
    with GridContext(xlat, xlon, device=device) as ctx:
        return ctx.grid_angle()


# ------------------------------------------------------------------------------- grid heuristics
def GridResolution(X):
    """Mean |d lon| along the middle row, jumps >= 120 degrees ignored (gonzag/utils.py:68-77)."""
    X = np.asarray(X)
    ny = X.shape[0]
    vx = np.abs(X[ny // 2, 1:] - X[ny // 2, :-1])
    return np.mean(vx[np.where(vx < 120.)])


def IsEastWestPeriodic(X, tol=0.):
    """Number of overlapping columns of an East-West periodic grid, -1 if not periodic: the
    longitude one step past the last column of the middle row is looked for among the first five
    columns (gonzag/utils.py:80-94).  ``tol = 0`` (default) is the reference's test, exact float
    equality -- which misses grids whose longitudes went through float32 or whose spacing is not a
    binary fraction (a synthetic 4322-column ORCA12 gives -1).  ``tol > 0`` accepts a
    match within ``tol`` grid steps, distances taken on the circle (0.01 is plenty)."""
    X = np.asarray(X)
    jj = X.shape[0] // 2
    iper = -1
    dx = X[jj, 1] - X[jj, 0]
    lon_last_p1 = (X[jj, -1] + dx) % 360.
    for it in range(min(5, X.shape[1])):
        if tol > 0.:
            d = abs(lon_last_p1 - X[jj, it] % 360.)
            hit = min(d, 360. - d) <= tol * abs(dx)
        else:
            hit = lon_last_p1 == X[jj, it] % 360.
        if hit:
            iper = it
            break
    return iper


def _global_lonwise(nx, xmin, xmax, imin, imax, xminB, xmaxB, resd):
    """Decision part of IsGlobalLongitudeWise (gonzag/utils.py:166-181) from the six reductions."""
    l360 = True
    lglobal = False
    if (xmin < 1.5 * resd and xmax > 360. - 1.5 * resd) and (imax > imin):
        lglobal = True
        xmin = 0.
        xmax = 360.
    elif (xminB % 360. > xmaxB % 360.) and (imax < imin):
        l360 = False
        xmin = xminB
        xmax = xmaxB
    return lglobal, l360, xmin, xmax


def IsGlobalLongitudeWise(X, resd=1.):
    """(is global, work in the [0,360] frame, lon_min, lon_max) of a model longitude array
    (gonzag/utils.py:149-182).  Host NumPy; `ModGrid` takes the same reductions from the GPU."""
    X = np.asarray(X)
    nx = X.shape[1]
    X = np.mod(X, 360.)
    xb = degE_to_degWE(X)
    return _global_lonwise(nx, np.amin(X), np.amax(X), np.argmin(X) % nx, np.argmax(X) % nx,
                           np.amin(xb), np.amax(xb), resd)


# ------------------------------------------------------------------------------- time window
def scan_idx(vt, rt1, rt2):
    """Indices at which to start and stop scanning a time vector for the window [rt1, rt2]
    (gonzag/utils.py:205-220).  Vectorised restatement of the reference's two ``for ... break``
    loops, including what they leave behind when nothing matches (the last index of the range)."""
    vt = np.asarray(vt, dtype=np.float64)
    nt = len(vt)
    if nt < 2:
        raise ValueError("scan_idx needs a time vector of at least 2 records")
    c1 = (vt[:-1] <= rt1) & (vt[1:] > rt1)
    kt1 = int(np.argmax(c1)) if c1.any() else nt - 2
    c2 = (vt[kt1:-1] <= rt2) & (vt[kt1 + 1:] > rt2)
    kt2 = kt1 + int(np.argmax(c2)) if c2.any() else nt - 2
    return kt1, kt2 + 1


def GetEpochTimeOverlap(ncfile_sat, ncfile_mod):
    """((t_start, t_end), (Nt_sat, Nt_mod)): common time window of the two data sets
    (gonzag/utils.py:187-203)."""
    nts, (zt1_s, zt2_s) = as_source(ncfile_sat).GetTimeInfo()
    ntm, (zt1_m, zt2_m) = as_source(ncfile_mod).GetTimeInfo()
    if (zt1_m >= zt2_s) or (zt1_s >= zt2_m) or (zt2_m <= zt1_s) or (zt2_s <= zt1_m):
        MsgExit('No time overlap for Model and Track file')
    return (max(zt1_s, zt1_m), min(zt2_s, zt2_m)), (nts, ntm)


# ------------------------------------------------------------------------------- classes
class ModGrid:
    """Model (source) grid: ``size, shape, time, lat, lon, mask, HResDeg, HResKM, IsLonGlobal,
    l360, EWPer, IsDistorded, xangle, domain_bounds, file`` as the reference's `ModGrid`
    (gonzag/utils.py:369-466), plus ``ctx``: the grid resident on GPU ``device``.  ``ew_tolerance``: see
    `IsEastWestPeriodic` (0 = the reference's exact-equality periodicity test)."""

    def __init__(self, ncfile, rtu1, rtu2, nclsm, varlsm, distorded_grid=False, device=0, ew_tolerance=0.):
        src = as_source(ncfile)
        self.file = src if not hasattr(src, "path") else src.path
        self.source = src

        rvt = src.GetTimeEpochVector()
        jt1, jt2 = scan_idx(rvt, rtu1, rtu2)
        self.size = jt2 - jt1 + 1
        self.time = src.GetTimeEpochVector(kt1=jt1, kt2=jt2)
        self.jt1, self.jt2 = jt1, jt2

        zlat = np.ma.getdata(src.GetModelCoor('latitude'))
        zlon = np.mod(np.ma.getdata(src.GetModelCoor('longitude')), 360.)
        if len(zlat.shape) == 1 and len(zlon.shape) == 1:
            if config.ivrb > 0:
                print(' *** Model latitude and longitude arrays are 1D => building the 2D version!')
            ny, nx = len(zlat), len(zlon)
            self.lat = np.empty((ny, nx))
            self.lon = np.empty((ny, nx))
            self.lat[:, :] = np.asarray(zlat, dtype=np.float64)[:, None]
            self.lon[:, :] = np.asarray(zlon, dtype=np.float64)[None, :]
        else:
            if zlat.shape != zlon.shape:
                MsgExit('[SatTrack()] => lat and lon disagree in shape')
            self.lat = zlat
            self.lon = zlon
        del zlat, zlon
        self.shape = self.lat.shape

        lsm = as_source(nclsm) if nclsm is not ncfile else src
        self.mask = lsm.GetModelLSM(varlsm)
        if self.mask.shape != self.shape:
            MsgExit('model land-sea mask has a wrong shape => ' + str(self.mask.shape))

        self.HResDeg = GridResolution(self.lon)
        if self.HResDeg > 5. or self.HResDeg < 0.001:
            MsgExit('Model resolution found is surprising, prefer to stop => check "GetModelResolution()" in utils.py')
        self.HResKM = self.HResDeg * deg2km

        # the grid goes to the GPU once; the whole-grid reductions come back from there
        self.ctx = GridContext(self.lat, self.lon, mask=self.mask, k_ew_per=-1, device=device)
        vals = np.empty(4)
        cols = np.empty(2, dtype=np.int64)
        self.ctx._ck(L.lib().gzb_lon_stats(self.ctx._h, L.ptr(vals), L.ptr(cols)))
        self.IsLonGlobal, self.l360, lon_min, lon_max = _global_lonwise(
            self.shape[1], vals[0], vals[1], int(cols[0]), int(cols[1]), vals[2], vals[3], self.HResDeg)
        self.EWPer = IsEastWestPeriodic(self.lon, tol=ew_tolerance) if self.IsLonGlobal else -1    # tol 0 = the reference's exact test
        self.ctx.set_ew_per(self.EWPer)
        la0, la1 = C.c_double(), C.c_double()
        self.ctx._ck(L.lib().gzb_lat_range(self.ctx._h, C.byref(la0), C.byref(la1)))
        lat_min, lat_max = la0.value, la1.value

        self.IsDistorded = distorded_grid
        self._xangle = None
        if distorded_grid:
            if config.ivrb > 0:
                print(' *** Computing angle distortion of the model grid ("-D" option invoked)...')
            self.ctx.grid_angle(out=False)       # K0; stays on the GPU, the host copy is made on demand

        self.domain_bounds = [lat_min, lon_min, lat_max, lon_max]

        if not self.l360:
            lon_we = np.empty(self.shape)
            self.ctx._ck(L.lib().gzb_lon_to_pm180(self.ctx._h, L.ptr(lon_we)))
            self.lon = lon_we

        if config.ivrb > 0:
            print('\n *** About model gridded (source) domain:')
            print('     * shape = ', self.shape)
            print('     * horizontal resolution: ', self.HResDeg, ' degrees or ', self.HResKM, ' km')
            print('     * Is this a global domain w.r.t longitude: ', self.IsLonGlobal)
            if self.IsLonGlobal:
                print('       ==> East West periodicity: ', (self.EWPer >= 0), ', with an overlap of ', self.EWPer, ' points')
            else:
                print('       ==> this is a regional domain, working in the ' + ('[0:360]' if self.l360 else '[-180:180]') + ' frame')
            print('     * lon_min, lon_max = ', round(lon_min, 2), round(lon_max, 2))
            print('     * lat_min, lat_max = ', round(lat_min, 2), round(lat_max, 2))
            print('     * number of time records of interest: ', self.size, ' (' + str(jt1) + ' to ' + str(jt2) + ')\n')

    @property
    def xangle(self):
        """(ny,nx) local grid rotation in degrees (`-D`), zeros without it: the reference's
        attribute, copied from the GPU the first time it is read."""
        if self._xangle is None:
            if self.IsDistorded:
                self._xangle = np.empty(self.shape)
                self.ctx._ck(L.lib().gzb_get_angle(self.ctx._h, L.ptr(self._xangle), L.GZB_HOST))   # a copy: K0 ran once, before the frame switch
            else:
                self._xangle = np.zeros(self.shape)
        return self._xangle

    def close(self):
        if getattr(self, "ctx", None) is not None:
            self.ctx.close()
            self.ctx = None


class SatTrack:
    """Satellite (target) track restricted to a time window and to the model domain:
    ``size, time, lat, lon, jt1, jt2, keepit, file`` as the reference's `SatTrack`
    (gonzag/utils.py:477-549)."""

    def __init__(self, ncfile, rtu1, rtu2, Np=0, domain_bounds=[-90., 0., 90., 360.], l_0_360=True):
        src = as_source(ncfile)
        self.file = src if not hasattr(src, "path") else src.path
        self.source = src

        if Np < 2500:
            rvt = src.GetTimeEpochVector(lquiet=True)
            jt1, jt2 = scan_idx(rvt, rtu1, rtu2)
        else:
            # two passes, as the reference: a 1-in-500 subsample narrows the range first (:506-514)
            kss = 500
            rvt = src.GetTimeEpochVector(isubsamp=kss, lquiet=True)
            j1, j2 = scan_idx(rvt, rtu1, rtu2)
            j1 = j1 * kss
            j2 = j2 * kss
            rvt = src.GetTimeEpochVector(kt1=j1, kt2=j2, lquiet=True)
            jt1, jt2 = scan_idx(rvt, rtu1, rtu2)
            jt1 = jt1 + j1
            jt2 = jt2 + j1
        self.jt1 = jt1
        self.jt2 = jt2

        vtime = src.GetTimeEpochVector(kt1=jt1, kt2=jt2)
        vlat = np.ma.getdata(src.GetSatCoor('latitude', jt1, jt2))
        vlon = np.ma.getdata(src.GetSatCoor('longitude', jt1, jt2))
        # same longitude convention as the model grid (gonzag/utils.py:526-529)
        vlon = np.mod(vlon, 360.) if l_0_360 else degE_to_degWE(vlon)

        [ymin, xmin, ymax, xmax] = domain_bounds
        keepit = np.where((vlat[:] >= ymin) & (vlat[:] <= ymax) & (vlon[:] >= xmin) & (vlon[:] <= xmax))
        self.time = vtime[keepit]
        self.lat = vlat[keepit]
        self.lon = vlon[keepit]
        self.size = len(self.time)
        self.keepit = keepit

        if config.ivrb > 0:
            print('\n *** About satellite track (target) domain:')
            print('     * number of time records of interest for the interpolation to come: ', self.size)
            print('       ==> time record indices: ' + str(jt1) + ' to ' + str(jt2) + ', included\n')
