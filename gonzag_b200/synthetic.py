"""Seeded synthetic inputs shaped like the reference's data.

The reference's test archive is not available offline, so parity tests and the
benchmark run on generated inputs:

* ``orca_like_grid``   - curvilinear global grid: Mercator-stretched rows, E-W
  overlap columns that are *exact copies* (as in ORCA files, so distance ties are
  systematic), and a northern cap bent towards a displaced pole so that the local
  grid rotation (`GridAngle`, reference gonzag/utils.py:226) exceeds 45 degrees
  on a few per cent of the cells (exercises ``-D``).
* ``regional_grid``    - regular regional box (Example-2 shape), no periodicity.
* ``land_mask``        - smooth pseudo-continents, ~30 % land, ``mask[0,0]=0``.
* ``ssh_records``      - analytic SSH(t, lat, lon) in fp32 (file-like) or fp64.
* ``orbit_track``      - circular-orbit ground track sampled at 1 Hz
  (SARAL-like: 98.55 deg / 6035 s, Jason-like: 66.04 deg / 6745.7 s), land
  points dropped to create realistic gaps.

Everything here is NumPy, runs on the host and is input preparation only: it is
not part of the hot path.
"""
from __future__ import annotations

import numpy as np

SARAL = dict(incl_deg=98.55, period_s=6035.0)
JASON = dict(incl_deg=66.04, period_s=6745.7)

_OMEGA_EARTH = 2.0 * np.pi / 86164.0905  # rad/s, sidereal


# --------------------------------------------------------------------------- land
def _land_waves(seed: int = 1, nwaves: int = 7):
    rng = np.random.default_rng(seed)
    kvec = rng.normal(size=(nwaves, 3))
    kvec *= (rng.uniform(1.5, 4.5, size=(nwaves, 1)) / np.linalg.norm(kvec, axis=1, keepdims=True))
    phase = rng.uniform(0.0, 2.0 * np.pi, size=nwaves)
    amp = rng.uniform(0.5, 1.0, size=nwaves)
    return kvec, phase, amp


def land_function(lat_deg, lon_deg, seed: int = 1):
    """Smooth scalar on the sphere; land is where it exceeds ``LAND_LEVEL``."""
    kvec, phase, amp = _land_waves(seed)
    la = np.radians(np.asarray(lat_deg, dtype=np.float64))
    lo = np.radians(np.asarray(lon_deg, dtype=np.float64))
    x = np.cos(la) * np.cos(lo)
    y = np.cos(la) * np.sin(lo)
    z = np.sin(la)
    f = np.zeros_like(x)
    for (kx, ky, kz), ph, a in zip(kvec, phase, amp):
        f += a * np.sin(kx * x + ky * y + kz * z + ph)
    return f


LAND_LEVEL = 0.78  # calibrated on seed 1: ~30 % of the sphere is land


def land_mask(lat, lon, seed: int = 1, level: float = LAND_LEVEL):
    """int64 {0,1} mask (1 = ocean) like `GetModelLSM` (reference ncio.py:164-191)."""
    msk = (land_function(lat, lon, seed) <= level).astype(np.int64)
    msk[0, 0] = 0
    return msk


# --------------------------------------------------------------------------- grids
def _stretched_lat(nj: int, lat_s: float, lat_n: float, lat_c: float):
    """Row latitudes: Mercator (isotropic cells) up to ``lat_c``, then constant
    spacing up to ``lat_n`` (so that polar rows do not pile up)."""
    def merc(p):
        return np.log(np.tan(np.pi / 4.0 + np.radians(p) / 2.0))
    span_m = merc(lat_c) - merc(lat_s)                      # in Mercator radians
    span_c = np.radians(lat_n - lat_c) / np.cos(np.radians(lat_c))
    n_m = int(round((nj - 1) * span_m / (span_m + span_c)))
    n_m = min(max(n_m, 2), nj - 3)
    y = np.linspace(merc(lat_s), merc(lat_c), n_m + 1)
    lat_m = np.degrees(2.0 * np.arctan(np.exp(y)) - np.pi / 2.0)
    lat_cap = np.linspace(lat_c, lat_n, nj - n_m)[1:]
    return np.concatenate([lat_m, lat_cap])


def orca_like_grid(nj: int, ni: int, overlap: int = 2, lat_south: float = -78.0,
                   lat_north: float = 88.0, cap_from: float = 30.0,
                   pole_shift_deg: float = 20.0, pole_lon_deg: float = 255.0,
                   born_fp32: bool = True, lon_start: float = 0.0):
    """Global curvilinear grid ``(lat, lon)`` of shape ``(nj, ni)``, fp64, lon in [0,360).

    Columns ``1 .. ni-overlap`` are the interior; with ``overlap=2`` column 0 is a
    copy of column ``ni-2`` and column ``ni-1`` a copy of column 1, i.e. the
    layout `BilinTrack` expects for ``k_ew_per=2`` (reference
    gonzag/bilin_mapping.py:206-210).
    """
    if overlap not in (0, 2):
        raise ValueError("overlap must be 0 or 2")
    nint = ni - overlap
    dx = 360.0 / nint
    lon1 = lon_start + dx * np.arange(nint)
    lat1 = _stretched_lat(nj, lat_south, lat_north, max(cap_from, 45.0))
    lon2, lat2 = np.meshgrid(lon1, lat1)

    # bend the northern cap: rotate each point about an equatorial axis by an
    # angle growing smoothly with latitude, which moves the grid pole away from
    # the geographic pole towards ``pole_lon_deg``.
    s = np.clip((lat2 - cap_from) / (lat_north - cap_from), 0.0, 1.0)
    w = s * s * (3.0 - 2.0 * s)
    ang = np.radians(pole_shift_deg) * w
    la = np.radians(lat2)
    lo = np.radians(lon2 - pole_lon_deg)          # frame where the new pole lies on lon 0
    x = np.cos(la) * np.cos(lo)
    y = np.cos(la) * np.sin(lo)
    z = np.sin(la)
    # rotation about the y axis, tipping +z towards +x
    xr = x * np.cos(ang) + z * np.sin(ang)
    zr = -x * np.sin(ang) + z * np.cos(ang)
    lat_i = np.degrees(np.arcsin(np.clip(zr, -1.0, 1.0)))
    lon_i = np.mod(np.degrees(np.arctan2(y, xr)) + pole_lon_deg, 360.0)

    if born_fp32:  # coordinates as read from a float32 NetCDF variable
        lat_i = lat_i.astype(np.float32).astype(np.float64)
        lon_i = np.mod(lon_i.astype(np.float32).astype(np.float64), 360.0)

    lat = np.empty((nj, ni), dtype=np.float64)
    lon = np.empty((nj, ni), dtype=np.float64)
    if overlap == 2:
        lat[:, 1:ni - 1] = lat_i
        lon[:, 1:ni - 1] = lon_i
        lat[:, 0] = lat[:, ni - 2]
        lon[:, 0] = lon[:, ni - 2]
        lat[:, ni - 1] = lat[:, 1]
        lon[:, ni - 1] = lon[:, 1]
    else:
        lat[:, :] = lat_i
        lon[:, :] = lon_i
    return lat, lon


def regional_grid(nj: int, ni: int, lat0: float = -55.0, lon0: float = 140.0,
                  res_deg: float = 1.0 / 12.0, born_fp32: bool = True, shear: float = 0.0):
    """Regular regional box (Example-2 shape); ``shear`` tilts rows to make it curvilinear."""
    lon1 = lon0 + res_deg * np.arange(ni)
    lat1 = lat0 + res_deg * np.arange(nj)
    lon2, lat2 = np.meshgrid(lon1, lat1)
    if shear != 0.0:
        lat2 = lat2 + shear * (lon2 - lon0)
        lon2 = lon2 + 0.5 * shear * (lat2 - lat0)
    lon2 = np.mod(lon2, 360.0)
    if born_fp32:
        lat2 = lat2.astype(np.float32).astype(np.float64)
        lon2 = np.mod(lon2.astype(np.float32).astype(np.float64), 360.0)
    return np.ascontiguousarray(lat2), np.ascontiguousarray(lon2)


def grid_resolution_deg(lon):
    """Same estimate as the reference's `GridResolution` (gonzag/utils.py:68-76)."""
    nj = lon.shape[0]
    d = np.abs(lon[nj // 2, 1:] - lon[nj // 2, :-1])
    return float(np.mean(d[d < 120.0]))


# --------------------------------------------------------------------------- field
def ssh_record(lat, lon, k: int, dtype=np.float32, seed: int = 2):
    """Analytic SSH record ``k`` (metres, O(1)) on the grid, in ``dtype``."""
    rng = np.random.default_rng(seed + 7919 * (k + 1))
    eps = rng.uniform(-0.05, 0.05)
    la = np.radians(lat)
    lo = np.radians(lon)
    f = (0.6 * np.sin(la) * np.cos(lo + 0.3 * k)
         + 0.25 * np.cos(3.0 * la + 0.11 * k) * np.sin(2.0 * lo)
         + eps)
    return np.ascontiguousarray(f.astype(dtype))


def ssh_records(lat, lon, nrec: int, dtype=np.float32, seed: int = 2):
    out = np.empty((nrec,) + lat.shape, dtype=dtype)
    for k in range(nrec):
        out[k] = ssh_record(lat, lon, k, dtype=dtype, seed=seed)
    return out


# --------------------------------------------------------------------------- tracks
def orbit_track(n_seconds: int, t0: float = 0.0, incl_deg: float = 98.55,
                period_s: float = 6035.0, phase: float = 0.37, lon0_deg: float = 11.3,
                dt: float = 1.0, drop_land_seed: int | None = 1,
                lat_bounds=None, lon_bounds=None, lon_frame_360: bool = True, start_s: float = 0.0):
    """Ground track of a circular orbit: returns ``time, lat, lon`` (fp64).

    ``drop_land_seed`` removes points over the pseudo-continents of
    `land_function` (gaps/jumps like a real altimeter product);
    ``lat_bounds`` / ``lon_bounds`` emulate the bounding-box filter of the
    reference's `SatTrack` (gonzag/utils.py:531-539).
    """
    n = int(round(n_seconds / dt))
    tt = start_s + dt * np.arange(n, dtype=np.float64)      # start_s: a later window of the same orbit
    u = 2.0 * np.pi * tt / period_s + phase
    inc = np.radians(incl_deg)
    lat = np.degrees(np.arcsin(np.sin(inc) * np.sin(u)))
    lon = np.degrees(np.arctan2(np.cos(inc) * np.sin(u), np.cos(u)) - _OMEGA_EARTH * tt) + lon0_deg
    lon = np.mod(lon, 360.0)
    keep = np.ones(n, dtype=bool)
    if drop_land_seed is not None:
        keep &= land_function(lat, lon, drop_land_seed) <= LAND_LEVEL
    if not lon_frame_360:
        lon = np.where(lon > 180.0, lon - 360.0, lon)
    if lat_bounds is not None:
        keep &= (lat >= lat_bounds[0]) & (lat <= lat_bounds[1])
    if lon_bounds is not None:
        keep &= (lon >= lon_bounds[0]) & (lon <= lon_bounds[1])
    return (t0 + tt)[keep], lat[keep], lon[keep]


def multi_mission_track(n_seconds: int, n_missions: int = 6, t0: float = 0.0, seed: int = 3):
    """C5-style batch: several missions with different phases / inclinations, each
    returned as its own ``(time, lat, lon)`` triple."""
    out = []
    for m in range(n_missions):
        rng = np.random.default_rng(seed + m)
        base = SARAL if m % 2 == 0 else JASON
        out.append(orbit_track(n_seconds, t0=t0, phase=float(rng.uniform(0, 2 * np.pi)),
                               lon0_deg=float(rng.uniform(0, 360)), **base))
    return out


# --------------------------------------------------------------------------- configs
def search_params(res_deg: float, rfactor: float = 0.75, deg2km: float = 111.11,
                  box_w_km: float = 500.0):
    """(rd_found_km, np_box_r) exactly as `Model2SatTrack` derives them
    (reference gonzag/mod2sat.py:35,39 and gonzag/utils.py:110)."""
    rd_found_km = rfactor * res_deg * deg2km
    np_box_r = int(0.5 * box_w_km / (res_deg * deg2km))
    return rd_found_km, np_box_r


CONFIGS = {
    # name: (Nj, Ni, kind)
    "C1": dict(nj=292, ni=362, kind="global", track=SARAL, days=1.0, rec_dt=30.0 * 86400.0),
    "C2": dict(nj=800, ni=1000, kind="regional", track=SARAL, days=1.0, rec_dt=3600.0),
    "C3": dict(nj=3059, ni=4322, kind="global", track=JASON, days=10.0, rec_dt=3600.0),
    "C4": dict(nj=9000, ni=13000, kind="global", track=JASON, days=30.0, rec_dt=3600.0),
}
