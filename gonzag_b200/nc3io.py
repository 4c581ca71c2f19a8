"""NetCDF-3 readers with the call surface of the reference's gonzag/ncio.py:68-270
(`GetTimeInfo, GetTimeEpochVector, GetModelCoor, GetModelLSM, GetModel2DVar, GetSatCoor,
GetSatSSH`), through `scipy.io.netcdf_file` (classic and 64-bit-offset files, memory-mapped).

The reference reads with the netCDF4 library, which is not a requirement of this package (and is
absent offline).  NetCDF-3 is the one NetCDF flavour that needs no external library: files of that
kind given BY NAME to `ModGrid`, `SatTrack`, `GetEpochTimeOverlap` and `Model2SatTrack` are read
here; NetCDF-4 / HDF5 and zarr stores still go to the reference's own I/O layer when that is
importable (`sources.as_source`).

What netCDF4 does silently and the reference relies on is restated here:
  * auto-masking: values equal to ``_FillValue`` / ``missing_value`` (else the library's default
    fill value of the type) come back masked -- the `-l 0` land-sea mask is derived from that
    (gonzag/ncio.py:178-182), and masked satellite SSH becomes `rmissval` (:268);
  * auto-scaling: ``scale_factor`` / ``add_offset``;
  * ``num2date`` of the first time value, for the calendars standard / gregorian /
    proleptic_gregorian / noleap / 365_day / all_leap / 366_day / 360_day, which
    `ToEpochTime` (gonzag/ncio.py:20-64) turns into UNIX epoch seconds by reading the calendar date
    as a real one; the rest of the vector is ``t0 + (X - X[0]) * seconds_per_unit``.

A record of the model field is delivered as one native-endian 2-D array per call (NetCDF-3 stores
big-endian): `Model2SatTrack` stages the records of a slab into pinned memory and overlaps the
copies with the gather kernel, exactly as for any per-record provider.
"""
from __future__ import annotations

import calendar as _cal
import datetime as _dt
import re

import numpy as np

from .config import rmissval

_TIME_NAMES = ['time', 'time_counter', 'TIME', 'record', 't']
_MODEL_LAT = ['lat', 'latitude', 'nav_lat', 'gphit', 'LATITUDE']
_MODEL_LON = ['lon', 'longitude', 'nav_lon', 'glamt', 'LONGITUDE']
_SAT_LAT = ['lat', 'latitude', 'LATITUDE']
_SAT_LON = ['lon', 'longitude', 'LONGITUDE']

# netCDF4.default_fillvals
_DEFAULT_FILL = {'f4': 9.969209968386869e+36, 'f8': 9.969209968386869e+36, 'i4': -2147483647, 'i2': -32767,
                 'i1': -127, 'S1': None}


def is_netcdf3(path):
    """True for a regular file that starts with the NetCDF classic / 64-bit-offset magic."""
    try:
        with open(path, 'rb') as f:
            return f.read(4) in (b'CDF\x01', b'CDF\x02')
    except (OSError, TypeError):
        return False


def _attr(var, name, default=None):
    v = getattr(var, name, default)
    if isinstance(v, bytes):
        v = v.decode()
    return v


# ------------------------------------------------------------------------------- time
_UNIT_SECONDS = {'second': 1., 'seconds': 1., 'sec': 1., 'secs': 1., 's': 1.,
                 'minute': 60., 'minutes': 60., 'min': 60., 'mins': 60.,
                 'hour': 3600., 'hours': 3600., 'hr': 3600., 'hrs': 3600., 'h': 3600.,
                 'day': 86400., 'days': 86400., 'd': 86400.}
_MONTH_DAYS = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]


def _parse_units(units):
    m = re.match(r'\s*(\w+)\s+since\s+(-?\d+)-(\d+)-(\d+)(?:[ T](\d+):(\d+)(?::(\d+(?:\.\d*)?))?)?\s*(Z|UTC|[+-]\d{1,2}:?\d{0,2})?\s*$',
                 units)
    if not m:
        raise ValueError("unsupported time units: '%s'" % units)
    unit = m.group(1).lower()
    if unit not in _UNIT_SECONDS:
        raise ValueError("unsupported time unit: '%s'" % units)
    y, mo, d = int(m.group(2)), int(m.group(3)), int(m.group(4))
    hh, mi = int(m.group(5) or 0), int(m.group(6) or 0)
    ss = float(m.group(7) or 0.)
    tz = m.group(8)
    off = 0.
    if tz and tz not in ('Z', 'UTC'):
        sign = -1. if tz[0] == '-' else 1.
        digits = tz[1:].replace(':', '')
        off = sign * (int(digits[:2] if len(digits) > 2 else digits) * 3600. + (int(digits[2:]) * 60. if len(digits) > 2 else 0.))
    return _UNIT_SECONDS[unit], (y, mo, d, hh, mi, ss), off


def _days_before_year(y, cal):
    """Days from 0001-01-01 to the start of year y in a fixed-length-year calendar."""
    n = {'noleap': 365, 'all_leap': 366, '360_day': 360}[cal]
    return (y - 1) * n


def _date_to_days(y, mo, d, cal):
    if cal == '360_day':
        return _days_before_year(y, cal) + (mo - 1) * 30 + (d - 1)
    md = list(_MONTH_DAYS)
    if cal == 'all_leap':
        md[1] = 29
    return _days_before_year(y, cal) + sum(md[:mo - 1]) + (d - 1)


def _days_to_date(n, cal):
    ny = {'noleap': 365, 'all_leap': 366, '360_day': 360}[cal]
    y = n // ny + 1
    r = n - (y - 1) * ny
    if cal == '360_day':
        return int(y), int(r // 30 + 1), int(r % 30 + 1)
    md = list(_MONTH_DAYS)
    if cal == 'all_leap':
        md[1] = 29
    mo = 0
    while r >= md[mo]:
        r -= md[mo]
        mo += 1
    return int(y), mo + 1, int(r + 1)


def _canon_calendar(cal):
    cal = (cal or 'standard').lower()
    if cal in ('standard', 'gregorian', 'proleptic_gregorian'):
        return 'real'
    if cal in ('noleap', '365_day'):
        return 'noleap'
    if cal in ('all_leap', '366_day'):
        return 'all_leap'
    if cal == '360_day':
        return '360_day'
    raise ValueError("unsupported calendar: '%s'" % cal)


def num2epoch(value, units, calendar='standard'):
    """``ToEpochTime`` of a scalar (gonzag/ncio.py:40-46): num2date in the file's calendar, rounded
    to the microsecond, its ``%Y-%m-%d %H:%M:%S`` reading taken as a real UTC date -> epoch seconds
    (float), plus the microseconds."""
    r2s, (y, mo, d, hh, mi, ss), tzoff = _parse_units(units)
    cal = _canon_calendar(calendar)
    # offset in integer microseconds (num2date's resolution)
    us = int(round((float(value) * r2s - tzoff) * 1e6)) + int(round(ss * 1e6)) + (hh * 3600 + mi * 60) * 1000000
    days, us_in_day = divmod(us, 86400 * 1000000)
    sec, micro = divmod(us_in_day, 1000000)
    if cal == 'real':
        date = _dt.date(y, mo, d) + _dt.timedelta(days=days)
        yy, mm, dd = date.year, date.month, date.day
    else:
        yy, mm, dd = _days_to_date(_date_to_days(y, mo, d, cal) + days, cal)
    # the calendar date read as a real one (strftime -> strptime -> timegm in the reference)
    if mm == 2 and dd > (29 if _cal.isleap(yy) else 28):
        raise ValueError("date %04d-%02d-%02d of calendar '%s' does not exist in the real calendar" % (yy, mm, dd, calendar))
    return float(_cal.timegm((yy, mm, dd, 0, 0, 0)) + sec + micro * 1e-6)


def to_epoch_time(X, units, calendar):
    """`ToEpochTime` (gonzag/ncio.py:20-64): scalar or vector of "<unit> since <date>" -> epoch seconds."""
    if np.shape(X) == ():
        return num2epoch(X, units, calendar)
    X = np.ma.getdata(X).astype(np.float64)
    t0 = num2epoch(X[0], units, calendar)
    u = units.strip()
    if u[0:13] == 'seconds since':
        r2s = 1.
    elif u[0:11] == 'hours since':
        r2s = 3600.
    elif u[0:10] == 'days since':
        r2s = 86400.
    else:
        raise ValueError('[ToEpochTime()] => unsupported time unit: ' + units)
    vet = np.zeros(len(X))
    vet[0] = t0
    vet[1:] = t0 + (X[1:] - X[0]) * r2s
    return vet


# ------------------------------------------------------------------------------- the source
def C_void(a):
    """Address of an array that may be read-only (a view of the memory map): `ndarray.ctypes.data` without the
    writeable check of `ctypes.data_as`."""
    import ctypes
    return ctypes.c_void_p(a.__array_interface__['data'][0])


class NetCDF3Source:
    """One NetCDF-3 file bound to the reader functions of the reference (`ncfile` argument
    dropped).  The file stays memory-mapped while the object lives; everything returned is a copy."""

    def __init__(self, path):
        from scipy.io import netcdf_file
        self.path = str(path)
        self._f = netcdf_file(self.path, 'r', mmap=True, maskandscale=False)

    def __str__(self):
        return self.path

    def close(self):
        f, self._f = self._f, None
        if f is not None:
            import warnings
            with warnings.catch_warnings():        # scipy warns that its own variable objects still view the map
                warnings.simplefilter("ignore", RuntimeWarning)
                try:
                    f.close()
                except Exception:
                    pass

    def __del__(self):
        self.close()

    # ---- what netCDF4 does on read
    def _read(self, var, index):
        raw = var[index] if var.shape != () else var.getValue()   # scipy: a view of the map
        x = np.array(raw, dtype=raw.dtype.newbyteorder('='))      # native-endian copy
        fill = None
        for att in ('_FillValue', 'missing_value'):
            v = getattr(var, att, None)
            if v is not None:
                fill = np.asarray(v).ravel()[0]
                break
        if fill is None:
            fill = _DEFAULT_FILL.get(x.dtype.str[1:], None)
        mask = None
        if fill is not None and x.dtype.kind in 'fiu':
            mask = np.isnan(x) if (x.dtype.kind == 'f' and np.isnan(fill)) else (x == np.asarray(fill).astype(x.dtype))
        sf, ao = getattr(var, 'scale_factor', None), getattr(var, 'add_offset', None)
        if sf is not None or ao is not None:
            x = x * (1.0 if sf is None else np.asarray(sf).ravel()[0]) + (0.0 if ao is None else np.asarray(ao).ravel()[0])
        if np.ndim(x) == 0:
            return np.ma.masked if (mask is not None and bool(mask)) else x[()]
        return np.ma.MaskedArray(x, mask=(mask if mask is not None else False), copy=False)

    def _var(self, names, what):
        for n in names:
            if n in self._f.variables:
                return n, self._f.variables[n]
        raise KeyError('could not find %s in %s (tried %s)' % (what, self.path, ', '.join(names)))

    def _time_var(self):
        _, v = self._var(_TIME_NAMES, 'a time-record variable')
        return v, _attr(v, 'units'), _attr(v, 'calendar', 'standard')

    # ---- gonzag/ncio.py:68-94
    def GetTimeInfo(self):
        dims = self._f.dimensions
        cd = next((c for c in _TIME_NAMES if c in dims), None)
        if cd is None:
            raise KeyError('found no time-record dimension in file ' + self.path)
        if cd not in self._f.variables:
            raise KeyError('name of time variable is different than name of time dimension')
        v = self._f.variables[cd]
        nt = int(v.shape[0])             # the record dimension's length is None in scipy: take the variable's
        units, cal = _attr(v, 'units'), _attr(v, 'calendar', 'standard')
        return nt, (to_epoch_time(self._read(v, 0), units, cal), to_epoch_time(self._read(v, nt - 1), units, cal))

    # ---- gonzag/ncio.py:98-136
    def GetTimeEpochVector(self, kt1=0, kt2=0, isubsamp=1, lquiet=False):
        v, units, cal = self._time_var()
        if kt1 > 0 and kt2 > 0:
            if kt1 >= kt2:
                raise ValueError('mind the indices when calling GetTimeEpochVector()')
            vdate = self._read(v, slice(kt1, kt2 + 1, isubsamp))
        else:
            vdate = self._read(v, slice(None, None, isubsamp))
        return to_epoch_time(vdate, units, cal)

    # ---- gonzag/ncio.py:139-163
    def GetModelCoor(self, what):
        if what == 'latitude':
            names = _MODEL_LAT
        elif what == 'longitude':
            names = _MODEL_LON
        else:
            raise ValueError('"what" argument of "GetModelCoor()" only supports "latitude" and "longitude"')
        _, v = self._var(names, what + ' array')
        nd = len(v.dimensions)
        if nd in (1, 2):
            return self._read(v, Ellipsis)
        if nd == 3:
            return self._read(v, 0)
        raise ValueError('Model ' + what + ' has a weird number of dimensions')

    # ---- gonzag/ncio.py:166-191
    def GetModelLSM(self, what):
        l_fill_val = (what[:10] == '_FillValue')
        ncvar = what[11:] if l_fill_val else what
        v = self._f.variables[ncvar]
        nd = len(v.dimensions)
        if l_fill_val:
            if nd not in (3, 4):
                raise ValueError(ncvar + ' is expected to have 3 or 4 dimensions')
            x = self._read(v, 0 if nd == 3 else (0, 0))
            return (1 - np.ma.getmaskarray(x)).astype(int)
        if nd not in (2, 3, 4):
            raise ValueError('Mask ' + ncvar + ' has a weird number of dimensions: ' + str(nd))
        x = self._read(v, {2: Ellipsis, 3: 0, 4: (0, 0)}[nd])
        return np.ma.getdata(x).astype(int)

    # ---- gonzag/ncio.py:194-210
    def GetModel2DVar(self, ncvar, kt=0):
        v = self._f.variables[ncvar]
        nd = len(v.dimensions)
        if nd == 3:
            return self._read(v, int(kt))
        if nd == 4:
            return self._read(v, (int(kt), 0))      # surface field
        raise ValueError('Model "' + ncvar + '" has a weird number of dimensions: ' + str(nd))

    def field_array(self, ncvar):
        """No whole-array view: NetCDF-3 data are big-endian and may carry fill values, so the
        records go through `GetModel2DVar` one at a time (streaming provider)."""
        return None

    def record_into(self, ncvar, kt, out):
        """The data of ``GetModel2DVar(ncvar, kt)`` (fill values included, as `numpy.ma.getdata` of it gives them)
        written straight into `out` -- a C-contiguous native-endian array, typically a slot of the pinned staging
        slab -- in one threaded byte-swapping pass over the memory map (`gzb_copy_swap`) instead of NumPy's three
        single-threaded ones (swap, fill-value mask, copy).  Returns False, leaving `out` untouched, when the fast
        path does not apply (scaled / offset variable, other dtype or shape): the caller then reads the record the
        ordinary way."""
        from . import _lib as L
        v = self._f.variables[ncvar]
        nd = len(v.dimensions)
        if nd not in (3, 4):
            raise ValueError('Model "' + ncvar + '" has a weird number of dimensions: ' + str(nd))
        if getattr(v, 'scale_factor', None) is not None or getattr(v, 'add_offset', None) is not None:
            return False
        raw = v[int(kt)] if nd == 3 else v[int(kt), 0]
        if (not isinstance(out, np.ndarray) or not out.flags['C_CONTIGUOUS'] or not out.flags['WRITEABLE']
                or not raw.flags['C_CONTIGUOUS'] or out.shape != raw.shape or not out.dtype.isnative
                or out.dtype != raw.dtype.newbyteorder('=') or raw.dtype.itemsize not in (1, 2, 4, 8)):
            return False
        L.check(L.lib().gzb_copy_swap(L.ptr(out), C_void(raw), raw.dtype.itemsize, raw.size, 0 if raw.dtype.isnative else 1))
        return True

    # ---- gonzag/ncio.py:213-245
    def GetSatCoor(self, what, kt1=0, kt2=0):
        if what == 'latitude':
            names = _SAT_LAT
        elif what == 'longitude':
            names = _SAT_LON
        else:
            raise ValueError('"what" argument of "GetSatCoor()" only supports "latitude" and "longitude"')
        _, v = self._var(names, what + ' array')
        if len(v.dimensions) != 1:
            raise ValueError('Satellite ' + what + ' has a weird number of dimensions (we expect only 1: the time-record!)')
        if kt1 > 0 and kt2 > 0:
            if kt1 >= kt2:
                raise ValueError('mind the indices when calling GetSatCoor()')
            return self._read(v, slice(kt1, kt2 + 1))
        return self._read(v, Ellipsis)

    # ---- gonzag/ncio.py:248-270
    def GetSatSSH(self, ncvar, kt1=0, kt2=0, ikeep=[]):
        v = self._f.variables[ncvar]
        if kt1 > 0 and kt2 > 0:
            if kt1 >= kt2:
                raise ValueError('mind the indices when calling GetSatSSH()')
            vssh = self._read(v, slice(kt1, kt2 + 1))
        else:
            vssh = self._read(v, Ellipsis)
        if len(ikeep) > 0:
            vssh = vssh[ikeep]
        vssh = np.ma.MaskedArray(vssh, dtype=np.float64)
        if np.ma.is_masked(vssh):
            vssh[np.where(np.ma.getmask(vssh))] = rmissval
        return np.ma.getdata(vssh)
