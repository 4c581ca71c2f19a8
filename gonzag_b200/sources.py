"""In-memory data sources with the call surface of the reference's file readers.

The reference reaches its data through ``gonzag/ncio.py`` (netCDF4) or ``gonzag/zarrio.py``
(``GetTimeEpochVector, GetModelCoor, GetModelLSM, GetModel2DVar, GetSatCoor, GetSatSSH``,
gonzag/ncio.py:98-275).  netCDF4 / xarray / zarr are not part of this package's requirements, so
`ModGrid`, `SatTrack` and `Model2SatTrack` accept, wherever the reference takes a file name, one of
the objects below: same functions, same argument meaning, same slicing quirks (``kt1, kt2`` are
ignored unless BOTH are > 0, gonzag/ncio.py:119-125), backed by arrays or by a per-record callable
(a streaming provider: one 2-D record at a time, as a file reader would deliver them).

A real file name works too: NetCDF-3 files (classic / 64-bit offset) are read by the built-in
`gonzag_b200.nc3io.NetCDF3Source` (scipy, no external library); anything else (NetCDF-4, zarr) goes
to ``gonzag.ncio`` / ``gonzag.zarrio`` when the reference package is importable.
"""
from __future__ import annotations

import numpy as np

from .config import rmissval


def _slice(v, kt1, kt2, isubsamp=1):
    """The reference's reading rule (gonzag/ncio.py:119-125, 236-242): a slice only when both
    indices are > 0, the whole vector otherwise."""
    if kt1 > 0 and kt2 > 0:
        if kt1 >= kt2:
            raise ValueError("mind the indices: kt1 >= kt2")
        return v[kt1:kt2 + 1:isubsamp]
    return v[::isubsamp]


class ModelArrays:
    """A model (source) data set held in memory.

    time    (nrec,) UNIX epoch seconds
    lat,lon 1-D (ny,), (nx,) or 2-D (ny,nx) coordinates, degrees
    fields  {name: (nrec,ny,nx) array | callable(k) -> (ny,nx) array}; masked arrays and a
            ``fill_value`` are honoured when the land-sea mask is derived from ``_FillValue``
    masks   {name: (ny,nx) array of {0,1}} (the `-l <file> -k <name>` case)
    """

    def __init__(self, time, lat, lon, fields=None, masks=None, fill_value=None, name="memory:model"):
        self.time = np.asarray(time, dtype=np.float64)
        self.lat = lat
        self.lon = lon
        self.fields = dict(fields or {})
        self.masks = dict(masks or {})
        self.fill_value = fill_value
        self.name = name

    def __str__(self):
        return self.name

    # ---- gonzag/ncio.py:98-136
    def GetTimeEpochVector(self, kt1=0, kt2=0, isubsamp=1, lquiet=False):
        return np.array(_slice(self.time, kt1, kt2, isubsamp), dtype=np.float64)

    # ---- gonzag/ncio.py:68-94
    def GetTimeInfo(self):
        return len(self.time), (float(self.time[0]), float(self.time[-1]))

    # ---- gonzag/ncio.py:139-163
    def GetModelCoor(self, what):
        if what == "latitude":
            return self.lat
        if what == "longitude":
            return self.lon
        raise ValueError('"what" only supports "latitude" and "longitude"')

    def record(self, ncvar, kt):
        f = self.fields[ncvar]
        return f(int(kt)) if callable(f) else f[int(kt)]

    # ---- gonzag/ncio.py:194-210
    def GetModel2DVar(self, ncvar, kt=0):
        return self.record(ncvar, kt)

    def field_array(self, ncvar):
        """The whole (nrec,ny,nx) array when the field is held as one, else None."""
        f = self.fields.get(ncvar)
        return None if (f is None or callable(f)) else f

    # ---- gonzag/ncio.py:166-191
    def GetModelLSM(self, what):
        if what[:10] == "_FillValue":
            x = self.record(what[11:], 0)
            if np.ma.isMaskedArray(x):
                m = np.ma.getmaskarray(x)
            elif self.fill_value is not None:
                m = (np.asarray(x) == self.fill_value) | np.isnan(x)
            else:
                m = np.isnan(np.asarray(x, dtype=np.float64))
            return (1 - m).astype(int)
        return np.asarray(np.ma.getdata(self.masks[what])).astype(int)


class TrackArrays:
    """A satellite along-track (target) data set held in memory: 1-D ``time, lat, lon`` and
    named 1-D variables (``{"ssh": array}``)."""

    def __init__(self, time, lat, lon, variables=None, name="memory:track"):
        self.time = np.asarray(time, dtype=np.float64)
        self.lat = np.asarray(lat)
        self.lon = np.asarray(lon)
        self.variables = dict(variables or {})
        self.name = name

    def __str__(self):
        return self.name

    def GetTimeEpochVector(self, kt1=0, kt2=0, isubsamp=1, lquiet=False):
        return np.array(_slice(self.time, kt1, kt2, isubsamp), dtype=np.float64)

    def GetTimeInfo(self):
        return len(self.time), (float(self.time[0]), float(self.time[-1]))

    # ---- gonzag/ncio.py:213-245
    def GetSatCoor(self, what, kt1=0, kt2=0):
        v = {"latitude": self.lat, "longitude": self.lon}[what]
        return np.array(_slice(v, kt1, kt2))

    # ---- gonzag/ncio.py:248-270
    def GetSatSSH(self, ncvar, kt1=0, kt2=0, ikeep=[]):
        v = _slice(self.variables[ncvar], kt1, kt2)
        v = np.ma.array(v, copy=True) if np.ma.isMaskedArray(v) else np.array(v, dtype=np.float64, copy=True)
        if len(ikeep) > 0:
            v = v[ikeep]
        if np.ma.is_masked(v):
            v[np.where(np.ma.getmask(v))] = rmissval
        return np.ma.getdata(v) if np.ma.isMaskedArray(v) else v


class _FileSource:
    """A file name bound to the reference's own readers (only when `gonzag` is importable)."""

    def __init__(self, path):
        try:
            from gonzag.config import IsZarr
            if not IsZarr:
                from gonzag import ncio as io
            else:
                from gonzag import zarrio as io
        except Exception as exc:
            raise RuntimeError("'%s' is a file name, and reading files needs the reference's I/O layer "
                               "(gonzag.ncio / gonzag.zarrio with netCDF4 / zarr), which is not importable here: %s.  "
                               "Pass a gonzag_b200.ModelArrays / TrackArrays object instead." % (path, exc))
        self.path = path
        self._io = io

    def __str__(self):
        return self.path

    def __getattr__(self, fn):
        f = getattr(self._io, fn)
        return lambda *a, **k: f(self.path, *a, **k)


def as_source(obj):
    """ModelArrays / TrackArrays (or any object with the reader functions) pass through; a file name
    goes to the built-in NetCDF-3 reader when the file is of that kind, else to the reference's I/O."""
    if isinstance(obj, (ModelArrays, TrackArrays)) or hasattr(obj, "GetTimeEpochVector"):
        return obj
    from .nc3io import NetCDF3Source, is_netcdf3
    if is_netcdf3(str(obj)):
        return NetCDF3Source(str(obj))
    return _FileSource(str(obj))
