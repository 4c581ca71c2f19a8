"""Output stage of the drop-in: `SaveTimeSeries` and `Save2Dfield` with the reference's call
surface (gonzag/ncio.py:277-341), written as NetCDF-3 (64-bit offset) through `scipy.io`, because
the netCDF4 library the reference uses is not a requirement of this package.  Same dimension,
variable and attribute names; float32 data variables, float64 time; no compression (a NetCDF-4
feature).  Files are readable by any NetCDF reader, including the reference's own.
"""
from __future__ import annotations

import numpy as np

cabout_nc = 'Created with Gonzag (https://github.com/brodeau/gonzag) -- gonzag_b200 output stage'


def _open(ncfile):
    from scipy.io import netcdf_file
    return netcdf_file(ncfile, 'w', version=2)


def Save2Dfield(ncfile, XFLD, xlon=[], xlat=[], name='field', unit='', long_name='', mask=[], clon='nav_lon',
                clat='nav_lat'):
    """2-D field (+ optional coordinates) -> ``ncfile``; points where ``mask < 0.5`` are written as
    NaN (gonzag/ncio.py:277-303)."""
    XFLD = np.ma.getdata(XFLD) if not np.ma.isMaskedArray(XFLD) else np.ma.filled(XFLD.astype(np.float64), np.nan)
    (nj, ni) = np.shape(XFLD)
    f_o = _open(ncfile)
    f_o.createDimension('y', nj)
    f_o.createDimension('x', ni)
    if np.shape(xlon) == (nj, ni) and np.shape(xlat) == (nj, ni):
        id_lon = f_o.createVariable(clon, 'f4', ('y', 'x',))
        id_lat = f_o.createVariable(clat, 'f4', ('y', 'x',))
        id_lon[:, :] = np.asarray(xlon, dtype=np.float32)
        id_lat[:, :] = np.asarray(xlat, dtype=np.float32)
    id_fld = f_o.createVariable(name, 'f4', ('y', 'x',))
    if long_name != '':
        id_fld.long_name = long_name
    if unit != '':
        id_fld.units = unit
    if np.shape(mask) != (0,):
        xtmp = np.array(XFLD, dtype=np.float64, copy=True)
        xtmp[np.where(np.asarray(mask) < 0.5)] = np.nan
        id_fld[:, :] = xtmp.astype(np.float32)
    else:
        id_fld[:, :] = np.asarray(XFLD, dtype=np.float32)
    f_o.about = cabout_nc
    f_o.close()
    return


def SaveTimeSeries(ivt, xd, vvar, ncfile, time_units='unknown', vunits=[], vlnm=[], missing_val=-9999.):
    """Nf time series of length Nt -> ``ncfile`` with an unlimited ``time`` dimension
    (gonzag/ncio.py:306-341).  ``xd``: (Nf, Nt); masked entries are written as ``missing_val``."""
    xd = np.ma.filled(xd, missing_val) if np.ma.isMaskedArray(xd) else np.asarray(xd)
    (Nf, Nt) = xd.shape
    if len(ivt) != Nt:
        raise ValueError('SaveTimeSeries() => disagreement in the number of records between "ivt" and "xd"')
    if len(vvar) != Nf:
        raise ValueError('SaveTimeSeries() => disagreement in the number of fields between "vvar" and "xd"')
    l_f_units = (np.shape(vunits) == (Nf,))
    l_f_lnm = (np.shape(vlnm) == (Nf,))
    f_o = _open(ncfile)
    f_o.createDimension('time', None)
    id_t = f_o.createVariable('time', 'f8', ('time',))
    id_t.calendar = 'gregorian'
    id_t.units = time_units
    id_d = []
    for jf in range(Nf):
        v = f_o.createVariable(vvar[jf], 'f4', ('time',))
        v._FillValue = np.float32(missing_val)
        if l_f_units:
            v.units = vunits[jf]
        if l_f_lnm:
            v.long_name = vlnm[jf]
        id_d.append(v)
    id_t[:] = np.asarray(ivt).astype(np.float64)
    for jf in range(Nf):
        id_d[jf][:] = xd[jf, :].astype(np.float32)
    f_o.about = cabout_nc
    f_o.close()
    return 0
