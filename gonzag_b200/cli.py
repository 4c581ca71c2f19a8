"""``python -m gonzag_b200.cli -s <sat file> -n <sat var> -m <model file> -v <model var> [-l <mask file>|0]
[-k <mask var>] [-D] [-o <dir>] [--device N]``

The sequence of the reference's ``alongtrack_sat_vs_nemo.py`` (same flags, :14-42; same steps, :50-102)
on the drop-in classes: time overlap, `ModGrid` (grid uploaded to the GPU once, `-D` angle computed
there), `SatTrack`, `Model2SatTrack` (K1-K4), then ``result.nc`` and ``xnp_msk.nc``.  File names are
read by the built-in NetCDF-3 readers, or by the reference's I/O layer for NetCDF-4 / zarr when that
package is importable (`gonzag_b200.sources.as_source`).  Outputs are NetCDF-3 (`gonzag_b200.ncout`).
"""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import config
from .mod2sat import Model2SatTrack
from .ncout import Save2Dfield, SaveTimeSeries
from .utils import EpochT2Str, GetEpochTimeOverlap, ModGrid, SatTrack


def parse(argv=None):
    p = argparse.ArgumentParser(description="Interpolates, in space and time, a 2-D field of an ocean model on a "
                                            "non-regular structured grid onto a satellite track (B200 GPU).")
    req = p.add_argument_group("required arguments")
    req.add_argument("-s", "--fsat", required=True, help="satellite along-track file")
    req.add_argument("-n", "--vsat", required=True, help="name of the field of interest in the satellite file")
    req.add_argument("-m", "--fmod", required=True, help="model file")
    req.add_argument("-v", "--vmod", required=True, help="name of the field of interest in the model file")
    p.add_argument("-l", "--fmsk", default="0", help='land-sea mask file, or "0" to use the _FillValue of the model field')
    p.add_argument("-k", "--vmsk", default="tmask", help="name of the land-sea mask in the mask file")
    p.add_argument("-D", "--distgrid", action="store_true", help="account for strong local distortion of the model grid")
    p.add_argument("-o", "--outdir", default=".", help="where result.nc and xnp_msk.nc go (default: current directory)")
    p.add_argument("--device", type=int, default=0, help="CUDA device")
    p.add_argument("--ew-tolerance", type=float, default=0., help="tolerance (in grid steps) of the East-West periodicity "
                   "test; 0 = the reference's exact float equality")
    return p.parse_args(argv)


def main(argv=None):
    a = parse(argv)
    flsm, vlsm = a.fmsk, a.vmsk
    if flsm == "0":
        flsm, vlsm = a.fmod, "_FillValue"
    (it1, it2), (Nts, Ntm) = GetEpochTimeOverlap(a.fsat, a.fmod)
    print(' *** Time overlap between model and satellite in UNIX epoch time: it1, it2', it1, '--', it2)
    print('   => UTC: "' + EpochT2Str(it1) + '" to "' + EpochT2Str(it2) + '"\n')
    clsm = vlsm + "@" + a.vmod if vlsm == "_FillValue" else vlsm
    MG = ModGrid(a.fmod, it1, it2, flsm, clsm, distorded_grid=a.distgrid, device=a.device, ew_tolerance=a.ew_tolerance)
    try:
        ST = SatTrack(a.fsat, it1, it2, Np=Nts, domain_bounds=MG.domain_bounds, l_0_360=MG.l360)
        RES = Model2SatTrack(MG, a.vmod, ST, a.vsat, device=a.device)
        c1, c2 = "Model SSH interpolated in space (", ") and time on satellite track"
        vvar = ["latitude", "longitude", a.vmod + "_np", a.vmod + "_bl", a.vsat, "distance"]
        vunits = ["deg.N", "deg.E", "m", "m", "m", "km"]
        vlongn = ["Latitude", "Longitude", c1 + "nearest-point" + c2, c1 + "bilinear" + c2, "Input satellite data",
                  "Cumulated distance from first point"]
        os.makedirs(a.outdir, exist_ok=True)
        # numpy.array([...]) drops the masks exactly as the reference's call does (:93-94): the file holds the data
        # under the mask, and -9999. where a series was never set
        SaveTimeSeries(RES.time, np.array([RES.lat, RES.lon, np.ma.getdata(RES.ssh_mod_np), np.ma.getdata(RES.ssh_mod),
                                           np.ma.getdata(RES.ssh_sat), RES.distance]),
                       vvar, os.path.join(a.outdir, "result.nc"), time_units="seconds since 1970-01-01 00:00:00",
                       vunits=vunits, vlnm=vlongn, missing_val=config.rmissval)
        if config.l_save_track_on_model_grid:
            xmsk = 1 - np.ma.getmaskarray(RES.XNPtrack).astype(np.int8)
            Save2Dfield(os.path.join(a.outdir, "xnp_msk.nc"), RES.XNPtrack, xlon=MG.lon, xlat=MG.lat, name="track", mask=xmsk)
    finally:
        MG.close()
    return RES


if __name__ == "__main__":
    main()
