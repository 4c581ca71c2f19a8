"""The utilities of gonzag/utils.py that sit on the hot path.

`GridAngle` runs on the GPU (kernel K0); `SearchBoxSize` and `degE_to_degWE` are the scalar
parameter helpers `Model2SatTrack` needs (gonzag/utils.py:97-110, 24-33).
"""
from __future__ import annotations

from math import copysign

import numpy as np

from .context import GridContext


def SearchBoxSize(res_mod, width_box):
    """Half-width, in grid points, of the first-attempt search box (gonzag/utils.py:97-110)."""
    return int(0.5 * width_box / res_mod)


def degE_to_degWE(X):
    """Longitude from the 0..360 frame to the -180..180 frame (gonzag/utils.py:24-33)."""
    if np.shape(X) == ():
        return copysign(1., 180. - X) * min(X, abs(X - 360.))
    return np.copysign(1., 180. - X) * np.minimum(X, np.abs(X - 360.))


def GridAngle(xlat, xlon, device=0):
    """Local rotation of the grid in degrees, `GridAngle` of gonzag/utils.py:226-262, on the GPU."""
    with GridContext(xlat, xlon, device=device) as ctx:
        return ctx.grid_angle()
