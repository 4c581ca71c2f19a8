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

from . import config
from .context import GridContext


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
