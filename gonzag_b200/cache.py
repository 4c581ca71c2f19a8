"""f4: persisting a mapping (NP, SM, WB) so that a later run on the same grid + track pair -- another
model variable, another experiment on the same mesh -- skips K1-K3.  The reference has no such cache; this is an
extension around the drop-in classes:

    R = gonzag_b200.Model2SatTrack(MG, "ssh", ST, "ssh", mapping=cache.load_mapping(path, MG, ST))      # None -> computed
    cache.save_mapping(path, R.BT, MG, ST)

The key is a FINGERPRINT, not a hash of every byte (hashing an ORCA12 grid, 211 MB, costs more than K1-K3 on a resident
grid): shapes, the search parameters derived from `MG.HResDeg`, `MG.EWPer`, whether `-D` is on, and a strided sample of
~4096 values (+ the four edges) of the grid coordinates, the angle and the track.  A file whose fingerprint differs is
ignored, never trusted.
"""
from __future__ import annotations

import hashlib
import os

import numpy as np

from . import config
from .mod2sat import MappingView

_VERSION = 1


def _sample(a, n=4096):
    a = np.asarray(a)
    if a.size == 0:
        return b""
    flat = a.reshape(-1)
    step = max(1, flat.size // n)
    parts = [np.ascontiguousarray(flat[::step]).tobytes(), np.ascontiguousarray(flat[-1:]).tobytes()]
    if a.ndim == 2:
        sj, si = max(1, a.shape[0] // 64), max(1, a.shape[1] // 64)
        parts += [np.ascontiguousarray(a[0, ::si]).tobytes(), np.ascontiguousarray(a[-1, ::si]).tobytes(),
                  np.ascontiguousarray(a[::sj, 0]).tobytes(), np.ascontiguousarray(a[::sj, -1]).tobytes()]
    return b"".join(parts)


def fingerprint(MG, ST):
    """Hex digest identifying (grid, -D angle, periodicity, search parameters, track)."""
    h = hashlib.blake2b(digest_size=20)
    shape = tuple(int(v) for v in MG.shape)
    ang = getattr(MG, "xangle", None)
    with_angle = ang is not None and np.shape(ang) == shape
    h.update(repr((_VERSION, shape, int(MG.EWPer), float(MG.HResDeg), float(config.rfactor), float(config.deg2km),
                   float(config.search_box_w_km), with_angle, int(ST.size))).encode())
    for a in (MG.lat, MG.lon, ang if with_angle else np.zeros(0), ST.lat, ST.lon):
        h.update(_sample(np.asarray(a, dtype=np.float64) if np.asarray(a).dtype != np.float64 else a))
    return h.hexdigest()


def save_mapping(path, BT, MG, ST):
    """Write NP / SM / WB of `BT` (anything with those attributes, e.g. ``Model2SatTrack(...).BT``) with the fingerprint."""
    tmp = str(path) + ".tmp.npz"
    np.savez(tmp, NP=np.asarray(BT.NP, dtype=np.int64), SM=np.asarray(BT.SM, dtype=np.int64),
             WB=np.asarray(BT.WB, dtype=np.float64), fingerprint=np.array(fingerprint(MG, ST)))
    os.replace(tmp, str(path))


def load_mapping(path, MG, ST):
    """The cached mapping as a `MappingView`, or None when there is no file or it belongs to another grid / track."""
    if not os.path.exists(str(path)):
        return None
    try:
        with np.load(str(path), allow_pickle=False) as z:
            if str(z["fingerprint"]) != fingerprint(MG, ST):
                return None
            NP, SM, WB = z["NP"], z["SM"], z["WB"]
    except Exception:
        return None
    n = int(ST.size)
    if NP.shape != (n, 2) or SM.shape != (n, 4, 2) or WB.shape != (n, 4):
        return None
    return MappingView(NP, SM, WB)
