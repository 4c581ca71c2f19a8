"""Restatement of what ``shapely`` does at gonzag/bilin_mapping.py:415-423:
``Polygon([...4 points...]).contains(Point(yT, xT))``.

THIS IS TEST INFRASTRUCTURE (see oracle/gz_oracle.py): it exists to pin
`gz_oracle.point_strictly_in_quad` -- the stand-in used by the oracle, by the golden
generator and mirrored by the CUDA kernel -- to the algorithm of the library the
reference really calls, which is absent offline.

Dependency: shapely 1.7.1 (util/conda/environment-zarr.yml:17), whose wheels bundle
GEOS 3.8.  Call chain, from the published GEOS sources (not under /root/reference):

  shapely ``BaseGeometry.contains``  ->  ``GEOSContains_r``
  -> ``geos::geom::Geometry::contains(g)``
       * false when the envelope of ``this`` does not contain the envelope of ``g``;
       * ``RectangleContains`` when ``this`` is an axis-parallel rectangle
         (inside the envelope and not on its boundary);
       * else ``relate(g)->isContains()`` (pattern ``T*****FF*``).
  -> for a Polygon / Point pair the only entry of the matrix that matters is the
     location of the point: ``RelateComputer`` labels the isolated node with
     ``PointLocator::locate`` -> ``locateInPolygonRing`` -> ``PointLocation::locateInRing``
     -> ``algorithm::RayCrossingCounter::locatePointInRing``.
     contains  <=>  location == INTERIOR  (BOUNDARY and EXTERIOR are "not contained").

``RayCrossingCounter`` counts the crossings of the closed ring with the ray that leaves
the test point in the +x direction (even-odd rule), with these conventions
(``RayCrossingCounter::countSegment``):

  1. a segment strictly left of the point is ignored;
  2. point == second end of a segment               -> on the boundary;
  3. horizontal segment at the height of the point  -> boundary iff x within its extent,
     never counted;
  4. an upward edge owns its start, a downward edge owns its end (half-open in y);
     the side is decided by ``CGAlgorithmsDD::orientationIndex`` -- a double-double
     evaluation that returns the EXACT sign of the orientation determinant for
     double inputs (0 -> on the boundary).

Shapely's constructor closes the ring (first point appended) and passes coordinates
through unchanged, the FIRST tuple member as x: here x = latitude, y = longitude,
which only fixes the direction of the ray.

`locate_in_ring` follows the list above with exact rational arithmetic for item 4
(``fractions.Fraction`` of doubles is exact), so it returns what GEOS returns for
every double input, including the degenerate ones (bow-tie quads, collapsed quads,
repeated vertices, points on edges or vertices).
"""
from __future__ import annotations

from fractions import Fraction

EXTERIOR, BOUNDARY, INTERIOR = 2, 1, 0        # geos::geom::Location values


def orientation_index(p1, p2, q):
    """Exact sign of (p2-p1) x (q-p1): +1 q left of p1->p2, -1 right, 0 collinear.
    GEOS ``CGAlgorithmsDD::orientationIndex``."""
    x1, y1 = Fraction(p1[0]), Fraction(p1[1])
    x2, y2 = Fraction(p2[0]), Fraction(p2[1])
    qx, qy = Fraction(q[0]), Fraction(q[1])
    det = (x2 - x1) * (qy - y1) - (y2 - y1) * (qx - x1)
    return (det > 0) - (det < 0)


def locate_in_ring(pt, ring):
    """``RayCrossingCounter::locatePointInRing``: `ring` is a closed list of (x, y)."""
    px, py = pt
    crossings = 0
    for k in range(1, len(ring)):
        p1, p2 = ring[k], ring[k - 1]          # GEOS feeds (ring[i], ring[i-1])
        if p1[0] < px and p2[0] < px:
            continue
        if px == p2[0] and py == p2[1]:
            return BOUNDARY
        if p1[1] == py and p2[1] == py:
            lo, hi = (p1[0], p2[0]) if p1[0] <= p2[0] else (p2[0], p1[0])
            if lo <= px <= hi:
                return BOUNDARY
            continue
        if (p1[1] > py and p2[1] <= py) or (p2[1] > py and p1[1] <= py):
            sign = orientation_index(p1, p2, pt)
            if sign == 0:
                return BOUNDARY
            if p2[1] < p1[1]:
                sign = -sign
            if sign > 0:
                crossings += 1
    return INTERIOR if crossings % 2 == 1 else EXTERIOR


def _is_rectangle(ring):
    """``Polygon::isRectangle``: 5 points, every edge axis-parallel, the envelope's corners."""
    if len(ring) != 5:
        return False
    xs = [p[0] for p in ring]
    ys = [p[1] for p in ring]
    x0, x1, y0, y1 = min(xs), max(xs), min(ys), max(ys)
    for p in ring[:4]:
        if not (p[0] == x0 or p[0] == x1):
            return False
        if not (p[1] == y0 or p[1] == y1):
            return False
    for k in range(4):
        a, b = ring[k], ring[k + 1]
        if (a[0] != b[0]) == (a[1] != b[1]):   # each edge changes exactly one coordinate
            return False
    return True


def polygon_contains_point(py, px, qy, qx):
    """``Polygon([(qy[k], qx[k]) for k in 0..3]).contains(Point(py, px))`` as GEOS 3.8
    evaluates it.  Argument order matches `gz_oracle.point_strictly_in_quad`."""
    ring = [(float(qy[k]), float(qx[k])) for k in range(4)]
    ring.append(ring[0])
    pt = (float(py), float(px))
    xs = [p[0] for p in ring]
    ys = [p[1] for p in ring]
    if not (min(xs) <= pt[0] <= max(xs) and min(ys) <= pt[1] <= max(ys)):
        return False                            # envelope short-circuit
    if _is_rectangle(ring):                     # RectangleContains: inside and not on the boundary
        return min(xs) < pt[0] < max(xs) and min(ys) < pt[1] < max(ys)
    return locate_in_ring(pt, ring) == INTERIOR
