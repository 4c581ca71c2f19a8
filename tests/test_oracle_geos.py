"""Pins the shapely stand-in of the heavy-duty quadrant search to GEOS's published algorithm.

The reference calls ``shapely`` ``Polygon.contains(Point)`` (gonzag/bilin_mapping.py:415-423);
shapely is absent offline, so the oracle, the golden generator and the CUDA kernel use
`gz_oracle.point_strictly_in_quad` (fp64 cross products, boundary = outside).  `oracle/geos_contains.py`
restates the GEOS 3.8 call chain behind ``contains`` (envelope test, rectangle shortcut,
RayCrossingCounter with an exact orientation predicate) in exact rational arithmetic.  Here the two are
compared on random, grid-cell-like, lattice (degenerate) and adversarial (on / next to an edge) inputs:
they must agree everywhere except in two documented corners, and every disagreement is classified.
"""
import math
import random
from fractions import Fraction

from oracle.gz_oracle import point_strictly_in_quad as stand_in
from oracle.geos_contains import polygon_contains_point as geos_contains, _is_rectangle, orientation_index


def _fp_sign_wrong(py, px, qy, qx):
    """True when, for some edge, the fp64 cross product of the stand-in has not the exact sign."""
    for k in range(4):
        ay, ax, by, bx = qy[k], qx[k], qy[(k + 1) % 4], qx[(k + 1) % 4]
        side = (by - ay) * (px - ax) - (bx - ax) * (py - ay)
        exact = (Fraction(by) - Fraction(ay)) * (Fraction(px) - Fraction(ax)) \
            - (Fraction(bx) - Fraction(ax)) * (Fraction(py) - Fraction(ay))
        if ((side > 0) - (side < 0)) != ((exact > 0) - (exact < 0)):
            return True
    return False


def _cell_quad(rng):
    """A grid-cell-like quad (1/12 degree, corners jittered by up to 30 %, either orientation, any
    starting corner) and a target in or around it."""
    y0, x0, d = rng.uniform(-80, 80), rng.uniform(0, 360), 1.0 / 12
    c = [(0, 0), (0, 1), (1, 1), (1, 0)]
    if rng.random() < 0.5:
        c = c[::-1]
    k = rng.randrange(4)
    c = c[k:] + c[:k]
    qy = [y0 + d * (a + rng.uniform(-.3, .3)) for a, _ in c]
    qx = [x0 + d * (b + rng.uniform(-.3, .3)) for _, b in c]
    return y0 + d * rng.uniform(-.5, 1.5), x0 + d * rng.uniform(-.5, 1.5), qy, qx


def test_random_quads_agree():
    """Convex, concave and self-intersecting (bow-tie) quads: even-odd (GEOS) and the stand-in's
    non-zero winding coincide for four vertices."""
    rng = random.Random(11)
    inside = 0
    for _ in range(20000):
        qy = [rng.uniform(-1, 1) for _ in range(4)]
        qx = [rng.uniform(-1, 1) for _ in range(4)]
        py, px = rng.uniform(-1, 1), rng.uniform(-1, 1)
        want = geos_contains(py, px, qy, qx)
        inside += want
        assert stand_in(py, px, qy, qx) == want
    assert 1000 < inside < 5000


def test_grid_cell_quads_agree():
    rng = random.Random(12)
    inside = 0
    for _ in range(20000):
        py, px, qy, qx = _cell_quad(rng)
        want = geos_contains(py, px, qy, qx)
        inside += want
        assert stand_in(py, px, qy, qx) == want
    assert inside > 3000


def test_regular_grid_rectangles_agree():
    """Axis-parallel cells (1-D lat / lon model grids): GEOS takes its rectangle shortcut."""
    rng = random.Random(13)
    for _ in range(5000):
        y0, x0 = rng.randint(-80, 80) / 4.0, rng.randint(0, 1400) / 4.0
        c = [(0, 0), (0, 1), (1, 1), (1, 0)]
        if rng.random() < 0.5:
            c = c[::-1]
        qy = [y0 + 0.25 * a for a, _ in c]
        qx = [x0 + 0.25 * b for _, b in c]
        py = y0 + rng.choice([-0.125, 0.0, 0.0625, 0.125, 0.25, 0.375])
        px = x0 + rng.choice([-0.125, 0.0, 0.0625, 0.125, 0.25, 0.375])
        assert _is_rectangle([(qy[k], qx[k]) for k in (0, 1, 2, 3, 0)])
        assert stand_in(py, px, qy, qx) == geos_contains(py, px, qy, qx)


def test_vertices_and_exact_edge_points_are_outside():
    rng = random.Random(14)
    for _ in range(3000):
        _, _, qy, qx = _cell_quad(rng)
        k = rng.randrange(4)
        assert not geos_contains(qy[k], qx[k], qy, qx)
        assert not stand_in(qy[k], qx[k], qy, qx)
    # exact points on edges need exactly representable coordinates: a lattice
    for _ in range(20000):
        qy = [float(rng.randint(0, 4)) for _ in range(4)]
        qx = [float(rng.randint(0, 4)) for _ in range(4)]
        k = rng.randrange(4)
        py, px = 0.5 * (qy[k] + qy[(k + 1) % 4]), 0.5 * (qx[k] + qx[(k + 1) % 4])     # mid-point of an edge: exact
        ring = [(qy[j], qx[j]) for j in (0, 1, 2, 3, 0)]
        if _is_rectangle(ring) and len(set(ring)) < 4:
            continue                                  # the quirk of the next test
        assert not geos_contains(py, px, qy, qx)
        assert not stand_in(py, px, qy, qx)


def test_lattice_quads_differ_only_through_the_rectangle_quirk():
    """Degenerate quads on a lattice (repeated vertices, collapsed, bow-ties with exact hits).  The one
    disagreement: a ring with a REPEATED vertex whose five points all sit on envelope corners and whose
    edges are axis-parallel passes GEOS's ``Polygon::isRectangle`` although it has no area, and the
    rectangle shortcut then reports its whole envelope as interior.  Needs two coincident corners of a
    mesh (distinct grid cells with identical coordinates, axis-parallel): not a case the reference can
    reach on a grid whose duplicated columns are E-W overlap columns (a mesh never holds both copies)."""
    rng = random.Random(15)
    quirk = 0
    for _ in range(60000):
        qy = [float(rng.randint(0, 4)) for _ in range(4)]
        qx = [float(rng.randint(0, 4)) for _ in range(4)]
        py, px = rng.randint(0, 8) / 2.0, rng.randint(0, 8) / 2.0
        a, b = stand_in(py, px, qy, qx), geos_contains(py, px, qy, qx)
        if a != b:
            ring = [(qy[j], qx[j]) for j in (0, 1, 2, 3, 0)]
            assert _is_rectangle(ring) and len(set(ring)) < 4, (py, px, qy, qx)
            assert b and not a
            quirk += 1
    assert quirk > 0          # the corner exists; if this fires the restatement lost the shortcut


def test_points_next_to_an_edge_differ_only_when_fp64_loses_the_sign():
    """Targets built ON an edge (rounded to double) and nudged by -1 / 0 / +1 ulp: GEOS decides the side
    with an exact predicate, the stand-in (and the CUDA kernel) with fp64 products.  A disagreement is
    allowed only where an fp64 cross product has the wrong sign, i.e. the target is within rounding
    error (~1e-16 relative) of an edge; the share of such cases is bounded."""
    rng = random.Random(16)
    n, diff = 40000, 0
    for _ in range(n):
        _, _, qy, qx = _cell_quad(rng)
        k = rng.randrange(4)
        s = rng.random()
        py = qy[k] + s * (qy[(k + 1) % 4] - qy[k])
        px = qx[k] + s * (qx[(k + 1) % 4] - qx[k])
        nudge = rng.choice([-1, 0, 1])
        if nudge:
            py = math.nextafter(py, math.inf * nudge)
        a, b = stand_in(py, px, qy, qx), geos_contains(py, px, qy, qx)
        if a != b:
            assert _fp_sign_wrong(py, px, qy, qx), (py, px, qy, qx)
            diff += 1
    assert diff < n // 1000


def test_orientation_predicate():
    assert orientation_index((0, 0), (1, 0), (0.5, 1e-300)) == 1
    assert orientation_index((0, 0), (1, 0), (0.5, -1e-300)) == -1
    assert orientation_index((0, 0), (1, 1), (0.5, 0.5)) == 0
    # a case plain fp64 gets wrong: the products differ by less than their rounding
    p1, p2 = (0.1, 0.1), (0.3, 0.3)
    q = (0.2, math.nextafter(0.2, 1.0))
    assert orientation_index(p1, p2, q) == 1


def test_heavy_duty_points_of_the_golden_case_get_the_same_quadrant(golden_global):
    """Every track point of the golden global -D case that enters the heavy-duty search (no light
    quadrant, or |grid angle| > 45 degrees): the oracle's quadrant with the GEOS restatement injected
    equals the one the golden vectors were generated with (stand-in injected into the reference)."""
    import numpy as np
    from oracle import gz_oracle as orc
    g = golden_global
    lat, lon, ang, kew = g["lat"], g["lon"], g["angle"], int(g["kew"])
    NP, SM = g["NP"], g["SM"]
    heavy = 0
    for jt in range(len(NP)):
        jP, iP = int(NP[jt, 0]), int(NP[jt, 1])
        if jP < 0:
            continue
        a = float(ang[jP, iP])
        pt = (float(g["ysat"][jt]), float(g["xsat"][jt]))
        calls = []

        def spy(py, px, qy, qx, calls=calls):
            calls.append(1)
            return geos_contains(py, px, qy, qx)

        iq = orc.quadrant(pt, lat, lon, jP, iP, k_ew_per=kew, grid_s_angle=a, contains=spy)
        if not calls:
            continue
        heavy += 1
        if iq in (1, 2, 3, 4):
            mesh = [(jP, iP)] + list(orc.mesh_from_quadrant(jP, iP, lat.shape[0], lat.shape[1], iq, k_ew_per=kew))
        else:
            mesh = [(-1, -1)] * 4
        assert np.array_equal(np.array(mesh, dtype=np.int64), SM[jt]), jt
    assert heavy > 50
