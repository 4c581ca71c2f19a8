"""CPU oracle for the gridded -> along-track mapping + interpolation path of brodeau/gonzag.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product (``gonzag_b200``) never does: its hot path is
the CUDA library behind ``include/gonzag_b200.h`` and fails loudly without it.

What it is: a NumPy / ``math`` restatement of the reference's algorithm, function by
function, keeping the reference's floating-point operation order and library calls
(``numpy`` ufuncs where the reference uses arrays, ``math.*`` where it uses scalars,
``numpy.linalg.det`` in the Newton solve) so that on the same machine it returns the
same bits as the reference.  Each function cites the reference lines it follows
(paths relative to the reference checkout).

Parity pinning: the reference ships no golden vectors for this path,
so the oracle is pinned against the *imported reference itself*:
``tests/golden/make_golden.py`` runs the unmodified reference (in the build container,
where it is mounted) on seeded synthetic inputs and commits the outputs as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against them
bit-for-bit, and ``tests/test_oracle_vs_reference.py`` re-runs the comparison live
whenever the reference checkout is present.

PARITY UNPINNED (by the reference's own vectors) for one branch: the "heavy-duty"
quadrant search (gonzag/bilin_mapping.py:407-426) calls ``shapely`` (GEOS
``Polygon.contains``), an un-vendored dependency (pinned ``shapely=1.7.1`` in
util/conda/environment-zarr.yml:17) that is absent offline, and the reference holds no
test of it.  `point_strictly_in_quad` is the stand-in (planar, interior only, boundary
= outside); the golden vectors for that branch were produced with the same stand-in
injected into the reference.  It is pinned instead to GEOS's published algorithm:
``oracle/geos_contains.py`` restates the call chain behind ``contains`` (envelope test,
rectangle shortcut, RayCrossingCounter with an exact orientation predicate) in exact
arithmetic, and ``tests/test_oracle_geos.py`` holds the stand-in to it on random,
grid-cell-like, degenerate and adversarial inputs.  Two documented differences remain:
targets within fp64 rounding (~1e-16 relative) of a mesh edge, and zero-area rings with
a repeated vertex that GEOS's ``isRectangle`` mistakes for rectangles.
"""
from __future__ import annotations

import math
from math import atan2, copysign, cos, log, pi, radians, sin, sqrt, tan, asin

import numpy as np

# ---- constants (reference gonzag/config.py:20-31) -------------------------------
DEG2KM = 111.11
R_EQ = 6378.137
R_PL = 6356.752
RMISSVAL = -9999.0
RFACTOR = 0.75
SEARCH_BOX_W_KM = 500.0
R_HAVERSINE = 6360.0        # gonzag/utils.py:310


# ================================================================== small helpers
def frame_pm180(x):
    """0..360 -> -180..180 longitude frame.  gonzag/utils.py:24-33 (`degE_to_degWE`)."""
    if np.shape(x) == ():
        return copysign(1.0, 180.0 - x) * min(x, abs(x - 360.0))
    return np.copysign(1.0, 180.0 - x) * np.minimum(x, np.abs(x - 360.0))


def haversine_km(plat, plon, glat, glon):
    """Distance (km) from one point to an array of points.  gonzag/utils.py:296-316."""
    d2r = pi / 180.0
    s_lat = np.sin(0.5 * ((glat - plat) * d2r))
    s_lon = np.sin(0.5 * ((glon - plon) * d2r))
    cc = np.cos(glat * d2r) * cos(plat * d2r)
    return 2.0 * R_HAVERSINE * np.arcsin(np.sqrt(s_lat * s_lat + cc * s_lon * s_lon))


def argmin_ji(a):
    """First-occurrence flat argmin -> (row, col).  gonzag/utils.py:48-55."""
    k = a.argmin()
    w = a.shape[1]
    return k // w, k % w


def search_box_radius(res_km, width_km):
    """gonzag/utils.py:97-110 (`SearchBoxSize`)."""
    return int(0.5 * width_km / res_km)


def heading(lat_a, lon_a, lat_b, lon_b):
    """Loxodromic bearing a->b in degrees, clockwise from north, in [0,360).
    gonzag/bilin_mapping.py:16-47."""
    tiny = 1.0e-9
    xa = radians(lon_a)
    xb = radians(lon_b)
    ya = -log(max(abs(tan(pi / 4.0 - radians(lat_a) / 2.0)), tiny))
    yb = -log(max(abs(tan(pi / 4.0 - radians(lat_b) / 2.0)), tiny))
    dlam = (xb - xa) % (2 * pi)
    if dlam >= pi:
        dlam = dlam - 2 * pi
    if dlam <= -pi:
        dlam = dlam + 2 * pi
    return (atan2(dlam, yb - ya) * 180.0 / pi) % 360.0


# ================================================================== Newton solve
def local_coords(vy, vx):
    """(alpha, beta) of point 0 inside the quad 1-2-3-4.  gonzag/bilin_mapping.py:50-141.

    ``vy``/``vx`` are length-5 float64 arrays (target first).  Quirks kept: first
    residual is hard-coded to (0.5, 0.5); the loop stops when the *last update* is
    <= 0.1; a zero determinant spins to the iteration cap; give-up and the
    four-equal-latitudes case both return (0.5, 0.5)."""
    max_iter = 100
    wrap = (abs(vx[1] - vx[4]) >= 180.0 or abs(vx[1] - vx[2]) >= 180.0
            or abs(vx[1] - vx[3]) >= 180.0)
    if wrap:
        saved = np.zeros(5)
        saved[:] = vx[:]
        vx[:] = frame_pm180(saved[:])

    step = 1000.0
    ex = 0.5
    ey = 0.5
    a = 0.0
    b = 0.0
    jac = np.zeros((2, 2))
    tmp = np.zeros((2, 2))
    it = 0
    while (step > 0.1) and (it < max_iter):
        e12 = vx[2] - vx[1]
        e41 = vx[1] - vx[4]
        e23 = vx[3] - vx[2]
        jac[0, 0] = e12 + (e41 + e23) * b
        jac[0, 1] = -e41 + (e41 + e23) * a
        jac[1, 0] = vy[2] - vy[1] + (vy[1] - vy[4] + vy[3] - vy[2]) * b
        jac[1, 1] = vy[4] - vy[1] + (vy[1] - vy[4] + vy[3] - vy[2]) * a
        det = np.linalg.det(jac)
        if det == 0.0:
            pass            # the reference's "give up" statement is a no-op comparison
        else:
            tmp[:, 0] = [ex, ey]
            tmp[:, 1] = jac[:, 1]
            da = np.linalg.det(tmp) / det
            tmp[:, 0] = jac[:, 0]
            tmp[:, 1] = [ex, ey]
            db = np.linalg.det(tmp) / det
            step = sqrt(da * da + db * db)
            a = a + da
            b = b + db
            ca = 1.0 - a
            cb = 1.0 - b
            ex = vx[0] - (ca * cb * vx[1] + a * cb * vx[2] + a * b * vx[3] + ca * b * vx[4])
            ey = vy[0] - (ca * cb * vy[1] + a * cb * vy[2] + a * b * vy[3] + ca * b * vy[4])
        it = it + 1

    if it >= max_iter:
        a = 0.5
        b = 0.5
    if vy[1] == vy[2] and vy[2] == vy[3] and vy[3] == vy[4]:
        a = 0.5
        b = 0.5
    if wrap:
        vx[:] = saved[:]
    return a, b


# ================================================================== nearest point
def nearest_point(pt, Ys, Xs, rd_found_km=10.0, j_prv=None, i_prv=None, np_box_r=10,
                  max_itr=5):
    """gonzag/bilin_mapping.py:144-191 (the never-used ``resolkm`` branch is dropped).

    Box pass around the previous hit (plain truthiness of the previous indices decides
    whether there is a box; the box is clamped, never wrapped), then whole-grid passes
    with the acceptance radius growing by 1.2 each time; a hit that needs the last pass
    is discarded."""
    yT, xT = pt
    Ny, Nx = Ys.shape
    use_box = (j_prv and i_prv)
    if use_box:
        ja, jb = max(j_prv - np_box_r, 0), min(j_prv + np_box_r + 1, Ny)
        ia, ib = max(i_prv - np_box_r, 0), min(i_prv + np_box_r + 1, Nx)
    else:
        ja, ia, jb, ib = 0, 0, Ny, Nx
    jy, jx = -1, -1
    found = False
    radius = rd_found_km
    n = 0
    while (not found) and n < max_itr:
        n += 1
        if use_box and n > 1:
            ja, ia, jb, ib = 0, 0, Ny, Nx
        dist = haversine_km(yT, xT, Ys[ja:jb, ia:ib], Xs[ja:jb, ia:ib])
        jy, jx = argmin_ji(dist)
        if (not use_box) and n == 1:
            n = 2
        found = (dist[jy, jx] < radius)
        if n > 1 and not found:
            radius = 1.2 * radius
    if use_box:
        jy, jx = jy + ja, jx + ia
    if jy < 0 or jx < 0 or jy >= Ny or jx >= Nx or n == max_itr:
        jy, jx = -1, -1
    return [jy, jx]


def edge_exclusion(jj, ji, Nj, Ni, k_ew_per):
    """gonzag/bilin_mapping.py:582-585 (operator precedence kept: column 0 is always excluded)."""
    if ji == 0 or ji == Ni - 1 and k_ew_per == -1:
        jj, ji = -1, -1
    if jj == 0 or jj == Nj - 1:
        jj, ji = -1, -1
    return jj, ji


# ================================================================== source mesh
def neighbours(jP, iP, Ny, Nx, k_ew_per=-1):
    """gonzag/bilin_mapping.py:194-212 (`surrounding_j_i`); -1 flags a missing index."""
    jm, jp, im, ip = jP - 1, jP + 1, iP - 1, iP + 1
    if jp == Ny:
        jp = -1
    if im == -1:
        if k_ew_per >= 0:
            im = Nx - k_ew_per - 1
    if ip == Nx:
        ip = -1
        if k_ew_per >= 0:
            ip = k_ew_per
    return jm, jp, im, ip


def mesh_from_quadrant(jP, iP, Ny, Nx, iq, k_ew_per=1):
    """gonzag/bilin_mapping.py:215-262 (`Iquadran2SrcMesh`): the three other corners,
    counter-clockwise from the nearest point."""
    jm, jp, im, ip = neighbours(jP, iP, Ny, Nx, k_ew_per=k_ew_per)
    p2 = p3 = p4 = (-1, -1)
    if -1 not in (jm, jp, im, ip):
        if iq == 1:
            p2, p3, p4 = (jP, ip), (jp, ip), (jp, iP)
        elif iq == 2:
            p2, p3, p4 = (jm, iP), (jm, ip), (jP, ip)
        elif iq == 3:
            p2, p3, p4 = (jP, im), (jm, im), (jm, iP)
        elif iq == 4:
            p2, p3, p4 = (jp, iP), (jp, im), (jP, im)
    return p2, p3, p4


def side_of_exit(yA, xA, yB, xB, vY, vX):
    """gonzag/bilin_mapping.py:267-312 (`IQH`, exported but not called on the path): the side of the quad
    (corners bottom-left, bottom-right, upper-right, upper-left) the segment A -> B leaves through, 1..4
    = down, right, up, left; -1 where the reference prints an error and exits.  Four independent tests on
    un-wrapped headings, last true wins -- as written there, only 'up' can ever be true for a quad whose
    corners sit in their usual directions."""
    lonA = xA % 360.0
    ht = heading(yA, lonA, yB, xB % 360.0)
    hBL = heading(yA, lonA, vY[0], vX[0] % 360.0)
    hBR = heading(yA, lonA, vY[1], vX[1] % 360.0)
    hUR = heading(yA, lonA, vY[2], vX[2] % 360.0)
    hUL = heading(yA, lonA, vY[3], vX[3] % 360.0)
    idir = -1
    if ht > hBL and ht < hBR:
        idir = 1
    if ht > hBR and ht < hUR:
        idir = 2
    if ht > hUR and ht < hUL:
        idir = 3
    if ht > hUL and ht < hBL:
        idir = 4
    return idir


def point_strictly_in_quad(py, px, qy, qx):
    """Planar point-in-quadrilateral, interior only (a point on an edge or vertex is
    outside).  Stand-in for ``shapely`` ``Polygon.contains(Point)`` at
    gonzag/bilin_mapping.py:421-423.  Winding-number walk over the four edges using the
    sign of the edge cross product; no division, products and one subtraction only, so
    the CUDA kernel can reproduce it bit-for-bit."""
    wn = 0
    for k in range(4):
        ay, ax = qy[k], qx[k]
        by, bx = qy[(k + 1) % 4], qx[(k + 1) % 4]
        side = (by - ay) * (px - ax) - (bx - ax) * (py - ay)
        if side == 0.0:
            if (min(ay, by) <= py <= max(ay, by)) and (min(ax, bx) <= px <= max(ax, bx)):
                return False
        if ax <= px:
            if bx > px and side > 0.0:
                wn += 1
        else:
            if bx <= px and side < 0.0:
                wn -= 1
    return wn != 0


def quadrant(pt, Ys, Xs, jP, iP, k_ew_per=1, grid_s_angle=0.0, contains=None):
    """Which of the four cells around the nearest point holds the target (1..4), -1 if
    undecided, None if a neighbour is missing.  gonzag/bilin_mapping.py:329-449.

    Light test: five loxodromic headings, four independent tests, the last true one
    wins.  Heavy-duty test (no quadrant found, or local grid rotation above 45 degrees):
    quads 4,3,2,1 in the (lat, lon) plane, first strict-interior hit wins, otherwise the
    light answer stands."""
    yT, xT = pt
    Ny, Nx = Ys.shape
    iq = -1
    jm, jp, im, ip = neighbours(jP, iP, Ny, Nx, k_ew_per=k_ew_per)
    if -1 in (jm, jp, im, ip):
        return None
    if (im < 0) or (jm < 0) or (ip > Nx - 1):
        return None

    lon0 = Xs[jP, iP] % 360.0
    lat0 = Ys[jP, iP]
    zT = xT % 360.0
    ht = heading(lat0, lon0, yT, zT)
    hN = heading(lat0, lon0, Ys[jp, iP], Xs[jp, iP] % 360.0)
    hE = heading(lat0, lon0, Ys[jP, ip], Xs[jP, ip] % 360.0)
    hS = heading(lat0, lon0, Ys[jm, iP], Xs[jm, iP] % 360.0)
    hW = heading(lat0, lon0, Ys[jP, im], Xs[jP, im] % 360.0)
    hz = frame_pm180(ht)
    if hN > hE:
        hN = hN - 360.0
    if hz > hN and hz <= hE:
        iq = 1
    if ht > hE and ht <= hS:
        iq = 2
    if ht > hS and ht <= hW:
        iq = 3
    if ht > hW and hz <= hN:
        iq = 4

    if (iq not in (1, 2, 3, 4)) or (abs(grid_s_angle) > 45.0):
        test = contains if contains is not None else point_strictly_in_quad
        for cand in (4, 3, 2, 1):
            (j2, i2), (j3, i3), (j4, i4) = mesh_from_quadrant(jP, iP, Ny, Nx, cand, k_ew_per=k_ew_per)
            qy = (Ys[jP, iP], Ys[j2, i2], Ys[j3, i3], Ys[j4, i4])
            qx = (Xs[jP, iP], Xs[j2, i2], Xs[j3, i3], Xs[j4, i4])
            if test(yT, xT, qy, qx):
                iq = cand
                break
    return iq


def source_mesh(pt, Ys, Xs, jP, iP, k_ew_per=-1, grid_s_angle=0.0):
    """(4,2) int64 mesh P1(nearest),P2,P3,P4 or all -1.  gonzag/bilin_mapping.py:453-486."""
    Ny, Nx = Ys.shape
    iq = quadrant(pt, Ys, Xs, jP, iP, k_ew_per=k_ew_per, grid_s_angle=grid_s_angle)
    if iq in (1, 2, 3, 4):
        p1 = (jP, iP)
        p2, p3, p4 = mesh_from_quadrant(jP, iP, Ny, Nx, iq, k_ew_per=k_ew_per)
    else:
        p1 = p2 = p3 = p4 = (-1, -1)
    return np.array([p1, p2, p3, p4], dtype=np.int64)


def bilinear_weights(pt, Ys, Xs, mesh):
    """Four weights for one target.  gonzag/bilin_mapping.py:490-527."""
    yT, xT = pt
    [[j1, i1], [j2, i2], [j3, i3], [j4, i4]] = mesh[:, :]
    if j1 >= 0:
        c = np.zeros((2, 5))
        c[:, 0] = [yT, xT]
        c[:, 1] = [Ys[j1, i1], Xs[j1, i1]]
        c[:, 2] = [Ys[j2, i2], Xs[j2, i2]]
        c[:, 3] = [Ys[j3, i3], Xs[j3, i3]]
        c[:, 4] = [Ys[j4, i4], Xs[j4, i4]]
        a, b = local_coords(c[0, :], c[1, :])
        w1 = (1.0 - a) * (1.0 - b)
        w2 = a * (1.0 - b)
        w3 = a * b
        w4 = (1.0 - a) * b
    else:
        w1, w2, w3, w4 = 1.0, 0.0, 0.0, 0.0
    return np.array([w1, w2, w3, w4])


# ================================================================== the mapping object
def nearest_chain(Yt, Xt, Ys, Xs, k_ew_per, rd_found_km, np_box_r, seed=None, idx=None):
    """Sequentially seeded nearest points.  gonzag/bilin_mapping.py:566-593.
    ``seed`` (default ``(r, r)``) is the pair handed to the first point."""
    Nj, Ni = Ys.shape
    n = len(Yt)
    out = np.zeros((n, 2), dtype=np.int64)
    jj, ji = (np_box_r, np_box_r) if seed is None else seed
    for t in range(n):
        jj, ji = nearest_point((Yt[t], Xt[t]), Ys, Xs, rd_found_km=rd_found_km,
                               j_prv=jj, i_prv=ji, np_box_r=np_box_r)
        jj, ji = edge_exclusion(jj, ji, Nj, Ni, k_ew_per)
        out[t, :] = [jj, ji]
    return out


def source_meshes(Yt, Xt, Ys, Xs, NP, k_ew_per, angle=None, idx=None):
    """gonzag/bilin_mapping.py:595-608.  Rows whose nearest point is (-1,-1) stay ZERO.
    ``idx`` restricts the work to a subset of track points (rows of the result)."""
    pts = range(len(Yt)) if idx is None else idx
    out = np.zeros((len(pts), 4, 2), dtype=np.int64)
    has_angle = angle is not None and np.shape(angle) == np.shape(Ys)
    for r, t in enumerate(pts):
        jP, iP = NP[t, :]
        if [jP, iP] != [-1, -1]:
            ang = angle[jP, iP] if has_angle else 0.0
            out[r, :, :] = source_mesh((Yt[t], Xt[t]), Ys, Xs, jP, iP, k_ew_per=k_ew_per,
                                       grid_s_angle=ang)
    return out


def weights(Yt, Xt, Ys, Xs, SM, idx=None):
    """gonzag/bilin_mapping.py:610-617.  ``SM`` rows align with ``idx`` when given."""
    pts = range(len(Yt)) if idx is None else idx
    out = np.zeros((len(pts), 4))
    for r, t in enumerate(pts):
        out[r, :] = bilinear_weights((Yt[t], Xt[t]), Ys, Xs, SM[r, :, :])
    return out


class TrackMapping:
    """Same results as the reference's `BilinTrack` (gonzag/bilin_mapping.py:530-617)."""

    def __init__(self, Yt, Xt, Ys, Xs, src_grid_local_angle=(), k_ew_per=-1,
                 rd_found_km=100.0, np_box_r=4, seed=None):
        self.Nt = len(Yt)
        self.Nj, self.Ni = Ys.shape
        ang = src_grid_local_angle if np.shape(src_grid_local_angle) == np.shape(Ys) else None
        self.NP = nearest_chain(Yt, Xt, Ys, Xs, k_ew_per, rd_found_km, np_box_r, seed=seed)
        self.SM = source_meshes(Yt, Xt, Ys, Xs, self.NP, k_ew_per, angle=ang)
        self.WB = weights(Yt, Xt, Ys, Xs, self.SM)


# ================================================================== grid angle (-D)
def grid_angle(xlat, xlon, j0=None, j1=None):
    """Local rotation of the grid's j-direction w.r.t. north, degrees.
    gonzag/utils.py:226-262.  Rows 0 and Ny-1 stay zero.  ``j0:j1`` limits the rows
    computed (the others are left at zero) so a band can be timed on large grids."""
    d2r = pi / 180.0
    q = pi / 4.0
    Ny, Nx = np.shape(xlat)
    out = np.zeros((Ny, Nx))
    ja = 1 if j0 is None else max(j0, 1)
    jb = Ny - 1 if j1 is None else min(j1, Ny - 1)
    for i in range(Nx):
        for j in range(ja, jb):
            if abs(xlon[j + 1, i] % 360.0 - xlon[j - 1, i] % 360.0) < 1.0e-8:
                continue
            t0 = tan(q - d2r * xlat[j, i] / 2.0)
            px = 0.0 - 2.0 * cos(d2r * xlon[j, i]) * t0
            py = 0.0 - 2.0 * sin(d2r * xlon[j, i]) * t0
            pn = px * px + py * py
            t1 = tan(q - d2r * xlat[j + 1, i] / 2.0)
            t2 = tan(q - d2r * xlat[j - 1, i] / 2.0)
            vx = 2.0 * (cos(d2r * xlon[j + 1, i]) * t1 - cos(d2r * xlon[j - 1, i]) * t2)
            vy = 2.0 * (sin(d2r * xlon[j + 1, i]) * t1 - sin(d2r * xlon[j - 1, i]) * t2)
            nv = sqrt(pn * (vx * vx + vy * vy))
            nv = max(nv, 1.0e-12)
            s = (px * vy - py * vx) / nv
            c = (px * vx + py * vy) / nv
            out[j, i] = atan2(s, c) * 180.0 / pi
    return out


# ================================================================== interpolation
def interp_track(t_sat, t_mod, get_record, mask, SM, WB, faithful_cost=False,
                 rmissval=RMISSVAL, j_slice=None):
    """Time-linear + space-bilinear interpolation.  gonzag/mod2sat.py:74-139.

    ``get_record(k)`` returns record ``k`` of the model field as a 2-D array (the
    reference's `GetModel2DVar`).  With ``faithful_cost=True`` the time interpolation is
    evaluated on the WHOLE grid for every track point exactly like gonzag/mod2sat.py:119
    (this is what the CPU baseline times); otherwise only the four corners are
    evaluated, with the same element-wise arithmetic, hence the same bits.
    Returns ``(ssh_np, ssh_bl)``, float64, untouched entries = ``rmissval``."""
    n = len(t_sat)
    out_np = np.zeros(n)
    out_np[:] = rmissval
    out_bl = np.zeros(n)
    out_bl[:] = rmissval
    k1 = 0
    k1_old, k2_old = -10, -10
    X1 = X2 = slope = None
    for t in range(n):
        now = t_sat[t]
        k = k1
        while not (t_mod[k] <= now and t_mod[k + 1] > now):
            k = k + 1
        k1, k2 = k, k + 1
        if (k1 > k1_old) and (k2 > k2_old):
            if (k1_old == -10) or (k1 > k2_old):
                X1 = get_record(k1)
            else:
                X1 = X2
            X2 = get_record(k2)
            slope = (X2 - X1) / float(t_mod[k2] - t_mod[k1])
        dt = float(now - t_mod[k1])
        [[j1, i1], [j2, i2], [j3, i3], [j4, i4]] = SM[t, :, :]
        [w1, w2, w3, w4] = WB[t, :]
        if faithful_cost:
            Xm = X1[:, :] + slope[:, :] * dt
            v1, v2, v3, v4 = Xm[j1, i1], Xm[j2, i2], Xm[j3, i3], Xm[j4, i4]
        else:
            v1 = X1[j1, i1] + slope[j1, i1] * dt
            v2 = X1[j2, i2] + slope[j2, i2] * dt
            v3 = X1[j3, i3] + slope[j3, i3] * dt
            v4 = X1[j4, i4] + slope[j4, i4] * dt
        ocean = mask[j1, i1] + mask[j2, i2] + mask[j3, i3] + mask[j4, i4]
        if ocean == 4:
            out_np[t] = v1
            sw = np.sum([w1, w2, w3, w4])
            if not (abs(sw - 1.0) > 0.001):
                out_bl[t] = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4
        k1_old, k2_old = k1, k2
    return out_np, out_bl


def earth_radius_km(lat):
    """gonzag/utils.py:267-278."""
    p = radians(lat)
    c = (R_EQ ** 2 * cos(p)) ** 2
    d = (R_PL ** 2 * sin(p)) ** 2
    e = (R_EQ * cos(p)) ** 2
    f = (R_PL * sin(p)) ** 2
    return sqrt((c + d) / (e + f))


def haversine_scalar_km(lat1, lon1, lat2, lon2):
    """gonzag/utils.py:281-293."""
    R = earth_radius_km(0.5 * (lat1 + lat2))
    dlat = radians(lat2 - lat1)
    dlon = radians(lon2 - lon1)
    p1 = radians(lat1)
    p2 = radians(lat2)
    a = sin(dlat / 2) ** 2 + cos(p1) * cos(p2) * sin(dlon / 2) ** 2
    return R * (2 * asin(sqrt(a)))


def along_track_distance(lat, lon):
    """Cumulative distance from the first point, km.  gonzag/mod2sat.py:146-148."""
    n = len(lat)
    d = np.zeros(n)
    for t in range(1, n):
        d[t] = d[t - 1] + haversine_scalar_km(lat[t], lon[t], lat[t - 1], lon[t - 1])
    return d


def output_mask(ssh_bl, ssh_sat):
    """int8 validity flag.  gonzag/mod2sat.py:157-158."""
    m = np.zeros(len(ssh_bl), dtype=np.int8)
    m[np.where((ssh_bl > -100.0) & (ssh_bl < 100.0) & (ssh_sat > -100.0) & (ssh_sat < 100.0))] = 1
    return m


def xnp_track(NP, mask, rmissval=RMISSVAL):
    """Map of the last track index hitting each cell.  gonzag/mod2sat.py:177-186.
    Not-found points write cell [-1,-1]; land = -100; untouched ocean is masked."""
    Nj, Ni = mask.shape
    x = np.zeros((Nj, Ni))
    x[:, :] = rmissval
    for t in range(NP.shape[0]):
        jj, ji = NP[t, :]
        x[jj, ji] = float(t)
    x[np.where(mask == 0)] = -100.0
    return np.ma.masked_where(x == rmissval, x)


def track_products(MG_lat, MG_lon, MG_mask, MG_time, get_record, ST_time, ST_lat, ST_lon,
                   ssh_sat, res_deg, angle=(), k_ew_per=-1, faithful_cost=False):
    """Everything `Model2SatTrack` exposes (gonzag/mod2sat.py:20-186) as a dict."""
    rd = RFACTOR * res_deg * DEG2KM
    r = search_box_radius(res_deg * DEG2KM, SEARCH_BOX_W_KM)
    bt = TrackMapping(ST_lat, ST_lon, MG_lat, MG_lon, src_grid_local_angle=angle,
                      k_ew_per=k_ew_per, rd_found_km=rd, np_box_r=r)
    v_np, v_bl = interp_track(ST_time, MG_time, get_record, MG_mask, bt.SM, bt.WB,
                              faithful_cost=faithful_cost)
    dist = along_track_distance(ST_lat, ST_lon)
    im = output_mask(v_bl, ssh_sat)
    return dict(NP=bt.NP, SM=bt.SM, WB=bt.WB, ssh_np_raw=v_np, ssh_bl_raw=v_bl, mask=im,
                ssh_mod_np=np.ma.masked_where(im == 0, v_np),
                ssh_mod=np.ma.masked_where(im == 0, v_bl),
                ssh_sat=np.ma.masked_where(im == 0, ssh_sat),
                distance=dist, XNPtrack=xnp_track(bt.NP, MG_mask),
                rd_found_km=rd, np_box_r=r)
