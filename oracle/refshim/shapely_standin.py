"""TEST INFRASTRUCTURE -- exact-geometry stand-ins for the few `shapely` calls the
reference makes on the hot path (crowd_sim/envs/utils/helper.py:5-6,42-55,
164-231).  shapely/GEOS is not installable in this container.

Semantics chosen (documented in DESIGN.md, restated identically in
oracle/crowd_oracle.c and in the CUDA kernel):

* ``Point(x, y).buffer(r)`` is the exact closed disc.  GEOS builds a 64-gon
  whose axis-aligned extreme vertices lie exactly on the circle, so for the
  only use (touching an axis-aligned wall segment away from the corners) the
  two agree; near a corner they differ by at most r(1-cos(pi/64)) = 3.6e-4 m.
* ``LineString([a, b]).intersection(disc).is_empty`` == distance(segment, centre) > r.
* ``box`` / ``affinity.translate`` / ``affinity.rotate(use_radians=True)`` keep
  the 4 vertices in fp64 with shapely's own affine arithmetic (including its
  snap of |cos|,|sin| < 2.5e-16 to 0).
* ``a.intersects(b)`` between two (possibly degenerate) rectangles is the
  closed separating-axis test over the edge normals of both; touching counts.
  A zero-length rectangle (agent speed 0) is treated as the closed segment of
  width 2r it degenerates to.
"""
import math
import sys
import types


class _Empty:
    is_empty = True


class _NonEmpty:
    is_empty = False


class Disc:
    def __init__(self, x, y, r):
        self.x, self.y, self.r = float(x), float(y), float(r)

    def simplify(self, tol):
        return self

    def intersects(self, other):
        if isinstance(other, Poly):
            return other.intersects_disc(self)
        raise NotImplementedError


class Point:
    def __init__(self, x, y=None):
        if y is None:
            x, y = x
        self.x, self.y = float(x), float(y)

    def buffer(self, r):
        return Disc(self.x, self.y, r)


def _seg_point_dist2(ax, ay, bx, by, px, py):
    dx, dy = bx - ax, by - ay
    l2 = dx * dx + dy * dy
    if l2 == 0.0:
        t = 0.0
    else:
        t = ((px - ax) * dx + (py - ay) * dy) / l2
        t = min(1.0, max(0.0, t))
    cx, cy = ax + t * dx, ay + t * dy
    return (px - cx) ** 2 + (py - cy) ** 2


class LineString:
    def __init__(self, pts):
        self.pts = [(float(p[0]), float(p[1])) for p in pts]

    def intersection(self, other):
        if isinstance(other, Disc):
            (ax, ay), (bx, by) = self.pts[0], self.pts[1]
            d2 = _seg_point_dist2(ax, ay, bx, by, other.x, other.y)
            return _NonEmpty() if d2 <= other.r * other.r else _Empty()
        raise NotImplementedError


class Poly:
    """Convex polygon given by its vertex list (rectangles only in practice)."""

    def __init__(self, pts):
        self.pts = [(float(x), float(y)) for x, y in pts]

    def simplify(self, tol):
        return self

    def _axes(self):
        n = len(self.pts)
        for i in range(n):
            x0, y0 = self.pts[i]
            x1, y1 = self.pts[(i + 1) % n]
            yield (-(y1 - y0), x1 - x0)

    @staticmethod
    def _proj(pts, ax):
        vals = [p[0] * ax[0] + p[1] * ax[1] for p in pts]
        return min(vals), max(vals)

    def intersects(self, other):
        if isinstance(other, Disc):
            return self.intersects_disc(other)
        for ax in list(self._axes()) + list(other._axes()):
            a0, a1 = self._proj(self.pts, ax)
            b0, b1 = self._proj(other.pts, ax)
            if a1 < b0 or b1 < a0:
                return False
        return True

    def intersects_disc(self, disc):
        raise NotImplementedError("norm zones are out of scope (reward.norm_zones=False)")


def box(minx, miny, maxx, maxy):
    # shapely.geometry.box default ccw=True vertex order
    return Poly([(maxx, miny), (maxx, maxy), (minx, maxy), (minx, miny)])


def Polygon(pts):
    return Poly(list(pts))


def translate(geom, xoff=0.0, yoff=0.0, zoff=0.0):
    return Poly([(1.0 * x + 0.0 * y + xoff, 0.0 * x + 1.0 * y + yoff) for x, y in geom.pts])


def rotate(geom, angle, origin="center", use_radians=False):
    if not use_radians:
        angle = angle * math.pi / 180.0
    cosp, sinp = math.cos(angle), math.sin(angle)
    if abs(cosp) < 2.5e-16:
        cosp = 0.0
    if abs(sinp) < 2.5e-16:
        sinp = 0.0
    if origin == "center":
        xs = [p[0] for p in geom.pts]
        ys = [p[1] for p in geom.pts]
        x0, y0 = (min(xs) + max(xs)) / 2.0, (min(ys) + max(ys)) / 2.0
    elif origin == "centroid":
        raise NotImplementedError
    else:
        x0, y0 = float(origin[0]), float(origin[1])
    xoff = x0 - x0 * cosp + y0 * sinp
    yoff = y0 - x0 * sinp - y0 * cosp
    return Poly([(cosp * x + (-sinp) * y + xoff, sinp * x + cosp * y + yoff) for x, y in geom.pts])


def install():
    """Register `shapely`, `shapely.geometry`, `shapely.affinity` in sys.modules."""
    shp = types.ModuleType("shapely")
    geo = types.ModuleType("shapely.geometry")
    aff = types.ModuleType("shapely.affinity")
    geo.Point, geo.LineString, geo.box, geo.Polygon = Point, LineString, box, Polygon
    aff.translate, aff.rotate = translate, rotate
    shp.geometry, shp.affinity = geo, aff
    sys.modules["shapely"] = shp
    sys.modules["shapely.geometry"] = geo
    sys.modules["shapely.affinity"] = aff
