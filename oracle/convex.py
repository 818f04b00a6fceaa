"""The handful of shapely operations ``MeshRenderer.crop_field_affine`` uses (feabas/renderer.py:436-449), for
convex polygons: ``shapely.geometry.box``, ``shapely.affinity.affine_transform``, ``.intersection``, ``.area``,
``shapely.contains_xy``.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  shapely is not installed in the build container; the region of
a rectangular mesh and the footprint of a block are convex, which is all the parity harness needs.
"""
import numpy as np


class ConvexPoly:
    """Convex polygon, counter-clockwise vertex list (possibly empty)."""

    def __init__(self, pts):
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
        if pts.shape[0] >= 3 and _signed_area(pts) < 0:
            pts = pts[::-1]
        self.pts = pts

    @property
    def area(self):
        return abs(_signed_area(self.pts)) if self.pts.shape[0] >= 3 else 0.0

    def intersection(self, other):
        """Sutherland-Hodgman clip of ``other`` against the edges of ``self``."""
        out = other.pts
        n = self.pts.shape[0]
        if n < 3:
            return ConvexPoly(np.empty((0, 2)))
        for i in range(n):
            if out.shape[0] == 0:
                break
            a, b = self.pts[i], self.pts[(i + 1) % n]
            edge = b - a
            side = edge[0] * (out[:, 1] - a[1]) - edge[1] * (out[:, 0] - a[0])      # >= 0: on the inner side
            nxt = np.roll(out, -1, axis=0)
            side_n = np.roll(side, -1)
            kept = []
            for p, q, sp, sq in zip(out, nxt, side, side_n):
                if sp >= 0:
                    kept.append(p)
                if (sp >= 0) != (sq >= 0):
                    kept.append(p + (q - p) * (sp / (sp - sq)))
            out = np.array(kept, dtype=np.float64).reshape(-1, 2)
        return ConvexPoly(out)


def _signed_area(pts):
    x, y = pts[:, 0], pts[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def box(xmin, ymin, xmax, ymax):
    return ConvexPoly([[xmin, ymin], [xmax, ymin], [xmax, ymax], [xmin, ymax]])


def affine_transform(poly, m):
    a, b, d, e, xoff, yoff = (float(v) for v in m)          # shapely: x' = a x + b y + xoff, y' = d x + e y + yoff
    x, y = poly.pts[:, 0], poly.pts[:, 1]
    return ConvexPoly(np.stack((a * x + b * y + xoff, d * x + e * y + yoff), axis=-1))


def contains_xy(poly, x, y):
    """shapely.contains_xy: strictly inside (boundary points are not contained)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    n = poly.pts.shape[0]
    if n < 3 or poly.area == 0:
        return np.zeros(x.shape, dtype=bool)
    inside = np.ones(x.shape, dtype=bool)
    for i in range(n):
        a, b = poly.pts[i], poly.pts[(i + 1) % n]
        inside &= ((b[0] - a[0]) * (y - a[1]) - (b[1] - a[1]) * (x - a[0])) > 0
    return inside


