"""Stub of shapely.Polygon (GEOS is absent): bounds / area / intersection(square).area / is_valid / buffer,
backed by the oracle's float64 clip area. Geometry parity with real GEOS is UNPINNED (oracle/region.py)."""
import numpy as np

from oracle import region as _r


class _Area:
    def __init__(self, a):
        self.area = a


class Polygon:
    def __init__(self, pts):
        self._v = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
        self._edges = None

    @property
    def is_valid(self):
        return True

    def buffer(self, d):
        return self

    @property
    def area(self):
        return _r.polygon_area(self._v)

    @property
    def bounds(self):
        return _r.polygon_bounds(self._v)

    def intersection(self, other):
        # `other` is always the axis-aligned patch square (region_samplers.py:125-133,180-188)
        x0, y0, x1, y1 = other.bounds
        if self._edges is None:
            self._edges = _r.build_edges(self._v)
        return _Area(float(_r.clip_area(self._edges, x0, y0, x1 - x0)[0]))
