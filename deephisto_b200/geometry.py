"""Host-side annotation geometry for the region samplers: runs once per dataset, float64 numpy.

Mirrors the set-up work of the reference (patch_samplers/region_samplers.py):
  RegionAnnotation.__init__      :28-73    vertices / layer, validity, area
  _parse_annotations             :194-249  JSON -> regions per class / per image
  _calc_area_weights/_calc_weights :339-482 sampling weights
and packs the result into flat device tables (struct dh_region_tables, include/deephisto_b200.h)
consumed by the sm_100a kernels. shapely's `buffer(0)` repair of invalid polygons is not available
(GEOS is not a dependency). A self-intersecting ring is, by default, KEPT with winding-number semantics -- the kernels' clip area is
|sum of signed edge integrals| = |sum over the ring's faces of winding x area|, which equals `buffer(0)`'s area up to the faces a
crossing creates (buffer(0) keeps the positively wound faces once; a small hand-drawn loop changes the result by the loop's own area) --
and named in a warning; `invalid="skip"` drops it instead (round-1 behaviour)."""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib

EDGE_STRIDE = 8


def polygon_area(v: np.ndarray) -> float:
    x, y = v[:, 0], v[:, 1]
    return float(abs(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)) * 0.5)


def polygon_bounds(v: np.ndarray) -> tuple[float, float, float, float]:
    return float(v[:, 0].min()), float(v[:, 1].min()), float(v[:, 0].max()), float(v[:, 1].max())


def build_edges(v: np.ndarray) -> np.ndarray:
    """[E,8] edge table (layout: include/deephisto_b200.h). Non-horizontal edges, oriented yA < yB."""
    p = np.ascontiguousarray(v, dtype=np.float64)
    q = np.roll(p, -1, axis=0)
    keep = p[:, 1] != q[:, 1]
    p, q = p[keep], q[keep]
    up = p[:, 1] < q[:, 1]
    a = np.where(up[:, None], p, q)
    b = np.where(up[:, None], q, p)
    dx, dy = b[:, 0] - a[:, 0], b[:, 1] - a[:, 1]
    m = dx / dy
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(dx == 0.0, 0.0, dy / np.where(dx == 0.0, 1.0, dx))
    out = np.zeros((len(a), EDGE_STRIDE), dtype=np.float64)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = a[:, 0], a[:, 1], b[:, 0], b[:, 1]
    out[:, 4], out[:, 5], out[:, 6] = m, r, np.where(up, 1.0, -1.0)
    return out


def _segments_intersect_properly(v: np.ndarray) -> bool:
    """True if two non-adjacent edges of the closed ring cross (shapely `not polygon.is_valid`, :69)."""
    n = len(v)
    if n < 4:
        return False
    a, b = v, np.roll(v, -1, axis=0)

    def orient(p, q, r):
        return np.sign((q[..., 0] - p[..., 0]) * (r[..., 1] - p[..., 1]) - (q[..., 1] - p[..., 1]) * (r[..., 0] - p[..., 0]))

    # all pairs (i, j >= i + 2) of non-adjacent edges, a block of first edges at a time by broadcasting: O(block * n) memory instead
    # of O(n^2) (a 10 000-vertex annotation has 5e7 pairs)
    def cross(px, py, qx, qy, rx, ry):
        return np.sign((qx - px) * (ry - py) - (qy - py) * (rx - px))

    ax, ay, bx, by = a[:, 0], a[:, 1], b[:, 0], b[:, 1]
    block = max(1, (1 << 20) // n)
    cols = np.arange(n)
    for i0 in range(0, n - 2, block):
        i1 = min(i0 + block, n - 2)
        pax, pay, pbx, pby = (t[i0:i1, None] for t in (ax, ay, bx, by))
        o1 = cross(pax, pay, pbx, pby, ax[None, :], ay[None, :])
        o2 = cross(pax, pay, pbx, pby, bx[None, :], by[None, :])
        o3 = cross(ax[None, :], ay[None, :], bx[None, :], by[None, :], pax, pay)
        o4 = cross(ax[None, :], ay[None, :], bx[None, :], by[None, :], pbx, pby)
        rows = np.arange(i0, i1)[:, None]
        ok = (cols[None, :] >= rows + 2) & ~((rows == 0) & (cols[None, :] == n - 1))
        if bool(np.any(ok & (o1 * o2 < 0) & (o3 * o4 < 0))):
            return True
    return False


def area_weights(areas, area_influence: float) -> np.ndarray:
    """region_samplers.py:339-378."""
    assert -1 <= area_influence <= 1
    a = np.asarray(list(areas), dtype=np.float64)
    w_default = np.ones(len(a), dtype=np.float64) / len(a)
    if area_influence == 0:
        return w_default
    if area_influence > 0:
        target = a / sum(a.tolist())
        f = area_influence
    else:
        inv = [1 / x for x in a.tolist()]
        target = np.array(inv) / sum(inv)
        f = -area_influence
    w = w_default + (target - w_default) * f
    # NB the reference writes `sum(w)` over a NUMPY array here and `sum(areas)` over a list of Python floats above: on the
    # Python 3.12 it pins (environment.yaml) the builtin sum is Neumaier-compensated for exact floats only, so the two sums round
    # differently. Both spellings are kept as they are (tests/golden/golden_weights_v1.json pins the result bit for bit).
    return w / sum(w)


@dataclass
class Region:
    """One annotated polygon on one image at the sampler's layer scale."""

    image: int
    region_idx: int
    class_: str
    vertices: np.ndarray  # float64 [V,2] (x, y), already divided by layer
    area: float
    bounds: tuple[float, float, float, float]
    edges: np.ndarray = field(repr=False, default=None)


class RegionTables:
    """Flat device tables for dh_region_sample / dh_region_accept_dense / dh_rasterize_polygons."""

    def __init__(self, regions: list[Region], classes: list[str], img_hw: list[tuple[int, int]], tables: list[dict[str, list[int]]],
                 table_weights: np.ndarray, area_influence: float, device="cuda", sorted_classes_per_table: bool = False):
        C_ = len(classes)
        self.classes, self.n_regions, self.n_tables = classes, len(regions), len(tables)
        edge_off = np.zeros(len(regions) + 1, dtype=np.int32)
        for i, r in enumerate(regions):
            edge_off[i + 1] = edge_off[i] + len(r.edges)
        edges = np.concatenate([r.edges for r in regions], axis=0) if regions else np.zeros((0, EDGE_STRIDE))
        tbl_cls_off, tbl_cls, cat_off, cat_region, cat_cdf = [0], [], [0], [], []
        for t in tables:
            present = list(range(C_)) if sorted_classes_per_table else [classes.index(c) for c in t.keys()]
            tbl_cls += present
            tbl_cls_off.append(len(tbl_cls))
            for c in classes:
                regs = t.get(c, [])
                if regs:
                    cdf = np.cumsum(area_weights([regions[r].area for r in regs], area_influence))
                    cdf[-1] = 1.0
                    cat_region += regs
                    cat_cdf += cdf.tolist()
                cat_off.append(len(cat_region))
        img_cdf = np.cumsum(np.asarray(table_weights, dtype=np.float64))
        img_cdf[-1] = 1.0
        self.host = dict(
            edges=edges.reshape(-1), edge_off=edge_off,
            reg_bbox=np.asarray([r.bounds for r in regions], dtype=np.float64).reshape(-1),
            reg_area=np.asarray([r.area for r in regions], dtype=np.float64),
            reg_image=np.asarray([r.image for r in regions], dtype=np.int32),
            img_hw=np.asarray(img_hw, dtype=np.int32).reshape(-1),
            tbl_cls_off=np.asarray(tbl_cls_off, dtype=np.int32), tbl_cls=np.asarray(tbl_cls, dtype=np.int32),
            cat_off=np.asarray(cat_off, dtype=np.int32), cat_region=np.asarray(cat_region, dtype=np.int32),
            cat_cdf=np.asarray(cat_cdf, dtype=np.float64), img_cdf=img_cdf,
        )
        self.dev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in self.host.items()}
        self.struct = _lib.RegionTables(
            **{k: (self.dev[k].data_ptr() if self.dev[k].numel() else None) for k in self.host},
            n_tables=len(tables), n_classes=C_, n_regions=len(regions), n_images=len(img_hw),
        )


def make_region(image: int, region_idx: int, class_: str, vertices: np.ndarray, layer: int, invalid: str = "approximate") -> Region:
    """RegionAnnotation.__init__ (region_samplers.py:64-73) without shapely. `invalid`: what to do with a self-intersecting ring, which
    the reference repairs with buffer(0) (:69-71): "approximate" keeps it with winding-number area semantics and warns, "skip" raises."""
    v = np.asarray(vertices)
    if v.ndim != 2 or v.shape[1] != 2:
        raise RuntimeError("Invalid region shape. It should be (N, 2).")
    if v.dtype != np.float64:
        raise RuntimeError("Invalid region dtype. It should be float64.")
    if len(v) < 3:
        raise RuntimeError("A polygon needs at least 3 vertices.")
    v = v if layer == 1 else v.copy() / layer
    if _segments_intersect_properly(v):
        if invalid == "skip":
            raise RuntimeError("invalid (self-intersecting) polygon: buffer(0) repair needs GEOS and is not supported")
        import warnings

        warnings.warn(f"image {image}, region {region_idx} (class {class_!r}): self-intersecting polygon kept with winding-number area semantics "
                      f"(signed area {polygon_area(v):.1f} px^2); shapely's buffer(0) repair may differ by the area of the crossing loops",
                      RuntimeWarning, stacklevel=3)
    return Region(image, region_idx, class_, v, polygon_area(v), polygon_bounds(v), build_edges(v))
