"""Annotated-region sampling (CPU oracle): polygon clip area, acceptance, weights, random sampling.

PARITY UNPINNED for the geometry: the reference's acceptance test is shapely/GEOS
`polygon.intersection(square).area > ps*ps*ri` (patch_samplers/region_samplers.py:125-134,180-189) and
neither shapely nor GEOS is part of the reference tree (environment.yaml pins only python). The
area of (simple polygon ∩ axis-aligned square) is unique, so it is restated from its definition in
two independent ways that must agree to ~1e-9 relative:
  clip_area      boundary integral, the exact operation order of dh_region.cu (bit-exact target)
  clip_area_sh   Sutherland-Hodgman clip against the four half-planes + shoelace (cross-check)
Everything around the geometry (bounds, rounding, loop order, weights) follows the reference lines cited."""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from .philox import STREAM_ATTEMPT, STREAM_GROUP, STREAM_TABLE, bounded, philox4x32_10, u01

_LIB = None


def _clib():
    global _LIB
    if _LIB is None:
        p = Path(__file__).resolve().parent / "_build" / "liboracle.so"
        if p.exists():
            lib = C.CDLL(str(p))
            lib.oracle_clip_area.restype = C.c_double
            lib.oracle_clip_area.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double]
            lib.oracle_clip_area_many.restype = None
            lib.oracle_clip_area_many.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
            _LIB = lib
        else:
            _LIB = False
    return _LIB


# ---- polygon primitives --------------------------------------------------------------------------
def scale_vertices(vertices: np.ndarray, layer: int) -> np.ndarray:
    """region_samplers.py:64-68: (N,2) float64, divided by `layer` when layer != 1."""
    v = np.asarray(vertices)
    if v.ndim != 2 or v.shape[1] != 2:
        raise RuntimeError("Invalid region shape. It should be (N, 2).")
    if v.dtype != np.float64:
        raise RuntimeError("Invalid region dtype. It should be float64.")
    return v if layer == 1 else v.copy() / layer


def polygon_area(v: np.ndarray) -> float:
    """shapely Polygon.area (region_samplers.py:73) == |shoelace|."""
    x, y = v[:, 0], v[:, 1]
    return float(abs(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)) * 0.5)


def polygon_bounds(v: np.ndarray):
    """shapely Polygon.bounds (region_samplers.py:116,173): (minx, miny, maxx, maxy)."""
    return float(v[:, 0].min()), float(v[:, 1].min()), float(v[:, 0].max()), float(v[:, 1].max())


def build_edges(v: np.ndarray) -> np.ndarray:
    """Edge table [E,8] = xA,yA,xB,yB (yA<yB), m=(xB-xA)/(yB-yA), r=(yB-yA)/(xB-xA) or 0, sgn, 0; horizontal edges dropped.
    (Layout defined in include/deephisto_b200.h; one IEEE division each for m and r.)"""
    p = np.asarray(v, dtype=np.float64)
    q = np.roll(p, -1, axis=0)
    rows = []
    for (x1, y1), (x2, y2) in zip(p.tolist(), q.tolist()):
        if y1 == y2:
            continue
        if y1 < y2:
            xA, yA, xB, yB, sgn = x1, y1, x2, y2, 1.0
        else:
            xA, yA, xB, yB, sgn = x2, y2, x1, y1, -1.0
        m = (xB - xA) / (yB - yA)
        r = 0.0 if xB == xA else (yB - yA) / (xB - xA)
        rows.append([xA, yA, xB, yB, m, r, sgn, 0.0])
    return np.asarray(rows, dtype=np.float64).reshape(-1, 8)


def clip_area(edges: np.ndarray, x, y, ps: float) -> np.ndarray:
    """float64 area(polygon ∩ [x,x+ps]x[y,y+ps]) for arrays of candidates; same operation order as
    dh_region.cu::edge_term (edges accumulated sequentially, no FMA)."""
    xs_in = np.atleast_1d(np.asarray(x, dtype=np.float64))
    ys_in = np.atleast_1d(np.asarray(y, dtype=np.float64))
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    lib = _clib()
    if lib:
        out = np.empty(len(xs_in), dtype=np.float64)
        xs_c, ys_c = np.ascontiguousarray(xs_in), np.ascontiguousarray(ys_in)
        lib.oracle_clip_area_many(edges.ctypes.data, len(edges), xs_c.ctypes.data, ys_c.ctypes.data, len(xs_c), float(ps), out.ctypes.data)
        return out
    return clip_area_numpy(edges, xs_in, ys_in, ps)


def clip_area_numpy(edges: np.ndarray, xs_in: np.ndarray, ys_in: np.ndarray, ps: float) -> np.ndarray:
    xa, ya = xs_in, ys_in
    xb, yb = xa + float(ps), ya + float(ps)
    acc = np.zeros_like(xa)
    wx = xb - xa
    with np.errstate(invalid="ignore"):
        for xA, yA, xB, yB, m, r, sgn, _ in edges.tolist():
            ys = np.maximum(yA, ya)
            ye = np.minimum(yB, yb)
            live = ys < ye
            xs = np.where(ys == yA, xA, xA + (ys - yA) * m)
            xe = np.where(ye == yB, xB, xA + (ye - yA) * m)
            g = np.minimum(np.maximum(xs, xa), xb) - xa
            v_vert = (ye - ys) * g
            xmin, xmax = np.minimum(xs, xe), np.maximum(xs, xe)
            ar = abs(r)
            cl, ch = np.maximum(xmin, xa), np.minimum(xmax, xb)
            v = np.where(cl < ch, ((ch - cl) * ar) * (((cl - xa) + (ch - xa)) * 0.5), 0.0)
            ul = np.maximum(xmin, xb)
            v = np.where(ul < xmax, v + ((xmax - ul) * ar) * wx, v)
            val = np.where(xs == xe, v_vert, v)
            acc = acc + np.where(live, sgn * val, 0.0)
    return np.abs(acc)


def clip_area_sh(v: np.ndarray, x: float, y: float, ps: float) -> float:
    """Independent check: Sutherland-Hodgman clip of the polygon against the square, then |shoelace|.
    Exact for simple polygons up to float64 rounding (degenerate bridge edges of concave clips cancel)."""
    pts = [tuple(p) for p in np.asarray(v, dtype=np.float64).tolist()]

    def clip(pts, inside, inter):
        out = []
        for i in range(len(pts)):
            a, b = pts[i - 1], pts[i]
            ia, ib = inside(a), inside(b)
            if ib:
                if not ia:
                    out.append(inter(a, b))
                out.append(b)
            elif ia:
                out.append(inter(a, b))
        return out

    def ix(c):
        return lambda a, b: (c, a[1] + (b[1] - a[1]) * (c - a[0]) / (b[0] - a[0]))

    def iy(c):
        return lambda a, b: (a[0] + (b[0] - a[0]) * (c - a[1]) / (b[1] - a[1]), c)

    for inside, inter in (
        (lambda p: p[0] >= x, ix(x)),
        (lambda p: p[0] <= x + ps, ix(x + ps)),
        (lambda p: p[1] >= y, iy(y)),
        (lambda p: p[1] <= y + ps, iy(y + ps)),
    ):
        if not pts:
            return 0.0
        pts = clip(pts, inside, inter)
    if len(pts) < 3:
        return 0.0
    a = np.asarray(pts)
    return float(abs(np.sum(a[:, 0] * np.roll(a[:, 1], -1) - np.roll(a[:, 0], -1) * a[:, 1])) * 0.5)


# ---- D: dense coordinates ---------------------------------------------------------------------------
def dense_candidates(v: np.ndarray, layer_size, ps: int, stride: int):
    """Candidate grid of _extract_patch_coords_dense (region_samplers.py:171-179): bounds rounded with
    Python's round (banker's), x1/y1 clamped to w-ps / h-ps, range(y0,y1,stride) x range(x0,x1,stride)."""
    h, w = layer_size
    x0, y0, x1, y1 = polygon_bounds(v)
    x0, y0, x1, y1 = round(x0), round(y0), round(x1), round(y1)
    x1 = min(x1, w - ps)
    y1 = min(y1, h - ps)
    ny = len(range(y0, y1, stride))
    nx = len(range(x0, x1, stride))
    return y0, x0, ny, nx


def coords_dense(v: np.ndarray, layer_size, ps: int, stride: int, ri: float = 0.75):
    """RegionAnnotation._extract_patch_coords_dense (region_samplers.py:145-191): accepted (y, x) row-major.
    Returns (coords int32 [n,2], mask uint8 [ny*nx], areas float64 [ny*nx])."""
    y0, x0, ny, nx = dense_candidates(v, layer_size, ps, stride)
    edges = build_edges(v)
    if ny == 0 or nx == 0:
        return np.zeros((0, 2), np.int32), np.zeros(0, np.uint8), np.zeros(0)
    yy, xx = np.meshgrid(y0 + stride * np.arange(ny), x0 + stride * np.arange(nx), indexing="ij")
    areas = clip_area(edges, xx.reshape(-1), yy.reshape(-1), ps)
    mask = areas > ps * ps * ri                                           # :189 strict
    coords = np.stack([yy.reshape(-1)[mask], xx.reshape(-1)[mask]], axis=1).astype(np.int32)
    return coords, mask.astype(np.uint8), areas


# ---- F: weights ----------------------------------------------------------------------------------------
def area_weights(areas, area_influence: float) -> np.ndarray:
    """AnnoRegionRndSampler._calc_area_weights (region_samplers.py:339-378), operation for operation."""
    assert -1 <= area_influence <= 1
    areas = list(areas)
    areas_inv = [1 / a for a in areas]
    w_proportional = np.array(areas) / sum(areas)
    w_inv_proportional = np.array(areas_inv) / sum(areas_inv)
    w_default = np.ones(len(areas), dtype=np.float64) / len(areas)
    if area_influence == 0:
        w = w_default
    elif area_influence > 0:
        delta = (w_proportional - w_default) * area_influence
        w = w_default + delta
        w = w / sum(w)
    else:
        delta = (w_inv_proportional - w_default) * (-area_influence)
        w = w_default + delta
        w = w / sum(w)
    return w


class RegionSet:
    """Parsed annotations of a dataset (restates _parse_annotations :194-249 and _calc_weights :395-482).
    `images` is a list of (layer_hw, [ {"class","vertices"} ... ])."""

    def __init__(self, images, layer: int, area_influence: float = 0.5, classes=None, one_image_for_batch: bool = False):
        self.verts, self.edges, self.bbox, self.area, self.reg_image, self.reg_class_name = [], [], [], [], [], []
        self.img_hw = [tuple(hw) for hw, _ in images]
        per_image = [dict() for _ in images]
        all_regions: dict[str, list[int]] = {}
        for j, (hw, annos) in enumerate(images):
            for a in annos:
                cls = a["class"]
                if classes is not None and cls not in classes:
                    continue
                v = scale_vertices(np.array(a["vertices"], dtype=np.float64), layer)
                rid = len(self.verts)
                self.verts.append(v)
                self.edges.append(build_edges(v))
                self.bbox.append(polygon_bounds(v))
                self.area.append(polygon_area(v))
                self.reg_image.append(j)
                self.reg_class_name.append(cls)
                per_image[j].setdefault(cls, []).append(rid)
                all_regions.setdefault(cls, []).append(rid)
        self.classes = sorted(all_regions.keys())                          # :306
        self.one_image = one_image_for_batch
        C_ = len(self.classes)
        if one_image_for_batch:
            tables = per_image
            img_areas = [sum(sum(self.area[r] for r in regs) for regs in t.values()) for t in per_image]   # :469-472
            self.img_w = area_weights(img_areas, area_influence)            # :473-475
        else:
            tables = [all_regions]
            self.img_w = np.ones(1)
        self.n_tables = len(tables)
        if one_image_for_batch:
            self.tbl_cls = [[self.classes.index(c) for c in t.keys()] for t in tables]   # :550-551 (dict order)
        else:
            self.tbl_cls = [list(range(C_))]                                             # :576 randint(len(classes))
        self.cat_regions = [[list(t.get(c, [])) for c in self.classes] for t in tables]
        self.cat_cdf = [
            [np.cumsum(area_weights([self.area[r] for r in regs], area_influence)) if regs else np.zeros(0) for regs in row]
            for row in self.cat_regions
        ]
        for row in self.cat_cdf:
            for cdf in row:
                if len(cdf):
                    cdf[-1] = 1.0
        self.img_cdf = np.cumsum(self.img_w)
        self.img_cdf[-1] = 1.0


def _cdf_pick(cdf: np.ndarray, u: float) -> int:
    """first i with cdf[i] > u, clamped (dh_region.cu::cdf_search)."""
    i = int(np.searchsorted(cdf, u, side="right"))
    return min(i, len(cdf) - 1)


def sample(rs: RegionSet, n_slots: int, k: int, ps: int, ri: float = 0.75, miss_limit: int = 500, max_redraw: int = 64,
           fixed_class: int = -1, slots_per_table_draw: int = 1, seed: int = 0, slot_offset: int = 0):
    """Restates dh_region.cu::region_sample_kernel, which itself follows AnnoRegionRndSampler._gen_single_proc
    (region_samplers.py:544-591) + RegionAnnotation._extract_patch_coords_rnd (:114-143): per group of k slots draw
    (image,) class, region; per slot the first attempt (of miss_limit) whose clip area > ps*ps*ri wins; any failure
    redraws the group's class and region. Returns (coords int32 [S,2], labels int64, images int32, status uint8)."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    thr = ps * ps * ri
    coords = np.zeros((n_slots, 2), np.int32)
    labels = np.full(n_slots, -1, np.int64)
    images = np.full(n_slots, -1, np.int32)
    status = np.full(n_slots, 1, np.uint8)
    for s0 in range(0, n_slots, k):
        kk = min(k, n_slots - s0)
        g0 = slot_offset + s0
        table = 0
        if rs.n_tables > 1:
            chunk = g0 // slots_per_table_draw
            pt = philox4x32_10(chunk & 0xFFFFFFFF, chunk >> 32, 0, STREAM_TABLE, k0, k1)
            table = _cdf_pick(rs.img_cdf, float(u01(pt[0])))
        cls_list = rs.tbl_cls[table]
        fail = 1
        done = False
        for rd in range(max_redraw):
            pg = philox4x32_10(g0 & 0xFFFFFFFF, g0 >> 32, rd, STREAM_GROUP, k0, k1)
            cls = fixed_class if fixed_class >= 0 else cls_list[int(bounded(pg[0], len(cls_list)))]
            regs = rs.cat_regions[table][cls]
            if not regs:
                fail = 2
                continue
            region = regs[_cdf_pick(rs.cat_cdf[table][cls], float(u01(pg[1])))]
            if rs.area[region] < thr:                                      # :117-118
                fail = 1
                continue
            bx0, by0, bx1, by1 = rs.bbox[region]
            img = rs.reg_image[region]
            h, w = rs.img_hw[img]
            xlo, ylo = int(bx0), int(by0)                                   # numpy randint truncates float bounds
            xhi = int(min(max(bx0 + 1, bx1 - ps), w))                       # :123
            yhi = int(min(max(by0 + 1, by1 - ps), h))                       # :124
            xhi, yhi = min(xhi, w - ps + 1), min(yhi, h - ps + 1)           # stay inside the slide
            xlo, ylo = max(xlo, 0), max(ylo, 0)
            if xhi <= xlo or yhi <= ylo:
                fail = 2
                continue
            ok_group = True
            got = []
            for s in range(kk):
                gs = g0 + s
                hit_yx = None
                for a0 in range(0, miss_limit, 32):                         # 32 attempts at a time, first accepted wins (kernel order)
                    att = np.arange(a0, min(a0 + 32, miss_limit), dtype=np.uint64)
                    pa = philox4x32_10(gs & 0xFFFFFFFF, gs >> 32, (np.uint64(rd) << np.uint64(16)) | att, STREAM_ATTEMPT, k0, k1)
                    xs = xlo + bounded(pa[0], xhi - xlo)
                    ys = ylo + bounded(pa[1], yhi - ylo)
                    ok = clip_area(rs.edges[region], xs, ys, ps) > thr      # :133-134 strict
                    hit = np.flatnonzero(ok)
                    if len(hit):
                        hit_yx = (int(ys[hit[0]]), int(xs[hit[0]]))
                        break
                if hit_yx is None:
                    ok_group = False
                    break
                got.append(hit_yx)
            if not ok_group:
                fail = 1
                continue
            for s, (y, x) in enumerate(got):
                coords[s0 + s] = (y, x)
                labels[s0 + s] = cls
                images[s0 + s] = img
                status[s0 + s] = 0
            done = True
            break
        if not done:
            status[s0 : s0 + kk] = fail
    return coords, labels, images, status


def rasterize(rs_edges, rs_bbox, scale: float, mh: int, mw: int) -> np.ndarray:
    """Pixel-centre even-odd rasterisation restating dh_region.cu::rasterize_kernel (own definition; the
    reference only rasterises for display through PIL, anno/utils.py:308-320)."""
    px = (np.arange(mw, dtype=np.float64) + 0.5) * scale
    py = (np.arange(mh, dtype=np.float64) + 0.5) * scale
    PX, PY = np.meshgrid(px, py)
    lab = np.zeros((mh, mw), np.int32)
    for r, (edges, bb) in enumerate(zip(rs_edges, rs_bbox)):
        inside_bb = (PX >= bb[0]) & (PX <= bb[2]) & (PY >= bb[1]) & (PY <= bb[3])
        cross = np.zeros((mh, mw), np.int64)
        for xA, yA, xB, yB, m, _, _, _ in np.asarray(edges).tolist():
            sel = (yA <= PY) & (PY < yB)
            xi = xA + (PY - yA) * m
            cross += (sel & (xi > PX)).astype(np.int64)
        lab = np.where(inside_bb & ((cross & 1) == 1), r + 1, lab)
    return lab
