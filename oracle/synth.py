"""Synthetic inputs shared by tests and the benchmark's CPU legs.

synth_slide restates deephisto_b200/csrc/dh_dense.cu::synth_slide_kernel byte for byte, so a slide
generated on the device (no host copy) and one generated here are identical."""

import numpy as np


def _fmix32(h):
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x85EBCA6B)
    h ^= h >> np.uint32(13)
    h *= np.uint32(0xC2B2AE35)
    h ^= h >> np.uint32(16)
    return h


def synth_words(first_word: int, count: int, seed: int = 0) -> np.ndarray:
    """uint32 words [first_word, first_word+count) of the stream."""
    idx = np.arange(first_word, first_word + count, dtype=np.uint64)
    a = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32) ^ np.uint32(seed & 0xFFFFFFFF)
    b = (idx >> np.uint64(32)).astype(np.uint32) ^ np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        return _fmix32(_fmix32(a) + b * np.uint32(0x9E3779B9))


def synth_slide(H: int, W: int, seed: int = 0, y0: int = 0, rows: int | None = None) -> np.ndarray:
    """uint8 [rows, W, 3]: rows [y0, y0+rows) of the H x W synthetic slide."""
    rows = H - y0 if rows is None else rows
    k0 = y0 * 3 * W
    k1 = (y0 + rows) * 3 * W
    w0, w1 = k0 // 4, (k1 + 3) // 4
    words = synth_words(w0, w1 - w0, seed)
    b = words.view(np.uint8) if words.dtype.byteorder != ">" else words.byteswap().view(np.uint8)
    return b[k0 - 4 * w0 : k0 - 4 * w0 + (k1 - k0)].reshape(rows, W, 3).copy()


def synth_polygons(n: int, H: int, W: int, seed: int = 0, n_classes: int = 5, rmin: float = 600.0, rmax: float = 3000.0,
                   vmin: int = 24, vmax: int = 64):
    """n star-shaped simple polygons, float64 vertices (x, y), classes round-robin.
    Returns a list of {"class": str, "vertices": [[x, y], ...]} in the reference's JSON schema
    (region_samplers.py:218-227)."""
    rng = np.random.default_rng(seed)
    names = ["AT", "BG", "LP", "MM", "TUM", "DYS", "C6", "C7"][:n_classes]
    out = []
    for i in range(n):
        nv = int(rng.integers(vmin, vmax + 1))
        r_out = float(rng.uniform(rmin, rmax))
        r_out = min(r_out, min(H, W) / 2 - 2)
        cx = float(rng.uniform(r_out + 1, W - r_out - 1))
        cy = float(rng.uniform(r_out + 1, H - r_out - 1))
        ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
        rad = rng.uniform(0.55, 1.0, nv) * r_out
        verts = np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], axis=1)
        out.append({"class": names[i % n_classes], "vertices": verts.tolist()})
    return out
