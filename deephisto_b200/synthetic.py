"""Synthetic benchmark inputs: annotation polygons in the reference's JSON schema
(patch_samplers/region_samplers.py:218-227). The slide itself is generated in HBM by dh_synth_slide."""

from __future__ import annotations

import numpy as np


def synth_polygons(n: int, H: int, W: int, seed: int = 0, n_classes: int = 5, rmin: float = 600.0, rmax: float = 3000.0,
                   vmin: int = 24, vmax: int = 64) -> list[dict]:
    """n star-shaped simple polygons: centres uniform in the slide, vmin..vmax float64 vertices at sorted random
    angles with radii in [0.55, 1] * r_out, r_out uniform in [rmin, rmax]; classes round-robin."""
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
