"""Dense whole-slide sampling: coordinates and gather + normalise (CPU oracle)."""

import numpy as np


def dense_coords(h: int, w: int, ps: int, stride: int, batch_size: int):
    """Restates FullImageDenseSampler._create_batched_coords (patch_samplers/full_samplers.py:374-404):
    main grid (y outer, x inner) over range(0, h-ps, stride) x range(0, w-ps, stride), then the last
    column, the last row, the corner; chunks of batch_size, the last chunk padded with the corner.
    Returns (coords int32 [Npad, 2] (y, x), N)."""
    ys = list(range(0, h - ps, stride))
    xs = list(range(0, w - ps, stride))
    coords = [(y, x) for y in ys for x in xs]                   # :380-384
    coords += [(y, w - ps) for y in ys]                          # :386-389
    coords += [(h - ps, x) for x in xs]                          # :391-394
    coords.append((h - ps, w - ps))                              # :397
    n = len(coords)
    if batch_size > 0:
        while len(coords) % batch_size:                          # :400-402
            coords.append(coords[n - 1])
    return np.asarray(coords, dtype=np.int32).reshape(-1, 2), n


def gather(slide: np.ndarray, coords: np.ndarray, ps: int) -> np.ndarray:
    """uint8 [B, ps, ps, 3]: np.stack of data[y:y+ps, x:x+ps, :] (full_samplers.py:361-365,441-442).
    Pixels outside the slide read as 0 (the new build's documented out-of-bounds rule)."""
    H, W, _ = slide.shape
    out = np.zeros((len(coords), ps, ps, 3), dtype=np.uint8)
    for b, (y, x) in enumerate(np.asarray(coords).tolist()):
        y0, y1, x0, x1 = max(y, 0), min(y + ps, H), max(x, 0), min(x + ps, W)
        if y1 > y0 and x1 > x0:
            out[b, y0 - y : y1 - y, x0 - x : x1 - x] = slide[y0:y1, x0:x1]
    return out


def normalize(patches_u8: np.ndarray, scale255: bool = True, mean=None, std=None, layout: str = "NHWC",
              flip: np.ndarray | None = None) -> np.ndarray:
    """float32 features. scale255: `.astype(np.float32) / 255` (full_samplers.py:441-443; equals
    predict_full_patched.py:67-70 and region_samplers.py:616 bit for bit); without it
    plain float32 values 0..255 (FullImageRndSampler.generator_torch, full_samplers.py:286).
    mean/std: (v - mean[c]) / std[c] in float32 (torchvision Normalize semantics; not in the reference).
    flip: per-patch bits 1 = horizontal, 2 = vertical (train.py:71-81 applies them batch-wide).
    layout NCHW = permute(0,3,1,2).contiguous() (predict_full_patched.py:71)."""
    f = patches_u8.astype(np.float32)
    if scale255:
        f = f / np.float32(255)
    if mean is not None or std is not None:
        m = np.asarray(mean if mean is not None else (0, 0, 0), dtype=np.float32)
        s = np.asarray(std if std is not None else (1, 1, 1), dtype=np.float32)
        f = (f - m) / s
    if flip is not None:
        f = f.copy()
        for b, fl in enumerate(np.asarray(flip).tolist()):
            if fl & 1:
                f[b] = f[b][:, ::-1]
            if fl & 2:
                f[b] = f[b][::-1]
    if layout == "NCHW":
        f = np.ascontiguousarray(f.transpose(0, 3, 1, 2))
    return f
