"""Whole-slide stitching of per-patch logits (CPU oracle)."""

import numpy as np


def stitch(logits: np.ndarray, coords: np.ndarray, h: int, w: int, ps: int, d: int, row_begin: int = 0, row_end: int | None = None):
    """Restates ImagePredictorPatched.process (examples/predict_full_patched.py:40-63):
        prediction = zeros([h//d, w//d, n], float32)
        for each patch in sampler order: prediction[y//d:(y+ps)//d, x//d:(x+ps)//d, :] += logits[i]
        argmax(axis=2)
    plus the count map the reference has commented out (:45,55-58). Returns (sum f32 [dh,dw,n],
    count int64 [dh,dw], argmax int64 [dh,dw]) restricted to rows [row_begin, row_end)."""
    dh, dw = h // d, w // d
    n = logits.shape[1]
    pred = np.zeros([dh, dw, n], dtype=np.float32)
    count = np.zeros([dh, dw], dtype=np.int64)
    lg = np.asarray(logits, dtype=np.float32)
    for i, (y, x) in enumerate(np.asarray(coords).tolist()):
        pred[y // d : (y + ps) // d, x // d : (x + ps) // d, :] += lg[i]
        count[y // d : (y + ps) // d, x // d : (x + ps) // d] += 1
    row_end = dh if row_end is None else row_end
    pred, count = pred[row_begin:row_end], count[row_begin:row_end]
    return pred, count, np.argmax(pred, axis=2)


def normalize(sum_map: np.ndarray, count: np.ndarray) -> np.ndarray:
    """sum / max(count, 1) in float32 (the north-star's count normalisation; reference :61 is commented out)."""
    c = np.maximum(count, 1).astype(np.float32)
    return (sum_map / c[..., None]).astype(np.float32)
