"""Prediction post-processing (CPU oracle, test infrastructure): restates perform_and_save_visualizations
(examples/predict_full_patched.py:81-113) minus the JPEG encoding.
  colorize   :89-95    colored_image[pred == anno.id] = anno.color
  thumbnail  :103-104  psim.get_region((0,0),(H,W), target_hw=(h,w)) -- psimage's resampling filter is unknown (the package is
                       not in the reference tree: PARITY UNPINNED); defined here, and in dh_colorize_overlay, as the integer area
                       average of the d x d block under every map cell, rounded half up
  overlay    :108-110  (img * alpha + colored_image * (1 - alpha)).astype(np.uint8), alpha = 0.6, float64"""

import numpy as np


def colorize(pred: np.ndarray, colors: dict[int, tuple[int, int, int]]) -> np.ndarray:
    h, w = pred.shape[:2]
    colored = np.zeros((h, w, 3), dtype=np.uint8)
    for cid, color in colors.items():
        colored[pred == cid] = color
    return colored


def thumbnail(slide: np.ndarray, dh: int, dw: int, d: int) -> np.ndarray:
    block = slide[: dh * d, : dw * d].reshape(dh, d, dw, d, 3).astype(np.uint64).sum(axis=(1, 3))
    return ((block + (d * d) // 2) // (d * d)).astype(np.uint8)


def overlay(img: np.ndarray, colored: np.ndarray, alpha: float = 0.6) -> np.ndarray:
    return (img * alpha + colored * (1 - alpha)).astype(np.uint8)
