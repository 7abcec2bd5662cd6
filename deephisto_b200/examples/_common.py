"""Shared argument handling of the example entry points: the reference hard-codes private dataset paths
(examples/sample_full_dense.py:14-16, sample_annotated_rnd.py:26-28); here every script takes `--image x.npy` /
`--dataset folder` or `--synthetic H W` so that it runs without the dataset."""

from __future__ import annotations

import argparse
from pathlib import Path


def slide_args(ap: argparse.ArgumentParser, default_hw=(8192, 8192)):
    ap.add_argument("--image", default=None, help=".npy slide (uint8 [H,W,3]), or .psi when psimage is installed")
    ap.add_argument("--synthetic", type=int, nargs=2, metavar=("H", "W"), default=None, help=f"synthetic slide (default {default_hw[0]} {default_hw[1]})")
    ap.add_argument("--layer", type=int, default=1)
    ap.set_defaults(_default_hw=default_hw)


def slide_source(args):
    from ..slide import SyntheticSlide

    if args.image:
        return args.image
    h, w = args.synthetic or args._default_hw
    return SyntheticSlide(h, w, seed=0)


def get_img_ano_paths(ds_folder: Path, sample: str = "train"):
    """The reference's utils.get_img_ano_paths (utils.py:4-14): pairs images/<sample>/*.{psi,npy} with annotations/<sample>/<stem>.json."""
    ds_folder = Path(ds_folder)
    imgs = sorted(p for p in (ds_folder / "images" / sample).iterdir() if p.suffix in (".psi", ".npy"))
    return [(p, ds_folder / "annotations" / sample / f"{p.stem}.json") for p in imgs]


def annotated_dataset(args, n_polygons: int = 50):
    """[(slide source, annotations)] from --dataset, or one synthetic slide with `n_polygons` synthetic star polygons."""
    from ..slide import SyntheticSlide
    from ..synthetic import synth_polygons

    if getattr(args, "dataset", None):
        return get_img_ano_paths(Path(args.dataset), args.sample)
    h, w = args.synthetic or args._default_hw
    return [(SyntheticSlide(h, w, seed=0), synth_polygons(n_polygons, h, w, seed=0))]


class Throughput:
    """Wall-clock items/s counter for the example drivers (the reference prints `items/s` in its annotated examples)."""

    def __init__(self):
        import time

        self._clock = time.perf_counter
        self.start = self._clock()
        self.items = 0

    def add(self, n: int):
        self.items += int(n)

    def report(self, extra: str = ""):
        import torch

        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dt = self._clock() - self.start
        print(f"{self.items / dt} items/s{extra}")
