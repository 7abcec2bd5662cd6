"""JPEG patches per class for a test set -- counterpart of the reference's examples/extract_patches_for_test_set.py
(extract_and_save_subset, region_samplers.py:874-909).

    python -m deephisto_b200.examples.extract_patches_for_test_set --out patches_test [--synthetic 8192 8192 | --dataset folder --sample test]"""

import argparse
from pathlib import Path

from ..patch_samplers.region_samplers import extract_and_save_subset
from ._common import Throughput, annotated_dataset, slide_args


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    slide_args(ap, default_hw=(8192, 8192))
    ap.add_argument("--dataset", default=None)
    ap.add_argument("--sample", default="test")
    ap.add_argument("--out", default="patches_test")
    ap.add_argument("--patches-per-class", type=int, default=100)
    ap.add_argument("--polygons", type=int, default=10, help="synthetic polygons when no dataset is given")
    opt = ap.parse_args(argv)

    out_dir = Path(opt.out)
    meter = Throughput()
    extract_and_save_subset(img_anno_paths=annotated_dataset(opt, opt.polygons), out_folder=out_dir, patch_size=224, layer=opt.layer,
                            patches_per_class=opt.patches_per_class)
    n = sum(1 for _ in out_dir.rglob("*.jpg"))
    meter.add(n)
    print(f"Total patches: {n} in {out_dir}")
    meter.report()


if __name__ == "__main__":
    main()
