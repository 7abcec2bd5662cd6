"""Dense grid inside every annotated region -- counterpart of the reference's examples/sample_annotated_dense.py.

    python -m deephisto_b200.examples.sample_annotated_dense [--synthetic 32768 32768 | --dataset folder --sample test]"""

import argparse

import numpy as np

from ..patch_samplers.region_samplers import AnnoRegionDenseSampler
from ._common import Throughput, annotated_dataset, slide_args


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    slide_args(ap, default_hw=(32768, 32768))
    ap.add_argument("--dataset", default=None)
    ap.add_argument("--sample", default="test")
    ap.add_argument("--stride", type=int, default=112)
    ap.add_argument("--polygons", type=int, default=8, help="synthetic polygons when no dataset is given")
    opt = ap.parse_args(argv)

    sampler = AnnoRegionDenseSampler(annotated_dataset(opt, opt.polygons), layer=opt.layer, patch_size=224, stride=opt.stride)
    per_class = np.zeros(len(sampler.classes), dtype=np.int64)
    meter = Throughput()
    for _patch, cls in sampler.structs_generator():
        per_class[cls] += 1
        meter.add(1)
    print(f"Total patches: {int(per_class.sum())}")
    meter.report()
    print("patches per class:", dict(zip(sampler.classes, per_class.tolist())))


if __name__ == "__main__":
    main()
