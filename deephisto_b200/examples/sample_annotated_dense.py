"""Example of using AnnoRegionDenseSampler (reference: examples/sample_annotated_dense.py).

    python -m deephisto_b200.examples.sample_annotated_dense [--synthetic 32768 32768 | --dataset folder --sample test]"""

import argparse
import time

import numpy as np

from ..patch_samplers.region_samplers import AnnoRegionDenseSampler
from ._common import annotated_dataset, slide_args

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    slide_args(parser, default_hw=(32768, 32768))
    parser.add_argument("--dataset", default=None)
    parser.add_argument("--sample", default="test")
    parser.add_argument("--stride", type=int, default=112)
    parser.add_argument("--polygons", type=int, default=8)
    args = parser.parse_args()

    dataset = AnnoRegionDenseSampler(annotated_dataset(args, args.polygons), patch_size=224, stride=args.stride, layer=args.layer)
    t0 = time.time()
    count = np.zeros([len(dataset.classes)], dtype=np.int32)
    print("Generating batches of structs")
    for i, (patch, cls) in enumerate(dataset.structs_generator()):
        count[cls] += 1
    t1 = time.time()
    print(f"Total patches: {np.sum(count)}")
    print(f"{np.sum(count) / (t1 - t0)} items/s")
    print(f"patches extracted for classes: {count}")
