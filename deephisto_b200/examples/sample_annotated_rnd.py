"""Example of using AnnoRegionRndSampler (reference: examples/sample_annotated_rnd.py; same flags plus the data source).

    python -m deephisto_b200.examples.sample_annotated_rnd --torch [--synthetic 32768 32768 | --dataset folder --sample train]"""

import argparse
import time

import numpy as np

from ..patch_samplers.region_samplers import AnnoRegionRndSampler
from ._common import annotated_dataset, slide_args

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--torch", action="store_true", help="if set, it will use torch tensor outputs")
    slide_args(parser, default_hw=(32768, 32768))
    parser.add_argument("--dataset", default=None, help="folder with images/<sample>/ and annotations/<sample>/ (utils.get_img_ano_paths)")
    parser.add_argument("--sample", default="train")
    parser.add_argument("-n", type=int, default=40, help="number of batches to extract")
    parser.add_argument("--batch-size", type=int, default=64)
    parser.add_argument("--quiet", action="store_true")
    args = parser.parse_args()

    n, b_size, b_per_worker = args.n, args.batch_size, 2
    dataset = AnnoRegionRndSampler(annotated_dataset(args), patch_size=224, layer=args.layer, patches_from_one_region=4, one_image_for_batch=True)
    t0 = time.time()
    count = np.zeros([len(dataset.classes)], dtype=np.int32)
    if args.torch:
        print("Generating batches with torch tensors")
        for f, cls, coords in dataset.torch_generator(batch_size=b_size, n_batches=n, batches_per_worker=b_per_worker):
            if not args.quiet:
                print(f"inputs: {f.shape}, cls: {cls.shape}, crds: {coords.shape}", flush=True)
            count += np.bincount(cls.cpu().numpy(), minlength=len(count)).astype(np.int32)
    else:
        print("Generating batches of structs")
        for batch in dataset.structs_generator(batch_size=b_size, n_batches=n, batches_per_worker=b_per_worker):
            if not args.quiet:
                print(f"batch of {len(batch)} patches with coords", flush=True)
            for patch, cls in batch:
                count[cls] += 1
    t1 = time.time()
    print(f"{n * b_size / (t1 - t0)} items/s")
    print(f"patches extracted for classes: {count}")
