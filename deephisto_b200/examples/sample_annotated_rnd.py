"""Random sampling inside annotated regions -- counterpart of the reference's examples/sample_annotated_rnd.py (its `--torch`
switch is kept: tensors on the device, or lists of (Patch, class index) structs).

    python -m deephisto_b200.examples.sample_annotated_rnd --torch [--synthetic 32768 32768 | --dataset folder --sample train]"""

import argparse

import numpy as np

from ..patch_samplers.region_samplers import AnnoRegionRndSampler
from ._common import Throughput, annotated_dataset, slide_args


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--torch", action="store_true", help="yield torch tensors instead of Patch structs")
    slide_args(ap, default_hw=(32768, 32768))
    ap.add_argument("--dataset", default=None, help="folder with images/<sample>/ and annotations/<sample>/")
    ap.add_argument("--sample", default="train")
    ap.add_argument("-n", type=int, default=40, help="batches to draw")
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--quiet", action="store_true", help="do not print one line per batch")
    opt = ap.parse_args(argv)

    sampler = AnnoRegionRndSampler(annotated_dataset(opt), layer=opt.layer, patch_size=224, patches_from_one_region=4, one_image_for_batch=True)
    per_class = np.zeros(len(sampler.classes), dtype=np.int64)
    meter = Throughput()
    if opt.torch:
        for feats, labels, origins in sampler.torch_generator(opt.batch_size, opt.n, batches_per_worker=2):
            meter.add(len(labels))
            per_class += np.bincount(labels.cpu().numpy(), minlength=len(per_class))
            if not opt.quiet:
                print(f"features {tuple(feats.shape)} labels {tuple(labels.shape)} origins {tuple(origins.shape)}", flush=True)
    else:
        for structs in sampler.structs_generator(opt.batch_size, opt.n, batches_per_worker=2):
            meter.add(len(structs))
            for _patch, cls in structs:
                per_class[cls] += 1
            if not opt.quiet:
                print(f"{len(structs)} Patch structs", flush=True)
    meter.report()
    print("patches per class:", dict(zip(sampler.classes, per_class.tolist())))


if __name__ == "__main__":
    main()
