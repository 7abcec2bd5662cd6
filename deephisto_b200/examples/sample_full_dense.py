"""Dense whole-slide sampling demo -- counterpart of the reference's examples/sample_full_dense.py.

    python -m deephisto_b200.examples.sample_full_dense [--synthetic 8192 8192 | --image slide.npy] [--stride 112]"""

import argparse

from ..patch_samplers.full_samplers import FullImageDenseSampler, SamplerExecutionMode
from ._common import Throughput, slide_args, slide_source


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    slide_args(ap)
    ap.add_argument("--stride", type=int, default=112)
    ap.add_argument("--batch-size", type=int, default=16)
    ap.add_argument("--quiet", action="store_true", help="do not print one line per batch")
    opt = ap.parse_args(argv)
    sampler = FullImageDenseSampler(slide_source(opt), opt.layer, 224, opt.batch_size, SamplerExecutionMode.INMEMORY_SINGLEPROC, stride=opt.stride)
    meter = Throughput()
    for batch, origins, progress in sampler.generator_torch():
        meter.add(batch.shape[0])
        if not opt.quiet:
            print(tuple(batch.shape), tuple(origins.shape), progress)
    meter.report()


if __name__ == "__main__":
    main()
