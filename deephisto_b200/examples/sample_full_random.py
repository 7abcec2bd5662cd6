"""Coverage-driven random whole-slide sampling demo -- counterpart of the reference's examples/sample_full_random.py.

    python -m deephisto_b200.examples.sample_full_random [--synthetic 8192 8192 | --image slide.npy] [--seed 0]"""

import argparse

from ..patch_samplers.full_samplers import FullImageRndSampler, SamplerExecutionMode
from ._common import Throughput, slide_args, slide_source


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    slide_args(ap)
    ap.add_argument("--batch-size", type=int, default=16)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--quiet", action="store_true", help="do not print one line per batch")
    opt = ap.parse_args(argv)
    sampler = FullImageRndSampler(slide_source(opt), opt.layer, 224, opt.batch_size, SamplerExecutionMode.INMEMORY_SINGLEPROC, seed=opt.seed)
    meter = Throughput()
    covered = 0.0
    for batch, origins, covered in sampler.generator_torch():
        meter.add(batch.shape[0])
        if not opt.quiet:
            print(tuple(batch.shape), tuple(origins.shape), covered)
    meter.report(f", filled_ratio {covered}")


if __name__ == "__main__":
    main()
