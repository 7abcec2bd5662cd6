"""Example of using FullImageRndSampler (reference: examples/sample_full_random.py).

    python -m deephisto_b200.examples.sample_full_random [--synthetic 8192 8192 | --image slide.npy]"""

import argparse
import time

from ..patch_samplers.full_samplers import FullImageRndSampler, SamplerExecutionMode
from ._common import slide_args, slide_source

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    slide_args(ap)
    ap.add_argument("--batch-size", type=int, default=16)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--quiet", action="store_true")
    args = ap.parse_args()
    patch_sampler = FullImageRndSampler(slide_source(args), layer=args.layer, patch_size=224, batch_size=args.batch_size,
                                        mode=SamplerExecutionMode.INMEMORY_SINGLEPROC, seed=args.seed)
    t0, n = time.time(), 0
    for inputs, coords, filled_ratio in patch_sampler.generator_torch():
        n += inputs.shape[0]
        if not args.quiet:
            print(inputs.shape, coords.shape, filled_ratio)
    print(f"{n / (time.time() - t0)} items/s, filled_ratio {filled_ratio}")
