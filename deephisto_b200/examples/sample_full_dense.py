"""Example of using FullImageDenseSampler (reference: examples/sample_full_dense.py).

    python -m deephisto_b200.examples.sample_full_dense [--synthetic 8192 8192 | --image slide.npy]"""

import argparse
import time

from ..patch_samplers.full_samplers import FullImageDenseSampler, SamplerExecutionMode
from ._common import slide_args, slide_source

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    slide_args(ap)
    ap.add_argument("--stride", type=int, default=112)
    ap.add_argument("--batch-size", type=int, default=16)
    ap.add_argument("--quiet", action="store_true")
    args = ap.parse_args()
    patch_sampler = FullImageDenseSampler(slide_source(args), layer=args.layer, patch_size=224, batch_size=args.batch_size, stride=args.stride,
                                          mode=SamplerExecutionMode.INMEMORY_SINGLEPROC)
    t0, n = time.time(), 0
    for inputs, coords, filled_ratio in patch_sampler.generator_torch():
        n += inputs.shape[0]
        if not args.quiet:
            print(inputs.shape, coords.shape, filled_ratio)
    import torch

    torch.cuda.synchronize()
    print(f"{n / (time.time() - t0)} items/s")
