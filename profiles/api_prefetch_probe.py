"""Steady-state throughput of AnnoRegionRndSampler.torch_generator (slide resident) against the prefetch-group size, with the
bench's consumer loop (labels + coords read back every step). Not a product path.
    python profiles/api_prefetch_probe.py"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from deephisto_b200.patch_samplers.region_samplers import AnnoRegionRndSampler  # noqa: E402
from deephisto_b200.slide import SyntheticSlide  # noqa: E402
from deephisto_b200.synthetic import synth_polygons  # noqa: E402

H = W = 32768
B, K = 256, 2560
src = SyntheticSlide(H, W, seed=0)
polys = synth_polygons(50, H, W, seed=0)
h_labels = torch.empty(B, dtype=torch.int64).pin_memory()
h_coords = torch.empty((B, 2), dtype=torch.float32).pin_memory()
for batches, gb in ((16, 2.5), (32, 5), (64, 10), (16, 2.5), (32, 5)):
    api = AnnoRegionRndSampler([(src, polys)], layer=1, patch_size=224, patches_from_one_region=4, one_image_for_batch=True, seed=1,
                               verbose=False, prefetch_bytes=int(gb * (1 << 30)) + (1 << 28), prefetch_batches=batches)
    for mode in ("both", "none"):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for f, l, c in api.torch_generator(batch_size=B, n_batches=K, batches_per_worker=2):
                if mode == "both":
                    h_labels.copy_(l, non_blocking=True)
                    h_coords.copy_(c, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        print(f"prefetch {batches:3d} batches, consumer reads {mode:5s}: {K * B / dt / 1e6:.3f} M patches/s", flush=True)
    del api
    torch.cuda.empty_cache()
