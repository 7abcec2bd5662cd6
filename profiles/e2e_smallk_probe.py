"""Where does the first torch_generator call over a pinned slide spend its time for small K? (diagnostic, not a product path)"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from deephisto_b200.patch_samplers.region_samplers import AnnoRegionRndSampler  # noqa: E402
from deephisto_b200.slide import PinnedSlide, SyntheticSlide  # noqa: E402
from deephisto_b200.synthetic import synth_polygons  # noqa: E402

H = W = 32768
B = 256
src = SyntheticSlide(H, W, seed=0)
slide = src.device_slide("cuda")
polys = synth_polygons(50, H, W, seed=0)
host = PinnedSlide.from_device(slide)
h_labels = torch.empty(B, dtype=torch.int64).pin_memory()
h_coords = torch.empty((B, 2), dtype=torch.float32).pin_memory()


def run(api, K, tag):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    marks = []
    for i, (f, l, c) in enumerate(api.torch_generator(batch_size=B, n_batches=K, batches_per_worker=2)):
        h_labels.copy_(l, non_blocking=True)
        h_coords.copy_(c, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        if i % 32 == 0 or i == K - 1:
            marks.append((i, round((time.perf_counter() - t0) * 1e3, 1)))
    print(tag, "K =", K, "total ms", round((time.perf_counter() - t0) * 1e3, 1), marks, flush=True)


warm = AnnoRegionRndSampler([(src, polys)], layer=1, patch_size=224, patches_from_one_region=4, one_image_for_batch=True, seed=1, verbose=False)
run(warm, 32, "warm")
del warm, slide
src._dev.clear()
torch.cuda.empty_cache()
for K in (100, 100, 33, 300, 100):
    api = AnnoRegionRndSampler([(host, polys)], layer=1, patch_size=224, patches_from_one_region=4, one_image_for_batch=True, seed=7, verbose=False)
    run(api, K, "first ")
    run(api, K, "steady")
    del api
    torch.cuda.empty_cache()
