"""Worker of tests/test_multigpu.py (one process per GPU, NCCL): row-band sharded whole-slide prediction must equal the
single-GPU result bit for bit (sum map) when the per-patch logits do not depend on batch composition."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from deephisto_b200.anno.utils import AnnoDescription  # noqa: E402
from deephisto_b200.examples import predict_full_patched as pfp  # noqa: E402
from deephisto_b200.patch_samplers import full_samplers as fs  # noqa: E402
from deephisto_b200.slide import SyntheticSlide  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


class FixedLogits(pfp.DeviceBatchPredictor):
    def logits(self, features):
        f = features.float()
        return torch.stack([f[:, 0, 3, 7], f[:, 1, 100, 50], f[:, 2, 223, 223], f[:, 0, 0, 0] * 2, f[:, 1, 17, 200] - f[:, 2, 5, 5]], 1).contiguous()


anno = AnnoDescription.with_auto_colors([f"c{i}" for i in range(5)])
mode = fs.SamplerExecutionMode.INMEMORY_SINGLEPROC
torch.manual_seed(0)
model = pfp.get_model(5)
for (H, W, stride, d) in [(3000, 2100, 112, 16), (1777, 1300, 100, 4)]:
    lazy = fs.FullImageDenseSampler(SyntheticSlide(H, W, seed=5), 1, 224, 64, mode, stride=stride, device=dev, lazy_slide=True)
    out = pfp.ImagePredictorPatched(None, lazy, FixedLogits(model, dev), anno, layer=1, downscale=d, device=dev, cnn_batch=96).process_device(
        want_sum=True, rank=rank, world=world)
    assert lazy._slide_dev is None
    full_s = fs.FullImageDenseSampler(SyntheticSlide(H, W, seed=5), 1, 224, 64, mode, stride=stride, device=dev)
    full = pfp.ImagePredictorPatched(None, full_s, FixedLogits(model, dev), anno, layer=1, downscale=d, device=dev, cnn_batch=128).process_device(want_sum=True)
    assert out["argmax"].shape == (H // d, W // d)
    assert torch.equal(out["sum"], full["sum"]), f"rank {rank}: banded sum map differs"
    assert torch.equal(out["argmax"], full["argmax"]), f"rank {rank}: banded class map differs"
    # the real CNN through the public process(rank, world): class map equal to the single-GPU map except where fp32 convolution
    # noise (different batch composition) flips a near-tie
    pred = pfp.DeviceBatchPredictor(model, dev)
    a = pfp.ImagePredictorPatched(None, lazy, pred, anno, layer=1, downscale=d, device=dev, cnn_batch=64).process(rank=rank, world=world)
    b = pfp.ImagePredictorPatched(None, full_s, pred, anno, layer=1, downscale=d, device=dev, cnn_batch=64).process()
    assert a.shape == b.shape and (a != b).mean() < 0.01, (a != b).mean()
# the reference's DEFAULT predict sampler (coverage-driven random sampling, predict_full_patched.py:156-163) through process(rank, world):
# per-rank band samplers, all-gather of the (coords, logits) lists, binned stitch per band, all-gather of the band maps. The assembled
# maps equal a single-GPU stitch of the concatenated list bit for bit, and every map cell is covered (d = speedup = 16).
from deephisto_b200 import ops  # noqa: E402

for (H, W, d) in [(2100, 1500, 16), (1333, 1000, 4)]:
    rs = fs.FullImageRndSampler(SyntheticSlide(H, W, seed=7), 1, 224, 16, mode, seed=3, device=dev, lazy_slide=True)
    ipp = pfp.ImagePredictorPatched(None, rs, FixedLogits(model, dev), anno, layer=1, downscale=d, device=dev, cnn_batch=96)
    out = ipp.process_device(want_sum=True, want_count=True, rank=rank, world=world)
    assert rs._slide_dev is None and out["argmax"].shape == (H // d, W // d) and sum(out["patches_per_rank"]) == len(out["coords"])
    assert min(out["patches_per_rank"]) > 0
    s1, c1, a1 = ops.stitch_binned(out["logits"], out["coords"], 224, d, H // d, W // d, want_count=True, want_argmax=True)
    assert torch.equal(out["sum"].view(torch.int32), s1.view(torch.int32)), f"rank {rank}: banded random-sampler sum map differs"
    assert torch.equal(out["argmax"], a1) and torch.equal(out["count"], c1)
    if d == 16:
        assert int(c1.min()) >= 1, "uncovered map cells"
    # FixedLogits are a function of the patch pixels: recompute them from the full slide at the gathered coordinates
    full = SyntheticSlide(H, W, seed=7).device_slide(dev)
    chk = FixedLogits(model, dev)
    idx = torch.arange(0, len(out["coords"]), max(1, len(out["coords"]) // 64), device=dev)
    again = chk.logits(chk.gather(full, out["coords"][idx].contiguous(), 224))
    assert torch.equal(again, out["logits"][idx]), f"rank {rank}: logits of the gathered list do not match the slide"
    b = pfp.ImagePredictorPatched(None, rs, FixedLogits(model, dev), anno, layer=1, downscale=d, device=dev).process(rank=rank, world=world)
    assert np.array_equal(b, a1.cpu().numpy().astype(np.int64))
# collective slide ingestion (slide.sharded_upload): 1/world of the rows per rank over PCIe, one NCCL all-gather; ragged shares

from deephisto_b200.slide import PinnedSlide, sharded_upload  # noqa: E402

for (H, W) in [(1000, 333), (world + 1, 40)]:
    a = np.random.default_rng(H).integers(0, 256, (H, W, 3), dtype=np.uint8)
    host = PinnedSlide.from_numpy(a)
    dev_slide, copied = sharded_upload(host, dev)
    assert np.array_equal(dev_slide.to_numpy(), a), f"rank {rank}: sharded upload differs from the host slide"
    t = torch.tensor([copied], device=dev)
    dist.all_reduce(t)
    assert int(t.item()) == host.nbytes
dist.barrier()
dist.destroy_process_group()
print(f"MULTIGPU OK rank {rank}/{world}")
