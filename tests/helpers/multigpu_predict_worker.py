"""Worker of tests/test_multigpu.py (one process per GPU, NCCL): row-band sharded whole-slide prediction must equal the
single-GPU result bit for bit (sum map) when the per-patch logits do not depend on batch composition."""
import os
import sys
from pathlib import Path

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
# collective slide ingestion (slide.sharded_upload): 1/world of the rows per rank over PCIe, one NCCL all-gather; ragged shares
import numpy as np  # noqa: E402

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
