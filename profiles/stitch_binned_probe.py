"""ncu target: one dh_stitch_binned call (sum map) on the 40k x 40k coverage-sampler coordinate list.
    python profiles/stitch_binned_probe.py [d] [tile_rows]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from deephisto_b200 import _lib, ops  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 4
th = int(sys.argv[2]) if len(sys.argv) > 2 else 0
what = sys.argv[3] if len(sys.argv) > 3 else "sum"
H = W = 40000
st = ops.CoverState(H, W, 224, 16, 2, 1024, seed=0)
parts = []
for _ in range(128):
    c, nz = st.next_coords()
    parts.append(c.clone())
coords = torch.cat(parts)
logits = torch.randn((coords.shape[0], 5), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
_lib.require_device().dh_stitch_binned_set_tile_rows(th)
for _ in range(2):
    out = ops.stitch_binned(logits, coords, 224, d, H // d, W // d, want_sum=what == "sum", want_argmax=what != "sum")
    torch.cuda.synchronize()
    del out
print("ok", coords.shape[0])
