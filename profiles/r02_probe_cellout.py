"""Class-map / count outputs of dh_stitch_binned: default dispatch (variant 0: cell-lane kernel with fused outputs where it applies)
vs the row-run (1) and segment (2) kernels. 40 000^2 coverage list. Not a product path."""
import sys

import os

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from deephisto_b200 import _lib, ops

lib = _lib.require_device()
PS, N = 224, 5


def timeit(fn, reps=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def cover_list(H, W, B=1024):
    st = ops.CoverState(H, W, PS, 16, 2, B, seed=0)
    cells = (H // 16) * (W // 16)
    parts = []
    while True:
        c, counts = st.next_group(16)
        parts.append(c.reshape(-1, 2))
        if int(counts[-1].item()) >= cells:
            keep = int((counts < cells).sum().item()) + 1
            parts[-1] = parts[-1][: keep * B]
            break
    return torch.cat(parts).contiguous()


hw = 40000
coords = cover_list(hw, hw)
logits = torch.randn((coords.shape[0], N), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
for d in ([int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else (16, 8, 4)):
    dh = dw = hw // d
    for label, kw in (("class map", dict(want_sum=False, want_argmax=True)), ("class + count", dict(want_sum=False, want_argmax=True, want_count=True)),
                      ("sum + class + count", dict(want_sum=True, want_argmax=True, want_count=True))):
        res = {}
        for variant in (0, 1, 2):
            lib.dh_stitch_binned_set_variant(variant)
            keep = {}

            def run():
                keep["o"] = None
                keep["o"] = ops.stitch_binned(logits, coords, PS, d, dh, dw, **kw)

            ms = timeit(run)
            res[variant] = keep["o"]
            keep.clear()
            print(f"d={d}  out={label}  variant={variant}  ms={ms:.4f}", file=sys.stderr)
        same = all((x is None and y is None) or torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x, y.view(torch.int32) if y.dtype == torch.float32 else y)
                   for x, y in zip(res[0], res[1]))
        print(f"d={d}  out={label}  variant 0 == variant 1: {same}", file=sys.stderr)
        del res
        torch.cuda.empty_cache()
lib.dh_stitch_binned_set_variant(0)
