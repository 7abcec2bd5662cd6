"""Calibration (not a product path): what a plain library kernel achieves on this B200 for write-only and copy streams of the
sizes our launches move. Gives the realistic ceiling for short, write-dominated launches (gather batch = 193 MB)."""
import json
import torch

def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mb in (154, 193, 1024, 4096):
    n = mb * (1 << 20) // 4
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    y = torch.empty(n, dtype=torch.float32, device="cuda")
    ms = timeit(lambda: x.fill_(1.0))
    out[f"fill_{mb}MB_GBs"] = mb * (1 << 20) / ms / 1e6
    ms = timeit(lambda: y.copy_(x))
    out[f"copy_{mb}MB_rw_GBs"] = 2 * mb * (1 << 20) / ms / 1e6
    # u8 -> f32 conversion (1 B read, 4 B written), library elementwise kernel
    u = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: y.copy_(u))
    out[f"u8_to_f32_{mb}MBout_GBs"] = 1.25 * mb * (1 << 20) / ms / 1e6
print(json.dumps(out, indent=1))
