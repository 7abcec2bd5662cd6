"""Roofline sweep of dh_gather_normalize (not a product path): CUDA-event time per launch for every output mode and
several batch sizes on a 32768^2 slide with random patch origins, plus the load-only / store-only ceilings of the
TMA-staged kernel. Prints one JSON object; run on a B200:  python profiles/gather_sweep.py > gpurun_out/gather_sweep.json"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deephisto_b200 import ops  # noqa: E402

H = W = 32768
PS = 224
slide = ops.DeviceSlide.synthetic(H, W, 0)
g = torch.Generator(device="cuda").manual_seed(0)
peak = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if \
    (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else 6650.0


def run(B, dtype, layout, variant="auto", reps=40, nbuf=3, slide=slide, tag=""):
    H, W = slide.H, slide.W
    coords = torch.stack([torch.randint(0, H - PS + 1, (nbuf * B,), generator=g, device="cuda"),
                          torch.randint(0, W - PS + 1, (nbuf * B,), generator=g, device="cuda")], 1).to(torch.int32).contiguous()
    shape = (B, PS, PS, 3) if layout == "NHWC" else (B, 3, PS, PS)
    outs = [torch.empty(shape, dtype=dtype, device="cuda") for _ in range(nbuf)]   # rotate: outputs larger than L2 in total
    ops.set_gather_variant(variant)
    for i in range(6):
        ops.gather_normalize(slide, coords[(i % nbuf) * B:(i % nbuf + 1) * B], PS, dtype=dtype, layout=layout, out=outs[i % nbuf])
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(reps * 60e-6 * 2e9))   # keep the GPU busy while the CPU queues every launch: no launch-rate gaps in the timings
    t0.record()
    for i in range(reps):
        evs[i][0].record()
        ops.gather_normalize(slide, coords[(i % nbuf) * B:(i % nbuf + 1) * B], PS, dtype=dtype, layout=layout, out=outs[i % nbuf])
        evs[i][1].record()
    t1.record()
    torch.cuda.synchronize()
    ops.set_gather_variant("auto")
    ms = sorted(a.elapsed_time(b) for a, b in evs)[reps // 2]
    back_to_back = t0.elapsed_time(t1) / reps
    esz = 4 if dtype == torch.float32 else 2
    alg = B * PS * PS * 3 * (1 + esz)
    return {"B": B, "dtype": str(dtype).split(".")[-1], "layout": layout, "variant": variant + tag, "us_median": 1e3 * ms, "us_back_to_back": 1e3 * back_to_back,
            "alg_MB": alg / 1e6, "GBs": alg / ms / 1e6, "GBs_b2b": alg / back_to_back / 1e6, "frac_of_measured": alg / ms / 1e6 / peak}


rows = []
for B in (256, 1024, 4096):
    for dtype in (torch.float32, torch.bfloat16):
        for layout in ("NHWC", "NCHW"):
            rows.append(run(B, dtype, layout))
for v in ("direct", "tma_noload", "tma_nostore", "tma_nomem", "tma_plainstore", "tma_blocked", "tma_evictfirst", "tma_blocked_evictfirst"):
    rows.append(run(256, torch.float32, "NHWC", v))
    rows.append(run(4096, torch.float32, "NHWC", v))
import os
for st in (2, 3, 6, 8):
    os.environ["DH_GATHER_STAGES"] = str(st)
    rows.append(run(256, torch.float32, "NHWC", tag=f"_stages{st}"))
    rows.append(run(4096, torch.float32, "NHWC", tag=f"_stages{st}"))
os.environ.pop("DH_GATHER_STAGES")
# read-locality experiment: the same bytes per patch, but slide rows so short that a patch is (nearly) one contiguous block
del slide
for w in (224,):
    narrow = ops.DeviceSlide.synthetic(3_000_000_000 // (3 * w), w, 1)
    rows.append(run(256, torch.float32, "NHWC", slide=narrow, tag=f"_W{w}"))
    rows.append(run(4096, torch.float32, "NHWC", slide=narrow, tag=f"_W{w}"))
    rows.append(run(4096, torch.float32, "NHWC", "tma_nostore", slide=narrow, tag=f"_W{w}"))
    del narrow
print(json.dumps({"peak_gbs": peak, "rows": rows}, indent=1))
for r in rows:
    print(f'{r["B"]:5d} {r["dtype"]:9s} {r["layout"]} {r["variant"]:18s} {r["us_median"]:8.1f} us  b2b {r["us_back_to_back"]:8.1f} us  {r["GBs"]:7.0f} GB/s  {r["frac_of_measured"]:.3f}', file=sys.stderr)
