"""Roofline sweep of the stitch kernels (not a product path): CUDA-event time per launch of dh_stitch_dense(_ex),
dh_stitch_scatter and dh_stitch_finalize for the 40k x 40k / stride 112 case (BASELINE configs[2]) at several downscales.
Algorithmic bytes = every requested output written once + the logits read once. Run on a B200:
    python profiles/stitch_sweep.py > gpurun_out/stitch_sweep.json"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from deephisto_b200 import ops  # noqa: E402

peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 40000     # e.g. 39999: dw * n is not a multiple of 4 floats -> the PHASED store path
PS, STRIDE, B, N = 224, 112, 64, 5
n, npad = ops.dense_count(H, W, PS, STRIDE, B)
g = torch.Generator(device="cuda").manual_seed(0)
logits = torch.randn((npad, N), generator=g, device="cuda")
coords = ops.dense_coords(H, W, PS, STRIDE, B)


def timeit(fn, reps):
    for _ in range(2):
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


rows = []
for d in (16, 4, 2, 1):
    dh, dw = H // d, W // d
    cells = dh * dw
    reps = 20 if d >= 4 else 5
    for label, kw, out_bytes in (("sum", dict(want_sum=True), cells * N * 4),
                                 ("sum+argmax", dict(want_sum=True, want_argmax=True), cells * (N * 4 + 1)),
                                 ("sum+count+argmax", dict(want_sum=True, want_count=True, want_argmax=True), cells * (N * 4 + 5)),
                                 ("argmax", dict(want_sum=False, want_argmax=True), cells)):
        keep = {}

        def run():
            keep["o"] = None                                  # free the previous outputs before allocating the next (d=1: 32 GB)
            keep["o"] = ops.stitch_dense(logits, H, W, PS, STRIDE, d, B, **kw)

        ms = timeit(run, reps)
        alg = out_bytes + npad * N * 4
        rows.append({"kernel": "stitch_dense", "d": d, "outputs": label, "ms": ms, "alg_MB": alg / 1e6, "GBs": alg / ms / 1e6, "frac_of_measured": alg / ms / 1e6 / peak})
        keep.clear()
    if d >= 2:
        sum_map = torch.zeros((dh, dw, N), dtype=torch.float32, device="cuda")
        ms = timeit(lambda: ops.stitch_scatter(logits, coords, PS, d, sum_map, None), reps)
        alg = cells * N * 4 + npad * N * 4
        rows.append({"kernel": "stitch_scatter", "d": d, "outputs": "sum (atomics)", "ms": ms, "alg_MB": alg / 1e6, "GBs": alg / ms / 1e6, "frac_of_measured": alg / ms / 1e6 / peak})
        ms = timeit(lambda: ops.stitch_finalize(sum_map, None, want_norm=False, want_argmax=True), reps)
        alg = cells * (N * 4 + 1)
        rows.append({"kernel": "stitch_finalize", "d": d, "outputs": "argmax", "ms": ms, "alg_MB": alg / 1e6, "GBs": alg / ms / 1e6, "frac_of_measured": alg / ms / 1e6 / peak})
        del sum_map
    torch.cuda.empty_cache()
# prediction post-processing: class-colour mask + area-average thumbnail + overlay in one pass over the slide
slide = ops.DeviceSlide.synthetic(H, W, 0)
lut = torch.randint(0, 256, (256, 3), dtype=torch.uint8, device="cuda")
for d in (16, 4):
    dh, dw = H // d, W // d
    am = torch.randint(0, N, (dh, dw), dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: ops.colorize_overlay(am, lut, slide, d, want_mask=True, want_thumb=True, want_overlay=True), 10)
    alg = dh * d * dw * d * 3 + dh * dw * (1 + 9)
    rows.append({"kernel": "colorize_overlay", "d": d, "outputs": "mask+thumb+overlay", "ms": ms, "alg_MB": alg / 1e6, "GBs": alg / ms / 1e6, "frac_of_measured": alg / ms / 1e6 / peak})
print(json.dumps({"peak_gbs": peak, "case": f"{H}x{W} ps{PS} stride{STRIDE} n{N}, {npad} patches", "rows": rows}, indent=1))
for r in rows:
    print(f'{r["kernel"]:16s} d={r["d"]:2d} {r["outputs"]:18s} {r["ms"]:9.3f} ms {r["alg_MB"]:10.1f} MB {r["GBs"]:8.0f} GB/s  {r["frac_of_measured"]:.3f}', file=sys.stderr)
