"""Roofline sweep of dh_stitch_binned (deterministic stitch of an arbitrary coordinate list; not a product path):
CUDA-event time of the whole call (count + alloc + fill + tile kernels) on the 40k x 40k case, for
  (a) the coverage-driven random sampler's own coordinates (FullImageRndSampler, dense_level 2, run to coverage 1.0), and
  (b) the dense enumeration at stride 112 (127 488 patches; the same list dh_stitch_dense handles in closed form),
next to dh_stitch_scatter (atomics) on the same list. Algorithmic bytes = requested outputs written once + logits read once.
    python profiles/stitch_binned_sweep.py > gpurun_out/stitch_binned_sweep.json"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from deephisto_b200 import _lib, ops  # noqa: E402

peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
lib = _lib.require_device()
PS, N = 224, 5
CODES = [int(c) for c in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 16, 32, 64, 128]
H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 40000     # e.g. 39999: rows not 16-byte aligned -> scalar-store tile kernel


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


st = ops.CoverState(H, W, PS, 16, 2, 1024, seed=0)
cells16 = (H // 16) * (W // 16)
parts = []
while True:
    c, nz = st.next_coords()
    parts.append(c.clone())
    if len(parts) % 16 == 0 and int(nz.item()) >= cells16:
        break
lists = {"coverage sampler": torch.cat(parts), "dense enumeration": ops.dense_coords(H, W, PS, 112, 64)}
rows = []
for name, coords in lists.items():
    P = coords.shape[0]
    logits = torch.randn((P, N), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
    for d in (16, 4, 2, 1):
        dh, dw = H // d, W // d
        cells = dh * dw
        reps = 10 if d >= 4 else 3
        for label, kw, out_bytes in (("sum", dict(want_sum=True), cells * N * 4), ("argmax", dict(want_sum=False, want_argmax=True), cells)):
            # override code = tile rows + 1000 * groups + 100000 * extra smem KB per CTA (0 = the library's heuristic)
            for th in ((0,) if label == "argmax" else CODES):
                lib.dh_stitch_binned_set_tile_rows(th)
                keep = {}

                def run():
                    keep["o"] = None
                    keep["o"] = ops.stitch_binned(logits, coords, PS, d, dh, dw, **kw)

                ms = timeit(run, reps)
                keep.clear()
                alg = out_bytes + P * N * 4
                rows.append({"kernel": "stitch_binned", "list": name, "P": P, "d": d, "outputs": label, "tile_rows": th, "ms": ms, "alg_MB": alg / 1e6,
                             "GBs": alg / ms / 1e6, "frac_of_measured": alg / ms / 1e6 / peak})
        lib.dh_stitch_binned_set_tile_rows(0)
        if d >= 2:
            sum_map = torch.zeros((dh, dw, N), dtype=torch.float32, device="cuda")
            ms = timeit(lambda: ops.stitch_scatter(logits, coords, PS, d, sum_map, None), reps)
            alg = cells * N * 4 + P * N * 4
            rows.append({"kernel": "stitch_scatter", "list": name, "P": P, "d": d, "outputs": "sum (atomics)", "tile_rows": 0, "ms": ms, "alg_MB": alg / 1e6,
                         "GBs": alg / ms / 1e6, "frac_of_measured": alg / ms / 1e6 / peak})
            del sum_map
        torch.cuda.empty_cache()
print(json.dumps({"peak_gbs": peak, "case": f"{H}x{W} ps{PS} n{N}", "rows": rows}, indent=1))
for r in rows:
    print(f'{r["kernel"]:15s} {r["list"]:18s} P={r["P"]:7d} d={r["d"]:2d} {r["outputs"]:14s} code={r["tile_rows"]:7d} {r["ms"]:9.3f} ms {r["alg_MB"]:9.1f} MB '
          f'{r["GBs"]:7.0f} GB/s  {r["frac_of_measured"]:.3f}', file=sys.stderr)
