"""Throughput of the coordinate-side kernels (not a product path): these move almost no bytes (<= 16 B per candidate), so
they are reported as candidates/s or batches/s and are latency / float64-ALU bound, not HBM bound.
Run on a B200:  python profiles/sampler_sweep.py > gpurun_out/sampler_sweep.json"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from deephisto_b200 import geometry, ops  # noqa: E402
from deephisto_b200.patch_samplers.region_samplers import build_tables  # noqa: E402
from deephisto_b200.synthetic import synth_polygons  # noqa: E402


def timeit(fn, reps=20):
    for _ in range(3):
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
H = W = 32768
polys = synth_polygons(50, H, W, seed=0)
tables, regions, classes = build_tables([((H, W), polys)], layer=1, area_influence=0.5, classes=None, one_image_for_batch=True, device="cuda")
thr = 224 * 224 * 0.75
for n_slots in (256, 4096, 65536, 1 << 20):
    out = ops.region_sample(tables.struct, n_slots, 4, 224, thr, seed=1)
    ms = timeit(lambda: ops.region_sample(tables.struct, n_slots, 4, 224, thr, seed=1, out=out))
    rows.append({"kernel": "region_sample_kernel (50 polygons, 24-64 vertices, k=4, ri=0.75)", "n": n_slots, "ms": ms, "per_s": n_slots / ms * 1e3, "unit": "accepted slots/s"})

# dense acceptance over the candidate grid of the largest polygon at stride 8
big = max(regions, key=lambda r: r.area)
edges = torch.from_numpy(big.edges.reshape(-1)).cuda()
x0, y0, x1, y1 = (round(v) for v in big.bounds)
for stride in (56, 8, 2):
    ny, nx = len(range(y0, min(y1, H - 224), stride)), len(range(x0, min(x1, W - 224), stride))
    ms = timeit(lambda: ops.region_accept_dense(edges, 0, len(big.edges), y0, x0, ny, nx, stride, 224, thr))
    rows.append({"kernel": f"region_accept_dense_kernel ({len(big.edges)} edges)", "n": ny * nx, "ms": ms, "per_s": ny * nx / ms * 1e3, "unit": "candidates/s"})

for (h, w, B) in ((8192, 8192, 64), (40000, 40000, 64), (40000, 40000, 1024)):
    st = ops.CoverState(h, w, 224, 16, 2, B, seed=0)
    ms = timeit(lambda: st.next_coords(), reps=30)
    rows.append({"kernel": f"dh_cover_sample ({h}x{w} slide, coarse grid {h // 16}x{w // 16}, incremental state, 1 launch)", "n": B, "ms": ms, "per_s": B / ms * 1e3, "unit": "patches/s (coordinates only)"})
    if (h, B) != (40000, 64):
        continue                                        # smaller cases reach full coverage inside the timing loop (launches become no-ops)
    st = ops.CoverState(h, w, 224, 16, 2, B, seed=1)
    ms = timeit(lambda: st.next_group(16), reps=30) / 16
    rows.append({"kernel": f"dh_cover_sample_group ({h}x{w} slide, 16 batches per call: per batch)", "n": B, "ms": ms, "per_s": B / ms * 1e3, "unit": "patches/s (coordinates only)"})

for (h, w) in ((40000, 40000), (100000, 100000)):
    n, npad = ops.dense_count(h, w, 224, 112, 64)
    ms = timeit(lambda: ops.dense_coords(h, w, 224, 112, 64))
    rows.append({"kernel": f"dense_coords_kernel ({h}x{w}, stride 112)", "n": npad, "ms": ms, "per_s": npad / ms * 1e3, "unit": "coords/s"})

e = [geometry.build_edges(np.asarray(p["vertices"])) for p in polys]
off = np.zeros(len(e) + 1, np.int32)
off[1:] = np.cumsum([len(x) for x in e])
bb = np.asarray([geometry.polygon_bounds(np.asarray(p["vertices"])) for p in polys]).reshape(-1)
ed, of, bbd = torch.from_numpy(np.concatenate(e).reshape(-1)).cuda(), torch.from_numpy(off).cuda(), torch.from_numpy(bb).cuda()
for scale in (16.0, 4.0):
    mh = int(H / scale)
    ms = timeit(lambda: ops.rasterize_polygons(ed, of, bbd, scale, mh, mh), reps=10)
    rows.append({"kernel": f"rasterize_kernel (50 polygons -> {mh}x{mh} label map)", "n": mh * mh, "ms": ms, "per_s": mh * mh / ms * 1e3, "unit": "mask pixels/s"})

print(json.dumps({"rows": rows}, indent=1))
for r in rows:
    print(f'{r["kernel"]:78s} n={r["n"]:9d} {r["ms"]:9.4f} ms  {r["per_s"]:.4g} {r["unit"]}', file=sys.stderr)
