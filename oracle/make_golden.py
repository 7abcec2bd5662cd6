"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) through numpy stubs of
its absent third-party modules (oracle/refstubs). Run in the build container:

    python -m oracle.make_golden

The GPU box has no /root/reference; it only reads the committed .npz files. Inputs are the seeded
synthetic slides of oracle/synth.py, so they are regenerated on any box instead of being stored."""

from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np

from . import reference_loader as rl
from . import synth

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"

# (H, W, ps, stride, batch) -- SURVEY 8c G1
DENSE_CASES = [
    (8192, 8192, 224, 224, 16),
    (8192, 8192, 224, 112, 16),
    (1000, 777, 224, 100, 7),     # non-multiple edges
    (448, 448, 224, 224, 4),      # (h - ps) % stride == 0
    (224, 500, 224, 64, 3),       # h == ps: empty main grid rows
    (300, 260, 32, 48, 5),        # stride > ps (gaps)
]
DENSE_COUNT_ONLY = [(40000, 40000, 224, 112, 64), (100000, 100000, 224, 112, 64)]
# cases whose pixels are digested per batch (float32 NHWC, as yielded by generator_torch)
PIXEL_CASES = [(8192, 8192, 224, 224, 16), (1000, 777, 224, 100, 7), (300, 260, 32, 48, 5)]
STITCH_CASES = [  # (H, W, ps, stride, batch, n, downscales)
    (2048, 2048, 224, 112, 64, 5, (16, 4, 1)),
    (1000, 777, 224, 100, 7, 5, (16, 3, 1)),
    (300, 260, 32, 48, 5, 3, (4, 1)),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def dense_goldens(fs, im, out: dict, manifest: dict):
    mode = fs.SamplerExecutionMode.INMEMORY_SINGLEPROC
    for H, W, ps, stride, B in DENSE_CASES:
        key = f"dense_{H}x{W}_ps{ps}_s{stride}_b{B}"
        slide = synth.synth_slide(H, W, seed=0)
        im.register(key, slide)
        s = fs.FullImageDenseSampler(key, layer=1, patch_size=ps, batch_size=B, mode=mode, stride=stride)
        coords = np.array([c for b in s._create_batched_coords() for c in b], dtype=np.int32)
        out[key + "_coords"] = coords
        entry = {"n_padded": int(len(coords))}
        if (H, W, ps, stride, B) in PIXEL_CASES:
            digests, csum = [], np.zeros(3, dtype=np.float64)
            first = None
            for feats, crd, progress in s.generator_torch():
                f = feats.numpy()
                assert f.dtype == np.float32
                digests.append(sha(f))
                csum += f.reshape(-1, 3).sum(axis=0, dtype=np.float64)
                if first is None:
                    first = f[:2].copy()
                    out[key + "_coords_f32_b0"] = crd.numpy()
            entry["batch_sha256"] = digests
            entry["channel_sum_f64"] = csum.tolist()
            if ps <= 64:
                out[key + "_first2"] = first
        manifest[key] = entry
    for H, W, ps, stride, B in DENSE_COUNT_ONLY:
        # the reference's list comprehension over 127k / 796k tuples is still cheap
        class _S:  # only the attributes _create_batched_coords touches
            pass
        s = _S()
        s.h, s.w, s.patch_size, s.stride, s.batch_size = H, W, ps, stride, B
        cb = fs.FullImageDenseSampler._create_batched_coords(s)
        flat = np.array([c for b in cb for c in b], dtype=np.int32)
        manifest[f"dense_{H}x{W}_ps{ps}_s{stride}_b{B}"] = {
            "n_padded": int(len(flat)), "n_batches": len(cb), "coords_sha256": sha(flat),
            "n_unpadded": int(len(flat) - (np.all(flat[::-1] == flat[-1], axis=1).cumprod().sum() - 1)),
        }


def stitch_goldens(pfp, fs, im, out: dict, manifest: dict):
    """ImagePredictorPatched.process (unmodified) driven by FullImageDenseSampler.generator() and a predictor
    that returns seeded float32 logits per batch; the pre-argmax sum map is captured by wrapping np.argmax."""
    mode = fs.SamplerExecutionMode.INMEMORY_SINGLEPROC

    class _Anno:
        def __init__(self, n):
            self.anno_classes = list(range(n))

    for H, W, ps, stride, B, n, ds in STITCH_CASES:
        key = f"stitch_{H}x{W}_ps{ps}_s{stride}_b{B}_n{n}"
        slide = np.zeros((H, W, 3), dtype=np.uint8)  # pixels are irrelevant to the stitcher
        im.register(key, slide)
        n_pad = len([c for b in fs.FullImageDenseSampler._create_batched_coords(
            type("S", (), dict(h=H, w=W, patch_size=ps, stride=stride, batch_size=B))()) for c in b])
        rng = np.random.default_rng(1234)
        logits = rng.standard_normal((n_pad, n)).astype(np.float32) * np.float32(3.0)
        out[key + "_logits"] = logits
        for d in ds:
            sampler = fs.FullImageDenseSampler(key, layer=1, patch_size=ps, batch_size=B, mode=mode, stride=stride)
            pos = {"i": 0}

            def predictor(patches):
                i = pos["i"]
                pos["i"] += len(patches)
                return logits[i : i + len(patches)]

            captured = {}
            real_argmax = np.argmax

            def spy(a, axis=None):
                captured["sum"] = a.copy()
                return real_argmax(a, axis=axis)

            pred = pfp.ImagePredictorPatched(key, sampler.generator(), predictor, _Anno(n), layer=1, downscale=d)
            pfp.np.argmax = spy
            try:
                amax = pred.process()
            finally:
                pfp.np.argmax = real_argmax
            s = captured["sum"]
            assert s.dtype == np.float32
            manifest[f"{key}_d{d}"] = {"sum_sha256": sha(s), "argmax_sha256": sha(amax.astype(np.uint8)), "shape": list(s.shape)}
            if s.nbytes <= 2_000_000:
                out[f"{key}_d{d}_sum"] = s
                out[f"{key}_d{d}_argmax"] = amax.astype(np.uint8)


def region_polygons():
    """SURVEY 8c G4 shapes: convex, concave star, integer rectangle (tie), bbox partly outside, too small."""
    rng = np.random.default_rng(7)
    ang = np.sort(rng.uniform(0, 2 * np.pi, 40))
    star = np.stack([1500 + (400 + 500 * (np.arange(40) % 2)) * np.cos(ang), 1400 + (400 + 500 * (np.arange(40) % 2)) * np.sin(ang)], 1)
    hexa = np.stack([1000 + 700 * np.cos(np.arange(6) * np.pi / 3 + 0.1), 1200 + 700 * np.sin(np.arange(6) * np.pi / 3 + 0.1)], 1)
    return {
        "convex": hexa,
        "star": star,
        "rect_tie": np.array([[100.0, 100.0], [100.0 + 224 * 3 + 56, 100.0], [100.0 + 224 * 3 + 56, 100.0 + 224 * 2], [100.0, 100.0 + 224 * 2]]),
        "outside": np.array([[1700.0, 1500.0], [2300.0, 1450.0], [2400.0, 2100.0], [1650.0, 2200.0]]),
        "small": np.array([[10.0, 10.0], [120.0, 10.0], [120.0, 130.0], [10.0, 130.0]]),
        "frac": np.array([[300.25, 200.5], [1200.75, 310.125], [1100.5, 1250.875], [250.125, 1000.25], [600.0, 600.0]]),
    }


def region_goldens(rs_mod, im, out: dict, manifest: dict):
    """RegionAnnotation._extract_patch_coords_dense (unmodified control flow; geometry through the shapely stub)."""
    H, W = 2048, 2048
    im.register("regions", np.zeros((H, W, 3), np.uint8))
    for name, verts in region_polygons().items():
        for layer in (1, 2):
            reg = rs_mod.RegionAnnotation(Path("regions"), 0, "X", verts.astype(np.float64), layer=layer, layer_size=(H // layer, W // layer))
            for ps, stride, ri in ((224, 112, 0.75), (224, 56, 0.5), (64, 32, 0.95)):
                coords = np.asarray(reg._extract_patch_coords_dense(ps, stride, ri), dtype=np.int32).reshape(-1, 2)
                key = f"region_{name}_l{layer}_ps{ps}_s{stride}_ri{ri}"
                out[key] = coords
                manifest[key] = {"n": int(len(coords)), "area": float(reg.area)}
        out[f"region_{name}_verts"] = verts


WEIGHT_CASES = [  # (area_influence, one_image_for_batch)
    (0.5, True), (0.5, False), (0.0, True), (-0.7, False), (1.0, True), (-1.0, True),
]


def weights_dataset(tmp: Path):
    """Two synthetic images with 11 and 7 polygons in 4 / 3 classes, written in the reference's JSON schema."""
    hw = (6000, 5000)
    items, polys_all = [], []
    for j, (n, ncls) in enumerate(((11, 4), (7, 3))):
        polys = synth.synth_polygons(n, *hw, seed=50 + j, rmin=250, rmax=900, n_classes=ncls)
        anno = tmp / f"w{j}.json"
        anno.write_text(json.dumps(polys))
        items.append((Path(f"weights_img{j}"), anno))
        polys_all.append(polys)
    return hw, items, polys_all


def weights_goldens(rs_mod, im, manifest: dict):
    """AnnoRegionRndSampler.__init__ (unmodified: _parse_annotations :194-249, _calc_weights :395-482, _calc_area_weights
    :339-378) on a synthetic two-image dataset; the weight tables it computes are the golden values of row F."""
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        hw, items, polys_all = weights_dataset(Path(td))
        for path, _ in items:
            im.register(path, np.zeros(hw + (3,), np.uint8))
        out = {"hw": list(hw), "cases": []}
        for infl, one in WEIGHT_CASES:
            s = rs_mod.AnnoRegionRndSampler(items, layer=1, patch_size=224, region_area_influence=infl, one_image_for_batch=one)
            out["cases"].append({
                "area_influence": infl, "one_image_for_batch": one, "classes": list(s.classes), "len": len(s),
                "reg_w_all": {c: np.asarray(w).tolist() for c, w in s._reg_w_all.items()},
                "reg_w_per_img": [{c: np.asarray(w).tolist() for c, w in d.items()} for d in s._reg_w_per_img],
                "img_w": {c: np.asarray(w).tolist() for c, w in s._img_w.items()},
                "img_w_all": np.asarray(s._img_w_all).tolist(),
                "areas_all": {c: [r.area for r in regs] for c, regs in s.regions.items()},
            })
    (OUT / "golden_weights_v1.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    manifest["weights"] = {"file": "golden_weights_v1.json", "cases": len(out["cases"])}


STAT_COVER_CASES = [(1100, 900, 224, 16), (2048, 2048, 224, 64)]     # (H, W, ps, batch)
STAT_SEEDS = list(range(12))


def statistics_goldens(fs, rs_mod, im, manifest: dict):
    """Row B / C statistics of the UNMODIFIED reference with its global numpy RNG seeded (np.random.seed): the random streams
    cannot be compared draw by draw with the Philox streams of the new build, their distributions can.
      cover:  FullImageRndSampler.generator() (full_samplers.py:263-274) -- batches until filled_ratio reaches 1, mean accumulator value
      region: RegionAnnotation._extract_patch_coords_rnd (region_samplers.py:82-143) -- mean / std of the accepted (y, x)"""
    mode = fs.SamplerExecutionMode.INMEMORY_SINGLEPROC
    out = {"seeds": STAT_SEEDS, "cover": [], "region": []}
    for H, W, ps, B in STAT_COVER_CASES:
        key = f"stat_cover_{H}x{W}"
        im.register(key, np.zeros((H, W, 3), np.uint8))
        n_batches, mean_acc, first_ratio = [], [], []
        for seed in STAT_SEEDS:
            np.random.seed(seed)
            s = fs.FullImageRndSampler(key, layer=1, patch_size=ps, batch_size=B, mode=mode)
            ratios = [fr for _patches, fr in s.generator()]
            n_batches.append(len(ratios))
            first_ratio.append(ratios[0])
            mean_acc.append(float(s._accum.mean()))
        out["cover"].append({"H": H, "W": W, "ps": ps, "batch": B, "n_batches": n_batches, "mean_accum": mean_acc, "first_ratio": first_ratio})
    polys = region_polygons()
    for name in ("star", "convex", "frac"):
        verts = polys[name].astype(np.float64)
        reg = rs_mod.RegionAnnotation(Path("regions"), 0, "X", verts, layer=1, layer_size=(2048, 2048))
        for ps, ri in ((224, 0.75), (64, 0.95)):
            np.random.seed(123)
            c = np.asarray(reg._extract_patch_coords_rnd(ps, 6000, ri, miss_limit=500), dtype=np.float64)
            out["region"].append({"polygon": name, "ps": ps, "ri": ri, "n": len(c), "mean_yx": c.mean(0).tolist(), "std_yx": c.std(0).tolist(),
                                  "min_yx": c.min(0).tolist(), "max_yx": c.max(0).tolist()})
    (OUT / "golden_stats_v1.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    manifest["statistics"] = {"file": "golden_stats_v1.json"}


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    out, manifest = {}, {}
    with rl.reference_modules("patch_samplers.full_samplers", "examples.predict_full_patched", "patch_samplers.region_samplers") as (fs, pfp, rs):
        from psimage.core import image as im  # the stub

        dense_goldens(fs, im, out, manifest)
        stitch_goldens(pfp, fs, im, out, manifest)
        region_goldens(rs, im, out, manifest)
        weights_goldens(rs, im, manifest)
        statistics_goldens(fs, rs, im, manifest)
    np.savez_compressed(OUT / "golden_v1.npz", **out)
    (OUT / "golden_v1.json").write_text(json.dumps(manifest, indent=1, sort_keys=True))
    print(f"wrote {OUT / 'golden_v1.npz'} ({(OUT / 'golden_v1.npz').stat().st_size / 1e6:.2f} MB), {len(out)} arrays, {len(manifest)} manifest entries")


if __name__ == "__main__":
    main()
