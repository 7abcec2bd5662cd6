"""The C-ABI library loads without a GPU and exports every symbol include/deephisto_b200.h declares."""

import re
import subprocess
from pathlib import Path

import pytest

from deephisto_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "deephisto_b200.h"


def declared_symbols():
    return sorted(set(re.findall(r"^DH_API [\w\s\*]+?\b(dh_\w+)\(", HEADER.read_text(), flags=re.M)))


def test_header_symbols_all_bound_and_exported():
    syms = declared_symbols()
    assert len(syms) >= 18
    assert sorted(_lib.SIGNATURES) == syms
    lib = _lib.load()  # no GPU needed
    for s in syms:
        assert hasattr(lib, s), s
    exported = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    names = {line.split()[-1] for line in exported.splitlines() if " T " in line}
    assert set(syms) <= names
    assert {n for n in names if n.startswith("dh_")} == set(syms), "exported dh_* symbol missing from the header"


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.dh_version() == 100
    from deephisto_b200 import ops

    assert ops.dense_count(8192, 8192, 224, 224, 16) == (1369, 1376)
    assert ops.dense_count(40000, 40000, 224, 112, 64) == (127449, 127488)
    assert ops.dense_count(100000, 100000, 224, 112, 64) == (795664, 795712)
    with pytest.raises(ValueError):
        ops.dense_count(100, 100, 224, 112, 64)
    assert "smaller than patch" in _lib.last_error()
    assert lib.dh_cover_scratch_words(2500, 2500) > 2500 * 2500 // 32


def test_profiling_switches_validate_their_argument():
    """The A/B switches between kernel formulations are host-only setters: every documented value is accepted, anything else comes
    back as a negative status with a message naming the valid values; the scratch size is the same for every variant (the tile
    geometries of all of them are covered by dh_stitch_binned_scratch_bytes)."""
    lib = _lib.load()
    try:
        for name, valid in (("dh_stitch_binned_set_variant", range(0, 5)), ("dh_stitch_dense_set_variant", range(0, 2)), ("dh_cover_set_variant", range(0, 2))):
            fn = getattr(lib, name)
            for v in valid:
                assert fn(v) == 0, (name, v)
            for bad in (-1, max(valid) + 1, 99):
                assert fn(bad) < 0, (name, bad)
                assert "variant" in _lib.last_error()
        sizes = set()
        for v in range(0, 4):
            assert lib.dh_stitch_binned_set_variant(v) == 0
            sizes.add(lib.dh_stitch_binned_scratch_bytes(124928, 224, 4, 5, 10000, 10000))
        assert len(sizes) == 1 and sizes.pop() > 124928 * 4
    finally:
        lib.dh_stitch_binned_set_variant(0)
        lib.dh_stitch_dense_set_variant(0)
        lib.dh_cover_set_variant(0)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.DeepHistoError):
        _lib.require_device()
    from deephisto_b200 import ops

    with pytest.raises(_lib.DeepHistoError):
        ops.gather_normalize(None, torch.zeros((1, 2), dtype=torch.int32), 8)


def test_missing_library_fails_loudly(tmp_path):
    """No CPU fallback: with the shared library absent every compute entry point raises (checked in a fresh interpreter)."""
    import os
    import sys

    code = ("import sys; sys.path.insert(0, %r)\n"
            "from deephisto_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.DeepHistoError as e:\n    print('RAISED', 'no CPU fallback' in str(e).lower() or 'not found' in str(e))\n") % str(ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, DEEPHISTO_B200_LIB=str(tmp_path / "nope.so")))
    assert "RAISED True" in out.stdout, out.stdout + out.stderr


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under deephisto_b200/ may import it."""
    offenders = []
    for p in (ROOT / "deephisto_b200").rglob("*.py"):
        txt = p.read_text()
        if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
            offenders.append(str(p))
    assert not offenders, offenders


def test_slide_sources_host_side():
    """The psimage seam (deephisto_b200/slide.py): duck-typed sources agree on layer sizes and regions without touching the GPU."""
    import numpy as np

    from deephisto_b200 import slide as sl

    a = np.arange(40 * 30 * 3, dtype=np.uint8).reshape(40, 30, 3)
    s = sl.open_slide(a)
    assert isinstance(s, sl.ArraySlide) and s.layer_size(1) == (40, 30) and s.layer_size(2) == (20, 15)
    assert np.array_equal(s.get_region_from_layer(1, (3, 4), (10, 9)), a[3:10, 4:9])
    assert np.array_equal(s.get_region_from_layer(2, (0, 0), (20, 15)), a[::2, ::2])
    assert s.get_region((0, 0), (40, 30), target_hw=(10, 6)).shape == (10, 6, 3)
    with pytest.raises(ValueError):
        sl.ArraySlide(a.astype(np.float32))
    with pytest.raises(TypeError):
        sl.open_slide(42)
    with pytest.raises(RuntimeError, match="psimage"):
        sl.open_slide("missing.psi")
    syn = sl.SyntheticSlide(64, 48, seed=3)
    assert syn.layer_size(1) == (64, 48)
    with pytest.raises(ValueError):
        syn.layer_size(2)


def test_pinned_slide_band_origin_host_side():
    """PinnedSlide as one rank's row band of a sharded slide: row arguments stay in layer coordinates, rows outside the band raise."""
    import numpy as np

    from deephisto_b200 import slide as sl

    a = np.arange(40 * 30 * 3, dtype=np.uint8).reshape(40, 30, 3)
    whole = sl.PinnedSlide.from_numpy(a)
    assert whole.layer_size(1) == (40, 30) and whole.nbytes == 40 * whole.pitch and whole.pitch % 16 == 0
    assert np.array_equal(whole.get_region_from_layer(1, (3, 4), (10, 9)), a[3:10, 4:9])
    band = sl.PinnedSlide(whole.host[12 * whole.pitch : 25 * whole.pitch], 13, 30, whole.pitch, y_origin=12, full_height=40)
    assert band.layer_size(1) == (40, 30) and band.nbytes == 13 * whole.pitch
    assert np.array_equal(band.get_region_from_layer(1, (14, 0), (20, 30)), a[14:20])
    with pytest.raises(ValueError, match="outside the band"):
        band.get_region_from_layer(1, (5, 0), (14, 30))
    with pytest.raises(ValueError):
        band.layer_size(2)


def test_tile_spans_cover_the_rectangles_once():
    """slide.tile_spans (input of dh_upload_rects): the spans cover every requested pixel, never overlap, stay 16-byte aligned."""
    import numpy as np

    from deephisto_b200 import ops, slide as sl

    H, W = 2000, 3001
    pitch = ops.DeviceSlide.pitch_for(W)
    rng = np.random.default_rng(0)
    rects = []
    for _ in range(12):
        y0, x0 = int(rng.integers(-50, H)), int(rng.integers(-50, W))
        rects.append((y0, y0 + int(rng.integers(1, 900)), x0, x0 + int(rng.integers(1, 900))))
    spans = sl.tile_spans(rects, H, W, pitch, tile=256)
    cov = np.zeros((H, pitch), dtype=np.int32)
    for y0, y1, b0, b1 in spans:
        assert 0 <= y0 < y1 <= H and 0 <= b0 < b1 <= pitch and b0 % 16 == 0 and (b1 % 16 == 0 or b1 == pitch)
        cov[y0:y1, b0:b1] += 1
    assert cov.max() == 1
    for y0, y1, x0, x1 in rects:
        y0, y1, x0, x1 = max(0, y0), min(H, y1), max(0, x0), min(W, x1)
        assert (cov[y0:y1, 3 * x0 : 3 * x1] == 1).all()
    assert len(sl.tile_spans([], H, W, pitch)) == 0


def test_dense_count_and_band_plans_property():
    """Host-only arithmetic against the oracle's restatement of the reference enumeration (full_samplers.py:374-404) on random
    shapes: dh_dense_count (C), bands.dense_grid (Python) and the oracle agree; every rank's band plan lists exactly the patches
    whose footprint touches its rows; dh_stitch_binned_scratch_bytes is monotone in the list length."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from deephisto_b200 import bands, ops
    from oracle import dense as odense

    lib = _lib.load()

    @settings(max_examples=60, deadline=None)
    @given(ps=st.sampled_from([32, 64, 100, 224]), extra_h=st.integers(0, 700), extra_w=st.integers(0, 700), stride=st.integers(8, 300),
           B=st.integers(1, 70), d=st.sampled_from([1, 3, 4, 16]), world=st.integers(1, 5))
    def check(ps, extra_h, extra_w, stride, B, d, world):
        H, W = ps + extra_h, ps + extra_w
        coords, n = odense.dense_coords(H, W, ps, stride, B)
        N, npad = ops.dense_count(H, W, ps, stride, B)
        g = bands.dense_grid(H, W, ps, stride, B)
        assert (N, npad) == (n, len(coords)) == (g.N, g.n_padded)
        for i in (0, N - 1, npad - 1, N // 2):
            assert bands.patch_origin(g, i) == tuple(int(v) for v in coords[i])
        dh = H // d
        covered = set()
        for rank in range(world):
            plan = bands.plan_band(H, W, ps, stride, d, B, rank, world)
            idx = set(bands.patch_indices(plan))
            if plan.row_end <= plan.row_begin:                      # world > dh: this rank owns no map rows, so no patches either
                assert idx == set() and plan.n_patches == 0
                continue
            touching = {i for i, (y, _x) in enumerate(coords.tolist()) if y // d < plan.row_end and min((y + ps) // d, dh) > plan.row_begin}
            assert touching <= idx, (rank, sorted(touching - idx)[:5])
            covered |= idx
        if dh > 0:
            assert {i for i, (y, _x) in enumerate(coords.tolist()) if min((y + ps) // d, dh) > y // d} <= covered
        a = lib.dh_stitch_binned_scratch_bytes(npad, ps, d, 5, max(dh, 1), max(W // d, 1))
        b = lib.dh_stitch_binned_scratch_bytes(2 * npad, ps, d, 5, max(dh, 1), max(W // d, 1))
        assert 0 < a <= b

    check()


def test_anno_utils_host_side():
    """anno.utils pieces that need no GPU (reference anno/utils.py:19-246, 371-408): class descriptions, visualiser parameters,
    patch-accent parsing, preview sizing, legend drawing."""
    from PIL import Image

    from deephisto_b200.anno import utils as au

    d = au.AnnoDescription.with_known_colors({"AT": (245, 119, 34), "BG": (153, 255, 255)})
    assert [c.id for c in d.anno_classes] == [0, 1] and d.color_by_label("BG") == (153, 255, 255)
    auto = au.AnnoDescription.with_auto_colors(["a", "b", "c"])
    assert len({c.color for c in auto.anno_classes}) == 3 and all(0 <= v <= 255 for c in auto.anno_classes for v in c.color)
    c = au.AnnoClass(3, "TUM", alternate_labels=("tumor",), description="x", color=(1, 2, 3))
    assert c.label_full == "TUM (tumor)" and "TUM (tumor)" in str(c)
    p = au.AnnoVisualizerParams.default()
    assert (p.fill, p.fill_transparency, p.line_width, p.show_legend, p.legend_placement, p.legend_size) == (True, 0.3, 2, True, "TR", 20)
    assert au.AnnoVisualizerParams.no_legend().show_legend is False
    acc = au.PatchVisAccent.parse("r28_LP_7_x17311_y14066", layer=2, patch_s=224)
    assert (acc.layer, acc.size, acc.x, acc.y, acc.label) == (2, 224, 17311, 14066, "LP")
    v = au.AnnoVisualizer(d)
    assert v._downscale(40000, 30000, None, 2000, False) == 20 and v._downscale(1000, 1000, 0.25, None, False) == 4
    assert v._downscale(500, 400, None, None, False) == 1
    with pytest.raises(RuntimeError, match="too big"):
        v._downscale(100000, 100000, None, None, False)
    assert v._downscale(100000, 100000, None, None, True) == 7
    with pytest.raises(ValueError):
        v._downscale(100, 100, 1.5, None, False)
    img = Image.new("RGB", (320, 200), (10, 10, 10))
    for place in ("TL", "TR", "BL", "BR"):
        v.vis_params = au.AnnoVisualizerParams(True, 0.3, 2, True, place, 12)
        out = v._add_legend(img.copy())
        assert out.size == (320, 200) and out.getcolors(maxcolors=1 << 16) is not None and len(out.getcolors(maxcolors=1 << 16)) > 2
