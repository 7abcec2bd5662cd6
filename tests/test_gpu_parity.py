"""Parity of the sm_100a kernels (called through the C-ABI) against the CPU oracle and the golden vectors
generated from the unmodified reference. Bit-exact for bytes / indices / fp32 sums; 1e-5 for the atomic scatter."""

import hashlib

import numpy as np
import pytest
import torch

from oracle import cover as ocover
from oracle import dense as odense
from oracle import region as oregion
from oracle import stitch as ostitch
from oracle import synth
from oracle.make_golden import DENSE_CASES, PIXEL_CASES, STITCH_CASES, region_polygons

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ops():
    from deephisto_b200 import _lib, ops

    _lib.require_device()  # fails loudly without a B200 / the built library
    return ops


def bits(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).cpu().numpy()
    if t.dtype == torch.float32:
        return t.view(torch.int32).cpu().numpy()
    return t.cpu().numpy()


# ---- synthetic slide -----------------------------------------------------------------------------------
@pytest.mark.parametrize("hw", [(64, 48), (37, 53), (300, 261), (1000, 777)])
def test_synth_slide_matches_oracle(ops, hw):
    H, W = hw
    for seed in (0, 7, (5 << 32) | 11):
        dev = ops.DeviceSlide.synthetic(H, W, seed)
        assert dev.pitch % 16 == 0
        assert np.array_equal(dev.to_numpy(), synth.synth_slide(H, W, seed))


# ---- A1 dense coordinates --------------------------------------------------------------------------------
@pytest.mark.parametrize("case", DENSE_CASES)
def test_dense_coords_bit_exact(ops, golden, case):
    z, man = golden
    H, W, ps, stride, B = case
    got = ops.dense_coords(H, W, ps, stride, B).cpu().numpy()
    assert np.array_equal(got, z[f"dense_{H}x{W}_ps{ps}_s{stride}_b{B}_coords"])
    a, b = len(got) // 3, len(got) - len(got) // 4
    part = ops.dense_coords(H, W, ps, stride, B, first=a, count=b - a).cpu().numpy()
    assert np.array_equal(part, got[a:b])


def test_dense_coords_full_size(ops, golden):
    _, man = golden
    for H, W, ps, stride, B in [(40000, 40000, 224, 112, 64), (100000, 100000, 224, 112, 64)]:
        got = ops.dense_coords(H, W, ps, stride, B).cpu().numpy()
        e = man[f"dense_{H}x{W}_ps{ps}_s{stride}_b{B}"]
        assert len(got) == e["n_padded"] and sha(got) == e["coords_sha256"]


# ---- A2 gather + normalise ----------------------------------------------------------------------------------
@pytest.mark.parametrize("case", PIXEL_CASES)
def test_gather_matches_reference_golden(ops, golden, case):
    """fp32 NHWC batches == the tensors the unmodified FullImageDenseSampler.generator_torch() yielded."""
    z, man = golden
    H, W, ps, stride, B = case
    key = f"dense_{H}x{W}_ps{ps}_s{stride}_b{B}"
    slide = ops.DeviceSlide.synthetic(H, W, 0)
    coords = ops.dense_coords(H, W, ps, stride, B)
    want = man[key]["batch_sha256"]
    # one launch over all patches, digested per reference batch
    feats = ops.gather_normalize(slide, coords, ps).cpu().numpy()
    for i, dg in enumerate(want):
        assert sha(feats[i * B : (i + 1) * B]) == dg, f"batch {i}"
    csum = feats.reshape(-1, 3).sum(axis=0, dtype=np.float64)
    assert np.allclose(csum, man[key]["channel_sum_f64"], rtol=1e-9)  # same values (digests), different summation order


@pytest.mark.parametrize("ps", [224, 32, 20, 7])          # 20: ps%4==0 only; 7: generic scalar kernel
@pytest.mark.parametrize("layout", ["NHWC", "NCHW"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.uint8])
def test_gather_variants_vs_oracle(ops, ps, layout, dtype):
    H, W = 523, 611  # odd width -> padded pitch
    host = synth.synth_slide(H, W, 3)
    slide = ops.DeviceSlide.from_numpy(host)
    rng = np.random.default_rng(ps)
    B = 19
    coords = np.stack([rng.integers(0, H - ps + 1, B), rng.integers(0, W - ps + 1, B)], 1).astype(np.int32)
    coords[0] = (0, 0)
    coords[1] = (H - ps, W - ps)
    coords[2] = (-3, 5)              # partly outside: zero fill
    coords[3] = (H - ps + 4, W - 5)
    cdev = torch.from_numpy(coords).cuda()
    raw = odense.gather(host, coords, ps)
    if dtype == torch.uint8:
        got = ops.gather_normalize(slide, cdev, ps, dtype=dtype, layout=layout)
        want = raw if layout == "NHWC" else np.ascontiguousarray(raw.transpose(0, 3, 1, 2))
        assert np.array_equal(got.cpu().numpy(), want)
        return
    flips = torch.from_numpy((np.arange(B) % 4).astype(np.uint8)).cuda()
    for scale255, mean, std, flip in [(True, None, None, None), (False, None, None, None),
                                      (True, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225), None), (True, None, None, flips)]:
        got = ops.gather_normalize(slide, cdev, ps, dtype=dtype, layout=layout, scale255=scale255, mean=mean, std=std, flip=flip)
        want = odense.normalize(raw, scale255, mean, std, layout, None if flip is None else flip.cpu().numpy())
        want_t = torch.from_numpy(want).to(dtype)  # bf16: round-to-nearest-even of the fp32 value
        assert got.shape == want_t.shape
        assert np.array_equal(bits(got), bits(want_t)), (scale255, mean, flip is not None)


def test_gather_out_index_and_value_range(ops):
    H, W, ps = 400, 400, 64
    slide = ops.DeviceSlide.synthetic(H, W, 1)
    host = synth.synth_slide(H, W, 1)
    coords = torch.tensor([[0, 0], [10, 20], [300, 336], [77, 5]], dtype=torch.int32, device="cuda")
    out = torch.full((6, ps, ps, 3), -1.0, device="cuda")
    idx = torch.tensor([5, 0, 3, 1], dtype=torch.int32, device="cuda")
    ops.gather_normalize(slide, coords, ps, out=out, out_index=idx)
    want = odense.normalize(odense.gather(host, coords.cpu().numpy(), ps))
    got = out.cpu().numpy()
    for b, s in enumerate([5, 0, 3, 1]):
        assert np.array_equal(got[s], want[b])
    assert (got[2] == -1).all() and (got[4] == -1).all()
    # all 256 byte values: exact IEEE division on the device
    ramp = np.arange(256, dtype=np.uint8).repeat(3).reshape(1, 256, 3).repeat(4, 0)
    rs = ops.DeviceSlide.from_numpy(np.ascontiguousarray(ramp))
    f = ops.gather_normalize(rs, torch.zeros((1, 2), dtype=torch.int32, device="cuda"), 4)
    g = ops.gather_normalize(rs, torch.tensor([[0, 252]], dtype=torch.int32, device="cuda"), 4)
    assert np.array_equal(f.cpu().numpy()[0, 0, :, 0], (np.arange(4, dtype=np.float32) / np.float32(255)))
    assert np.array_equal(g.cpu().numpy()[0, 0, :, 0], (np.arange(252, 256).astype(np.float32) / np.float32(255)))
    full = ops.gather_normalize(rs, torch.tensor([[0, 4 * i] for i in range(64)], dtype=torch.int32, device="cuda"), 4).cpu().numpy()
    assert np.array_equal(full[:, 0, :, 1].reshape(-1), np.arange(256).astype(np.float32) / np.float32(255))


def test_gather_full_size_properties(ops):
    """BASELINE C1 size (8192^2, stride 224): dense non-overlapping patches tile the slide, so the uint8 gather is a
    permutation of the covered pixels and re-assembling it reproduces the slide exactly."""
    H = W = 8192
    ps = 224
    slide = ops.DeviceSlide.synthetic(H, W, 0)
    coords = ops.dense_coords(H, W, ps, ps, 16)
    n, _ = ops.dense_count(H, W, ps, ps, 16)
    raw = ops.gather_normalize(slide, coords[:n], ps, dtype=torch.uint8)
    host = slide.storage.view(H, slide.pitch)[:, : 3 * W].view(H, W, 3)
    ny = nx = 36
    main = raw[: ny * nx].view(ny, nx, ps, ps, 3).permute(0, 2, 1, 3, 4).reshape(ny * ps, nx * ps, 3)
    assert torch.equal(main, host[: ny * ps, : nx * ps])
    assert torch.equal(raw[n - 1], host[H - ps :, W - ps :])
    f32 = ops.gather_normalize(slide, coords[:n], ps)
    bf = ops.gather_normalize(slide, coords[:n], ps, dtype=torch.bfloat16, layout="NCHW")
    # torch's CUDA `x / 255` multiplies by a rounded reciprocal (not IEEE division): build the expected values on the CPU
    lut = (torch.arange(256, dtype=torch.float32) / 255).cuda()
    want = lut[raw.long()]
    assert torch.equal(f32, want)
    assert torch.equal(bf, want.permute(0, 3, 1, 2).to(torch.bfloat16))


# ---- A4 stitch -------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", STITCH_CASES)
def test_stitch_dense_bit_exact(ops, golden, case):
    z, man = golden
    H, W, ps, stride, B, n, ds = case
    key = f"stitch_{H}x{W}_ps{ps}_s{stride}_b{B}_n{n}"
    logits = torch.from_numpy(z[key + "_logits"]).cuda()
    coords, _ = odense.dense_coords(H, W, ps, stride, B)
    for d in ds:
        s, cnt, am = ops.stitch_dense(logits, H, W, ps, stride, d, B, want_count=True, want_argmax=True)
        e = man[f"{key}_d{d}"]
        assert sha(s.cpu().numpy()) == e["sum_sha256"], d
        assert sha(am.cpu().numpy()) == e["argmax_sha256"], d
        _, ocnt, _ = ostitch.stitch(z[key + "_logits"], coords, H, W, ps, d)
        assert np.array_equal(cnt.cpu().numpy().astype(np.int64), ocnt)
        # argmax-only and band modes
        _, _, am2 = ops.stitch_dense(logits, H, W, ps, stride, d, B, want_sum=False, want_argmax=True)
        assert torch.equal(am, am2)
        dh = H // d
        r0, r1 = dh // 3, dh - dh // 5
        sb, cb, ab = ops.stitch_dense(logits, H, W, ps, stride, d, B, row_begin=r0, row_end=r1, want_count=True, want_argmax=True)
        assert torch.equal(sb, s[r0:r1]) and torch.equal(cb, cnt[r0:r1]) and torch.equal(ab, am[r0:r1])
        # finalize: argmax of the sum map (first maximum) and count-normalised map
        norm, am3 = ops.stitch_finalize(s, cnt, want_norm=True)
        assert torch.equal(am3, am)
        assert np.array_equal(norm.cpu().numpy(), ostitch.normalize(s.cpu().numpy(), ocnt))


def test_stitch_scatter_vs_oracle(ops):
    H, W, ps, d, n = 1500, 1300, 224, 16, 5
    rng = np.random.default_rng(0)
    P = 300
    coords = np.stack([rng.integers(0, H - ps + 1, P), rng.integers(0, W - ps + 1, P)], 1).astype(np.int32)
    logits = rng.standard_normal((P, n)).astype(np.float32)
    want, wcnt, _ = ostitch.stitch(logits, coords, H, W, ps, d)
    s = torch.zeros((H // d, W // d, n), device="cuda")
    c = torch.zeros((H // d, W // d), dtype=torch.int32, device="cuda")
    ops.stitch_scatter(torch.from_numpy(logits).cuda(), torch.from_numpy(coords).cuda(), ps, d, s, c)
    got = s.cpu().numpy()
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-5 * scale  # fp32 accumulation order only
    assert np.array_equal(c.cpu().numpy().astype(np.int64), wcnt)
    # band: rows [20, 60)
    sb = torch.zeros((40, W // d, n), device="cuda")
    ops.stitch_scatter(torch.from_numpy(logits).cuda(), torch.from_numpy(coords).cuda(), ps, d, sb, None, row_offset=20)
    assert np.abs(sb.cpu().numpy() - want[20:60]).max() <= 1e-5 * scale


@pytest.mark.parametrize("case", [
    # H, W, ps, d, n, P, overhang
    (1500, 1300, 224, 16, 5, 300, False),     # the reference's visualisation setting
    (1500, 1300, 224, 4, 5, 200, False),      # dw * n % 4 == 0 -> 16-byte stores
    (900, 1001, 224, 1, 5, 40, True),         # d = 1, dw * n % 4 != 0 -> scalar stores, patches overhanging the right/bottom edge
    (1000, 777, 100, 7, 3, 150, True),        # ps % d != 0: footprints of 14 or 15 cells
    (640, 640, 64, 2, 1, 500, False),         # one class
    (800, 800, 224, 8, 8, 120, False),        # n = 8: the most the cell kernel keeps in registers
    (800, 800, 224, 8, 11, 120, False),       # n > 8: class map through the sum map
])
def test_stitch_binned_bit_exact_vs_oracle(ops, case):
    """dh_stitch_binned == the reference loop over an arbitrary coordinate list in list order: bit-identical sums, counts, argmax."""
    H, W, ps, d, n, P, overhang = case
    rng = np.random.default_rng(P + d)
    ymax, xmax = (H - 1, W - 1) if overhang else (H - ps, W - ps)
    coords = np.stack([rng.integers(0, ymax + 1, P), rng.integers(0, xmax + 1, P)], 1).astype(np.int32)
    logits = (rng.standard_normal((P, n)) * 3).astype(np.float32)
    want, wcnt, wam = ostitch.stitch(logits, coords, H, W, ps, d)
    dh, dw = H // d, W // d
    lg, cd = torch.from_numpy(logits).cuda(), torch.from_numpy(coords).cuda()
    s, c, am = ops.stitch_binned(lg, cd, ps, d, dh, dw, want_count=True, want_argmax=True)
    assert np.array_equal(bits(s), want.view(np.int32))
    assert np.array_equal(c.cpu().numpy().astype(np.int64), wcnt)
    assert np.array_equal(am.cpu().numpy().astype(np.int64), wam)
    # outputs one at a time (different kernels / tile geometries) and a row band
    _, _, am1 = ops.stitch_binned(lg, cd, ps, d, dh, dw, want_sum=n > 8, want_argmax=True)
    _, c1, _ = ops.stitch_binned(lg, cd, ps, d, dh, dw, want_sum=False, want_count=True)
    assert torch.equal(am1, am) and torch.equal(c1, c)
    r0, r1 = dh // 3, dh // 3 + max(1, dh // 2)
    sb, cb, ab = ops.stitch_binned(lg, cd, ps, d, r1 - r0, dw, row_offset=r0, want_count=True, want_argmax=True)
    assert torch.equal(sb, s[r0:r1]) and torch.equal(cb, c[r0:r1]) and torch.equal(ab, am[r0:r1])
    # the atomic scatter agrees to rounding
    sc = torch.zeros_like(s)
    ops.stitch_scatter(lg, cd, ps, d, sc, None)
    assert float((sc - s).abs().max()) <= 1e-5 * float(np.abs(want).max())


def test_stitch_binned_long_lists_and_tile_heights(ops):
    """More patches over one tile than the shared-memory staging holds (global-memory path), duplicates of one origin included;
    every tile height gives the same bits."""
    H, W, ps, d, n = 1024, 1024, 224, 16, 5
    rng = np.random.default_rng(7)
    P = 900
    coords = np.stack([rng.integers(300, 340, P), rng.integers(300, 340, P)], 1).astype(np.int32)   # ~900 patches over the same tiles
    coords[100:200] = coords[100]
    logits = rng.standard_normal((P, n)).astype(np.float32)
    want, wcnt, wam = ostitch.stitch(logits, coords, H, W, ps, d)
    lg, cd = torch.from_numpy(logits).cuda(), torch.from_numpy(coords).cuda()
    from deephisto_b200 import _lib
    lib = _lib.require_device()
    try:
        for th in (0, 1, 5, 16, 64, 1000):
            lib.dh_stitch_binned_set_tile_rows(th)
            s, c, am = ops.stitch_binned(lg, cd, ps, d, H // d, W // d, want_count=True, want_argmax=True)
            assert np.array_equal(bits(s), want.view(np.int32)), th
            assert np.array_equal(c.cpu().numpy().astype(np.int64), wcnt) and np.array_equal(am.cpu().numpy().astype(np.int64), wam), th
    finally:
        lib.dh_stitch_binned_set_tile_rows(0)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8])
def test_stitch_binned_cell_lane_kernel_every_class_count(ops, n):
    """bin_cell_sum_kernel (one cell per lane, N class sums in registers; one instantiation per class count 1..8, rows 16-byte aligned
    or not): bit-identical to the reference loop on a crowded list (every tile sees many patches, some tiles more than the staging
    holds), on every tile height, and on a row band."""
    from deephisto_b200 import _lib
    lib = _lib.require_device()
    rng = np.random.default_rng(100 + n)
    ps = 224
    try:
        for (H, W, d, P, hot) in ((1500, 1300, 4, 500, 0), (1499, 1203, 4, 400, 0), (1100, 1037, 2, 200, 0), (1024, 1024, 16, 900, 700), (900, 1001, 8, 300, 200)):
            coords = np.stack([rng.integers(0, H - ps + 1, P), rng.integers(0, W - ps + 1, P)], 1).astype(np.int32)
            if hot:
                coords[:hot] = np.stack([rng.integers(300, 340, hot), rng.integers(300, 340, hot)], 1)   # more patches over one tile than fit the staging
            coords[5:9] = coords[4]
            logits = (rng.standard_normal((P, n)) * 3).astype(np.float32)
            want, _, _ = ostitch.stitch(logits, coords, H, W, ps, d)
            dh, dw = H // d, W // d
            lg, cd = torch.from_numpy(logits).cuda(), torch.from_numpy(coords).cuda()
            for th in (0, 16, 64, 128):
                lib.dh_stitch_binned_set_variant(3)
                lib.dh_stitch_binned_set_tile_rows(th)
                s, _, _ = ops.stitch_binned(lg, cd, ps, d, dh, dw)
                assert np.array_equal(bits(s), want.view(np.int32)), (H, W, d, n, th)
            lib.dh_stitch_binned_set_tile_rows(0)
            r0, r1 = dh // 3, dh // 3 + dh // 2
            sb, _, _ = ops.stitch_binned(lg, cd, ps, d, r1 - r0, dw, row_offset=r0)
            assert torch.equal(sb, s[r0:r1]), (H, W, d, n)
    finally:
        lib.dh_stitch_binned_set_variant(0)
        lib.dh_stitch_binned_set_tile_rows(0)


def test_stitch_binned_equals_dense_stitch_on_the_dense_enumeration(ops):
    """Fed with the dense sampler's own (padded) coordinate list, the binned stitch reproduces dh_stitch_dense bit for bit --
    two independent implementations of the reference order (predict_full_patched.py:47-54 over full_samplers.py:374-404)."""
    H, W, ps, stride, B, n = 6000, 5000, 224, 112, 64, 5
    N, npad = ops.dense_count(H, W, ps, stride, B)
    coords = ops.dense_coords(H, W, ps, stride, B)
    lg = torch.randn((npad, n), generator=torch.Generator(device="cuda").manual_seed(1), device="cuda")
    for d in (16, 4, 1):
        s, c, am = ops.stitch_dense(lg, H, W, ps, stride, d, B, want_count=True, want_argmax=True)
        s2, c2, am2 = ops.stitch_binned(lg, coords, ps, d, H // d, W // d, want_count=True, want_argmax=True)
        assert torch.equal(s.view(torch.int32), s2.view(torch.int32)) and torch.equal(c, c2) and torch.equal(am, am2)
        del s, s2
        torch.cuda.empty_cache()


# ---- B coverage sampler -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [(700, 900, 224, 16, 3), (2048, 2048, 224, 64, 0), (300, 5000, 224, 7, 11)])
def test_cover_sampler_bit_exact_vs_oracle(ops, cfg):
    H, W, ps, B, seed = cfg
    st = ops.CoverState(H, W, ps, 16, 2, B, seed)
    ref = ocover.CoverSampler(H, W, ps, B, seed)
    filled = 0.0
    it = 0
    while filled < 1:
        c, nz = st.next_coords()
        rc, filled = ref.next_coords()
        assert np.array_equal(c.cpu().numpy(), rc), f"batch {it}"
        assert int(nz.item()) / (st.dh * st.dw) == filled
        it += 1
        assert it < 2000
    assert np.array_equal(st.accum.cpu().numpy().astype(np.int64), ref.accum)


@pytest.mark.parametrize("ps", [224, 64, 30, 100])
def test_gather_space_to_depth_layout(ops, ps):
    """DH_S2D16 (the predictor's stem input): the kernel's output equals the 2x2 space-to-depth fold of its own NHWC bf16 output,
    channel p*8 + q*3 + c, channels 6/7/14/15 zero, a zero border of 2 / 1 pixels left untouched -- for interior patches (bulk-copy
    path), flipped patches and patches that overhang the slide (guarded path), with and without mean / std."""
    H, W, B = 700, 900, 37
    slide = ops.DeviceSlide.synthetic(H, W, seed=3)
    rng = np.random.default_rng(ps)
    yx = np.stack([rng.integers(0, H - ps + 1, B), rng.integers(0, W - ps + 1, B)], 1)
    yx[0] = (-5, 10)
    yx[1] = (H - ps + 7, W - ps + 3)
    yx[2] = (0, 0)
    yx[3] = (H - ps, W - ps)
    coords = torch.from_numpy(yx.astype(np.int32)).cuda()
    flip = torch.from_numpy(rng.integers(0, 4, B).astype(np.uint8)).cuda()
    side = ps // 2 + 3
    for kw in (dict(), dict(flip=flip), dict(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)), dict(scale255=False)):
        nhwc = ops.gather_normalize(slide, coords, ps, dtype=torch.bfloat16, layout="NHWC", **kw)
        out = torch.full((B, side, side, 16), 7.0, dtype=torch.bfloat16, device="cuda")             # canary: the border must stay untouched
        got = ops.gather_normalize(slide, coords, ps, dtype=torch.bfloat16, layout="S2D16", out=out, **kw)
        assert got.data_ptr() == out.data_ptr()
        want = torch.full_like(out, 7.0)
        v = nhwc.reshape(B, ps // 2, 2, ps // 2, 2, 3).permute(0, 1, 3, 2, 4, 5)                     # [B, y', x', p, q, c]
        inner = want[:, 2 : 2 + ps // 2, 2 : 2 + ps // 2]
        inner[...] = 0
        inner[..., 0:6] = v[:, :, :, 0].reshape(B, ps // 2, ps // 2, 6)
        inner[..., 8:14] = v[:, :, :, 1].reshape(B, ps // 2, ps // 2, 6)
        assert torch.equal(got.view(torch.int16), want.view(torch.int16)), kw.keys()
        if ps % 4 == 0:                                            # DH_S2D48: 4x4 blocks, channel p*12 + q*3 + c, same bytes as NHWC reordered
            got48 = ops.gather_normalize(slide, coords, ps, dtype=torch.bfloat16, layout="S2D48", **kw)
            want48 = nhwc.reshape(B, ps // 4, 4, ps // 4, 4, 3).permute(0, 1, 3, 2, 4, 5).reshape(B, ps // 4, ps // 4, 48)
            assert got48.shape == want48.shape and torch.equal(got48.view(torch.int16), want48.contiguous().view(torch.int16)), kw.keys()
    fresh = ops.gather_normalize(slide, coords, ps, dtype=torch.bfloat16, layout="S2D16")              # allocated here: zero border
    assert float(fresh[:, :2].abs().max()) == 0 and float(fresh[:, -1].abs().max()) == 0 and float(fresh[:, :, :2].abs().max()) == 0
    with pytest.raises(Exception):
        ops.gather_normalize(slide, coords, ps, dtype=torch.float32, layout="S2D16")


@pytest.mark.parametrize("cfg", [(1500, 1300, 224, 32, 5), (3000, 2600, 224, 1024, 1), (640, 800, 64, 200, 9)])
def test_cover_sampler_group_launch_equals_single_launches(ops, cfg):
    """dh_cover_sample_group (ONE persistent launch for a group of batches, count tree in shared memory) yields exactly the batches
    of one launch per batch, with either kernel variant (dh_cover_set_variant 1 = the per-batch kernel with a full scan), up to and
    beyond full coverage (batches after it are no-ops); accumulators equal."""
    from deephisto_b200 import _lib

    H, W, ps, B, seed = cfg
    lib = _lib.require_device()
    cells = (H // 16) * (W // 16)
    a = ops.CoverState(H, W, ps, 16, 2, B, seed)
    c = ops.CoverState(H, W, ps, 16, 2, B, seed)
    done, groups = False, 0
    while not done:
        n = 7 if groups % 2 else 16
        ga, na = a.next_group(n)
        try:
            lib.dh_cover_set_variant(1)
            gc, nc = c.next_group(n)
        finally:
            lib.dh_cover_set_variant(0)
        na = na.cpu().tolist()
        for i in range(n):
            live = i == 0 or na[i - 1] < cells                # batches enqueued after full coverage leave their (zeroed) slots untouched
            assert na[i] == int(nc[i].item())
            if live:
                assert torch.equal(ga[i], gc[i]), (groups, i)
        done = na[-1] >= cells
        groups += 1
        assert groups < 400
    assert torch.equal(a.accum, c.accum) and int(a.accum.min()) >= 1
    # one launch per batch through dh_cover_sample against a grouped run
    d = ops.CoverState(H, W, ps, 16, 2, B, seed)
    e = ops.CoverState(H, W, ps, 16, 2, B, seed)
    g1, n1 = d.next_group(5)
    prev = 0
    for i in range(5):
        ci, ni = e.next_coords(stop_when_full=True)
        assert int(ni.item()) == int(n1[i].item())
        if prev < cells:                                      # a launch after full coverage leaves its output untouched
            assert torch.equal(ci, g1[i])
        prev = int(ni.item())


# ---- C/D regions ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(region_polygons().keys()))
def test_region_accept_dense_bit_exact(ops, golden, name):
    from deephisto_b200 import geometry

    z, _ = golden
    verts = z[f"region_{name}_verts"].astype(np.float64)
    for layer in (1, 2):
        v = oregion.scale_vertices(verts, layer)
        edges_h = geometry.build_edges(v)
        assert np.array_equal(edges_h, oregion.build_edges(v))
        edges = torch.from_numpy(edges_h.reshape(-1)).cuda()
        for ps, stride, ri in ((224, 112, 0.75), (224, 56, 0.5), (64, 32, 0.95)):
            y0, x0, ny, nx = oregion.dense_candidates(v, (2048 // layer, 2048 // layer), ps, stride)
            want_c, want_m, want_a = oregion.coords_dense(v, (2048 // layer, 2048 // layer), ps, stride, ri)
            mask, area = ops.region_accept_dense(edges, 0, len(edges_h), y0, x0, ny, nx, stride, ps, ps * ps * ri, want_area=True)
            assert np.array_equal(area.cpu().numpy(), want_a)     # every float64 op reproduced
            assert np.array_equal(mask.cpu().numpy(), want_m)
            got = ops.compact_coords(mask, y0, x0, ny, nx, stride).cpu().numpy()
            assert np.array_equal(got, z[f"region_{name}_l{layer}_ps{ps}_s{stride}_ri{ri}"])


def _tables(polys_per_image, hw, one_image, device="cuda"):
    from deephisto_b200.patch_samplers.region_samplers import build_tables

    return build_tables([(hw, p) for p in polys_per_image], layer=1, area_influence=0.5, classes=None, one_image_for_batch=one_image,
                        device=device)


@pytest.mark.parametrize("one_image", [True, False])
def test_region_sample_bit_exact_vs_oracle(ops, one_image):
    hw = (6000, 6000)
    imgs = [synth.synth_polygons(12, *hw, seed=1, rmin=300, rmax=900), synth.synth_polygons(7, *hw, seed=2, rmin=200, rmax=700, n_classes=3)]
    tables, regions, classes = _tables(imgs, hw, one_image)
    rs = oregion.RegionSet([(hw, p) for p in imgs], layer=1, one_image_for_batch=one_image)
    assert classes == rs.classes
    for (n_slots, k, ri, seed, off, spt) in [(256, 4, 0.75, 9, 0, 128), (70, 3, 0.9, 1, 1000, 35), (33, 32, 0.5, 2, 7, 33)]:
        c, lab, img, st = ops.region_sample(tables.struct, n_slots, k, 224, 224 * 224 * ri, seed=seed, slot_offset=off, slots_per_table_draw=spt)
        oc, olab, oimg, ost = oregion.sample(rs, n_slots, k, 224, ri, seed=seed, slot_offset=off, slots_per_table_draw=spt)
        assert np.array_equal(st.cpu().numpy(), ost) and (ost == 0).all()
        assert np.array_equal(c.cpu().numpy(), oc)
        assert np.array_equal(lab.cpu().numpy(), olab)
        assert np.array_equal(img.cpu().numpy(), oimg)


def test_region_sample_failure_status(ops):
    hw = (3000, 3000)
    tiny = [{"class": "A", "vertices": [[10.0, 10.0], [100.0, 10.0], [100.0, 100.0], [10.0, 100.0]]}]  # area < ps*ps*ri
    tables, _, _ = _tables([tiny], hw, False)
    c, lab, img, st = ops.region_sample(tables.struct, 8, 4, 224, 224 * 224 * 0.75, max_redraw=3)
    assert (st.cpu().numpy() == 1).all() and (lab.cpu().numpy() == -1).all()


def test_rasterize_vs_oracle(ops):
    from deephisto_b200 import geometry

    polys = [np.asarray(p["vertices"]) for p in synth.synth_polygons(6, 4000, 4000, seed=4, rmin=300, rmax=900)]
    edges = [geometry.build_edges(v) for v in polys]
    bbox = [geometry.polygon_bounds(v) for v in polys]
    off = np.zeros(len(edges) + 1, np.int32)
    off[1:] = np.cumsum([len(e) for e in edges])
    lab = ops.rasterize_polygons(torch.from_numpy(np.concatenate(edges).reshape(-1)).cuda(), torch.from_numpy(off).cuda(),
                                 torch.from_numpy(np.asarray(bbox).reshape(-1)).cuda(), 16.0, 250, 250)
    want = oregion.rasterize(edges, bbox, 16.0, 250, 250)
    assert np.array_equal(lab.cpu().numpy(), want)
    assert len(np.unique(want)) > 2


@pytest.mark.parametrize("variant", ["direct", "tma"])
@pytest.mark.parametrize("ps", [224, 64, 16, 256, 112, 20])
def test_gather_kernel_variants_agree_with_oracle(ops, variant, ps):
    """Both gather kernels (direct LDG/STG and TMA-staged) against the oracle, all layouts / dtypes, flips, out-of-slide patches."""
    H, W = 700, 900
    host = synth.synth_slide(H, W, 11)
    slide = ops.DeviceSlide.from_numpy(host)
    rng = np.random.default_rng(ps)
    B = 37
    coords = np.stack([rng.integers(0, H - ps + 1, B), rng.integers(0, W - ps + 1, B)], 1).astype(np.int32)
    coords[0], coords[1], coords[2], coords[3] = (0, 0), (H - ps, W - ps), (-5, 7), (H - ps + 9, W - 3)
    cdev = torch.from_numpy(coords).cuda()
    raw = odense.gather(host, coords, ps)
    flips = torch.from_numpy(rng.integers(0, 4, B).astype(np.uint8)).cuda()
    ops.set_gather_variant(variant)
    try:
        for layout in ("NHWC", "NCHW"):
            for dtype in (torch.float32, torch.bfloat16):
                if variant == "tma" and dtype == torch.bfloat16 and ps % 8 != 0:
                    with pytest.raises(Exception, match="not supported by the TMA-staged kernel"):   # 16-byte bf16 units need ps % 8 == 0
                        ops.gather_normalize(slide, cdev, ps, dtype=dtype, layout=layout)
                    continue
                for scale255, mean, std, flip in [(True, None, None, None), (False, None, None, None),
                                                  (True, (0.5, 0.4, 0.3), (0.2, 0.25, 0.3), None), (True, None, None, flips)]:
                    got = ops.gather_normalize(slide, cdev, ps, dtype=dtype, layout=layout, scale255=scale255, mean=mean, std=std, flip=flip)
                    want = odense.normalize(raw, scale255, mean, std, layout, None if flip is None else flip.cpu().numpy())
                    assert np.array_equal(bits(got), bits(torch.from_numpy(want).to(dtype))), (layout, dtype, scale255, mean, flip is not None)
    finally:
        ops.set_gather_variant("auto")


@pytest.mark.parametrize("variant", [0, 1])
def test_stitch_dense_random_shapes_vs_oracle(ops, variant):
    """Randomised dense enumerations (slide, patch, stride, batch padding, class counts 1..64, downscales below and above the stride,
    row bands): dh_stitch_dense is bit-identical to the reference loop over the reference's own coordinate list -- with the logits
    staged per block and summed per column class (variant 0; strides too fine for the shared-memory budget fall back inside the same
    kernel) and with the per-row-class reads (variant 1)."""
    from deephisto_b200 import _lib

    lib = _lib.require_device()
    lib.dh_stitch_dense_set_variant(variant)
    try:
        rng = np.random.default_rng(77)
        for trial in range(28):
            ps = int(rng.choice([32, 48, 64, 100, 224]))
            stride = int(rng.choice([3, 8, ps // 4, ps // 2, ps, ps + 16]))
            d = int(rng.choice([1, 2, 3, 4, 7, 16, 32, 64]))
            d = min(d, ps)
            n = int(rng.choice([1, 2, 3, 5, 8, 9, 16, 64]))
            B = int(rng.choice([1, 7, 64]))
            H, W = int(rng.integers(ps, 700)), int(rng.integers(ps, 700))
            if trial % 5 == 0:
                H = ps + stride * int(rng.integers(0, 4))                  # (H - ps) % stride == 0: the last row coincides with nothing, edge of the enumeration
            while ((H - ps) // stride + 2) * ((W - ps) // stride + 2) > 9000:
                stride *= 2
            coords, N = odense.dense_coords(H, W, ps, stride, B)
            logits = (rng.standard_normal((coords.shape[0], n)) * 2).astype(np.float32)
            want, wcnt, wam = ostitch.stitch(logits, coords, H, W, ps, d)
            dh, dw = H // d, W // d
            if dh == 0 or dw == 0:
                continue
            lg = torch.from_numpy(logits).cuda()
            s, c, am = ops.stitch_dense(lg, H, W, ps, stride, d, B, want_count=True, want_argmax=True)
            tag = (trial, H, W, ps, stride, d, n, B)
            assert np.array_equal(bits(s), want.view(np.int32)), tag
            assert np.array_equal(c.cpu().numpy().astype(np.int64), wcnt), tag
            assert np.array_equal(am.cpu().numpy().astype(np.int64), wam), tag
            _, _, am1 = ops.stitch_dense(lg, H, W, ps, stride, d, B, want_sum=False, want_argmax=True)
            assert torch.equal(am1, am), tag
            r0 = int(rng.integers(0, dh))
            r1 = int(rng.integers(r0, dh)) + 1
            sb, cb, ab = ops.stitch_dense(lg, H, W, ps, stride, d, B, row_begin=r0, row_end=r1, want_count=True, want_argmax=True)
            assert torch.equal(sb, s[r0:r1]) and torch.equal(cb, c[r0:r1]) and torch.equal(ab, am[r0:r1]), tag
    finally:
        lib.dh_stitch_dense_set_variant(0)


# ---- BASELINE full sizes through size-independent properties ---------------------------------------------------------------
def test_stitch_full_size_properties(ops):
    """40k x 40k / stride 112 (BASELINE configs[2], 127 488 padded patches): with all logits = 1 the sum map IS the count map
    (small integers, exact in fp32); bands concatenate to the full map; total mass = sum of clipped patch footprints."""
    H = W = 40000
    ps, stride, B, n = 224, 112, 64, 5
    N, npad = ops.dense_count(H, W, ps, stride, B)
    ones = torch.ones((npad, n), device="cuda")
    for d in (16, 4):
        s, cnt, am = ops.stitch_dense(ones, H, W, ps, stride, d, B, want_count=True, want_argmax=True)
        dh, dw = H // d, W // d
        assert tuple(s.shape) == (dh, dw, n) and int(am.max()) == 0
        assert torch.equal(s[..., 0], cnt.to(torch.float32)) and torch.equal(s[..., n - 1], s[..., 0])
        assert int(cnt.min()) >= 1                                           # every cell is covered
        coords = ops.dense_coords(H, W, ps, stride, B).cpu().numpy().astype(np.int64)
        fy = np.minimum((coords[:, 0] + ps) // d, dh) - coords[:, 0] // d
        fx = np.minimum((coords[:, 1] + ps) // d, dw) - coords[:, 1] // d
        assert int(cnt.sum(dtype=torch.int64)) == int((fy * fx).sum())
        r0, r1 = dh // 3 + 1, dh // 3 + 1 + 257
        sb, cb, _ = ops.stitch_dense(ones, H, W, ps, stride, d, B, row_begin=r0, row_end=r1, want_count=True)
        assert torch.equal(sb, s[r0:r1]) and torch.equal(cb, cnt[r0:r1])
        # random logits: finalize(argmax) of the sum map == fused argmax
        g = torch.Generator(device="cuda").manual_seed(d)
        lg = torch.randn((npad, n), generator=g, device="cuda")
        s2, _, am2 = ops.stitch_dense(lg, H, W, ps, stride, d, B, want_argmax=True)
        _, am3 = ops.stitch_finalize(s2)
        assert torch.equal(am2, am3)
        del s, cnt, am, s2, sb, cb
        torch.cuda.empty_cache()


def test_region_sampling_full_size_properties(ops):
    """BASELINE configs[1] shape: 32768^2 slide, 50 synthetic polygons, 1000 batches of 256 slots in one launch. Every slot succeeds,
    every origin keeps the patch inside the slide and satisfies the acceptance criterion (oracle clip area on a subsample); classes
    are drawn uniformly (region_samplers.py:555-560); launch splitting does not change the stream."""
    from deephisto_b200.patch_samplers.region_samplers import build_tables
    from deephisto_b200.synthetic import synth_polygons

    H = W = 32768
    ps, k, ri = 224, 4, 0.75
    polys = synth_polygons(50, H, W, seed=0)
    tables, regions, classes = build_tables([((H, W), polys)], layer=1, area_influence=0.5, classes=None, one_image_for_batch=True)
    n = 256 * 1000
    c, lab, img, st = ops.region_sample(tables.struct, n, k, ps, ps * ps * ri, seed=42, slots_per_table_draw=512)
    assert int(st.max()) == 0
    cy, cx = c[:, 0], c[:, 1]
    assert int(cy.min()) >= 0 and int(cx.min()) >= 0 and int(cy.max()) <= H - ps and int(cx.max()) <= W - ps
    hist = torch.bincount(lab, minlength=len(classes)).cpu().numpy()
    assert hist.sum() == n and np.abs(hist / n - 1 / len(classes)).max() < 0.01
    # groups of k slots share one region: consecutive slots of a group carry the same label
    assert torch.equal(lab.view(-1, k)[:, 0].repeat_interleave(k), lab)
    cn, labn = c.cpu().numpy(), lab.cpu().numpy()
    by_class = {ci: [oregion.build_edges(np.asarray(p["vertices"], dtype=np.float64)) for p in polys if p["class"] == cls] for ci, cls in enumerate(classes)}
    for q in range(0, n, 997):
        y, x = cn[q]
        best = max(float(np.ravel(oregion.clip_area(e, float(x), float(y), float(ps)))[0]) for e in by_class[int(labn[q])])
        assert best > ri * ps * ps
    # the same slots drawn by two launches with offsets give the same coordinates
    a, _, _, _ = ops.region_sample(tables.struct, 4096, k, ps, ps * ps * ri, seed=42, slot_offset=0, slots_per_table_draw=512)
    b, _, _, _ = ops.region_sample(tables.struct, 4096, k, ps, ps * ps * ri, seed=42, slot_offset=4096, slots_per_table_draw=512)
    assert torch.equal(torch.cat([a, b]), c[:8192])


def test_cover_sampler_full_size_terminates(ops):
    """8192^2 slide (BASELINE configs[0] size), batch 256: coverage reaches 1.0, every origin is inside the clamp range, the
    accumulator equals the footprint histogram of the yielded coordinates."""
    H = W = 8192
    ps, B, sp = 224, 256, 16
    st = ops.CoverState(H, W, ps, sp, 2, B, seed=7)
    cells = (H // sp) * (W // sp)
    acc = torch.zeros((H // sp, W // sp), dtype=torch.int32, device="cuda")
    filled, batches = 0.0, 0
    while filled < 1.0:
        coords, nonzero = st.next_coords()
        assert int(coords.min()) >= 0 and int(coords[:, 0].max()) <= H - ps and int(coords[:, 1].max()) <= W - ps
        ones = torch.ones((B, 1), device="cuda")
        ops.stitch_scatter(ones, coords, ps, sp, None, acc)                 # the stitcher's footprint == the sampler's (y//s:(y+ps)//s)
        filled = int(nonzero.item()) / cells
        batches += 1
        assert batches < 5000
    assert torch.equal(acc, st.accum)
    assert batches * B >= 2 * cells / ((ps // sp) ** 2) * 0.5 and int(st.accum.min()) >= 1


@pytest.mark.parametrize("layout", ["NHWC", "NCHW"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gather_writes_only_its_output_canary(ops, layout, dtype):
    """compute-sanitizer is closed on this GPU pool, so out-of-bounds writes are checked with canaries: the output is a view in
    the middle of a larger buffer filled with a sentinel, for batch sizes that leave partial tiles / partial waves."""
    H, W, ps = 2000, 1800, 224
    slide = ops.DeviceSlide.synthetic(H, W, 3)
    for B in (1, 5, 149, 593):
        g = torch.Generator(device="cuda").manual_seed(B)
        coords = torch.stack([torch.randint(0, H - ps + 1, (B,), generator=g, device="cuda"), torch.randint(0, W - ps + 1, (B,), generator=g, device="cuda")], 1).to(torch.int32)
        n = B * ps * ps * 3
        pad = 4096
        big = torch.full((n + 2 * pad,), 7.0, dtype=dtype, device="cuda")
        shape = (B, ps, ps, 3) if layout == "NHWC" else (B, 3, ps, ps)
        out = big[pad : pad + n].view(shape)
        flips = torch.randint(0, 4, (B,), generator=g, device="cuda").to(torch.uint8)
        ops.gather_normalize(slide, coords, ps, dtype=dtype, layout=layout, flip=flips, out=out)
        assert bool((big[:pad] == 7.0).all()) and bool((big[pad + n :] == 7.0).all())
        assert float(out.float().max()) <= 1.0                                # every element was written ([0,1] after /255)


@pytest.mark.parametrize("layout,dtype", [("NHWC", torch.float32), ("NCHW", torch.bfloat16), ("NCHW", torch.float32)])
def test_gather_multi_slide_one_launch(ops, layout, dtype):
    """dh_gather_normalize_multi: a batch whose patches come from three resident slides of different sizes (descriptor table in
    HBM) equals the per-slide oracle gather -- flips, overhanging patches (zero fill) and both kernel paths included."""
    ps, B = 64, 61
    shapes = [(500, 700), (333, 401), (900, 260)]
    hosts = [synth.synth_slide(h, w, 40 + i) for i, (h, w) in enumerate(shapes)]
    table = ops.SlideTable([ops.DeviceSlide.from_numpy(a) for a in hosts])
    rng = np.random.default_rng(3)
    images = rng.integers(0, 3, B).astype(np.int32)
    coords = np.stack([[rng.integers(0, shapes[i][0] - ps + 1), rng.integers(0, shapes[i][1] - ps + 1)] for i in images]).astype(np.int32)
    coords[0], images[0] = (-7, 3), 1                      # overhangs slide 1: zero fill
    coords[1], images[1] = (900 - ps + 5, 260 - 9), 2      # overhangs slide 2
    flips = rng.integers(0, 4, B).astype(np.uint8)
    got = ops.gather_normalize_multi(table, torch.from_numpy(images).cuda(), torch.from_numpy(coords).cuda(), ps, dtype=dtype, layout=layout,
                                     flip=torch.from_numpy(flips).cuda())
    want = np.stack([odense.normalize(odense.gather(hosts[images[b]], coords[b : b + 1], ps), True, None, None, layout, flips[b : b + 1])[0]
                     for b in range(B)])
    assert np.array_equal(bits(got), bits(torch.from_numpy(want).to(dtype)))
    with pytest.raises(Exception, match="not supported"):
        ops.gather_normalize_multi(table, torch.from_numpy(images).cuda(), torch.from_numpy(coords).cuda(), 30)


def test_empty_inputs_and_error_reporting(ops):
    """Empty batches are no-ops; invalid arguments come back as negative status codes with a message (no exceptions from C,
    no partial launches)."""
    from deephisto_b200 import _lib

    slide = ops.DeviceSlide.synthetic(300, 300, 0)
    empty = torch.zeros((0, 2), dtype=torch.int32, device="cuda")
    for layout in ("NHWC", "NCHW"):
        out = ops.gather_normalize(slide, empty, 64, layout=layout)
        assert out.numel() == 0
    assert ops.dense_coords(300, 300, 64, 32, 4, first=5, count=0).shape == (0, 2)
    lg = torch.zeros((0, 5), device="cuda")
    sm = torch.zeros((10, 10, 5), device="cuda")
    ops.stitch_scatter(lg, empty, 64, 16, sm, None)                      # P = 0
    z, zc, za = ops.stitch_binned(lg, empty, 64, 16, 8, 8, want_count=True, want_argmax=True)   # P = 0: the reference's map stays zero
    assert int(z.abs().sum()) == 0 and int(zc.sum()) == 0 and int(za.sum()) == 0
    assert float(sm.abs().max()) == 0.0
    with pytest.raises(_lib.DeepHistoError, match="patch size"):
        ops.gather_normalize(slide, torch.zeros((1, 2), dtype=torch.int32, device="cuda"), 8200, dtype=torch.uint8)      # > 8192
    with pytest.raises(ValueError, match="smaller than patch"):
        ops.dense_count(100, 100, 224, 112, 64)
    with pytest.raises(_lib.DeepHistoError, match="rows"):
        ops.stitch_dense(torch.zeros((ops.dense_count(300, 300, 64, 32, 4)[1], 5), device="cuda"), 300, 300, 64, 32, 16, 4, row_begin=5, row_end=99)
    with pytest.raises(_lib.DeepHistoError, match="batch size"):
        ops.CoverState(64, 64, 32, 16, 2, 4096, seed=0).next_coords()
    with pytest.raises(_lib.DeepHistoError, match="must be a CUDA tensor"):
        ops.gather_normalize(slide, torch.zeros((1, 2), dtype=torch.int32), 64)
    with pytest.raises(TypeError):
        ops.gather_normalize(slide, torch.zeros((1, 2), dtype=torch.int64, device="cuda"), 64)


def test_cover_sampler_resume_from_saved_state(ops):
    """(accumulator, batch_index) is the whole sampler state: a run restored from a snapshot continues with exactly the coordinates
    of the uninterrupted run (the eligibility state is rebuilt by dh_cover_init)."""
    H, W, ps, B = 1500, 1300, 224, 32
    a = ops.CoverState(H, W, ps, 16, 2, B, seed=5)
    for _ in range(3):
        a.next_coords()
    snap, idx = a.accum.clone(), a.batch_index
    b = ops.CoverState(H, W, ps, 16, 2, B, seed=5)
    b.restore(snap, idx)
    for _ in range(6):
        ca, na = a.next_coords()
        cb, nb = b.next_coords()
        assert torch.equal(ca, cb) and int(na.item()) == int(nb.item())
    assert torch.equal(a.accum, b.accum)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_stitch_binned_random_shapes_vs_oracle(ops, variant):
    """Randomised shapes (slide size, patch size, downscale, class count, list length, overhanging and duplicated origins, row
    bands): every output of dh_stitch_binned is bit-identical to the reference loop -- with the default choice of tile kernel
    (variant 0), the row-run kernels only (1), the segment kernel wherever it applies (2) and the cell-lane sum kernel (3)."""
    from deephisto_b200 import _lib

    _lib.require_device().dh_stitch_binned_set_variant(variant)
    try:
        _stitch_binned_random_shapes(ops)
    finally:
        _lib.require_device().dh_stitch_binned_set_variant(0)


def _stitch_binned_random_shapes(ops):
    rng = np.random.default_rng(2024)
    for trial in range(24):
        ps = int(rng.choice([32, 50, 64, 100, 224]))
        d = int(rng.choice([1, 2, 3, 4, 7, 8, 16, 32]))
        if d > ps:
            d = ps
        n = int(rng.choice([1, 2, 3, 5, 8, 9, 16]))
        H, W = int(rng.integers(ps, 900)), int(rng.integers(ps, 900))
        if (H // d) * (W // d) * n > 6_000_000:
            H, W = max(ps, H // 3), max(ps, W // 3)
        P = int(rng.choice([0, 1, 7, 60, 300]))
        coords = np.stack([rng.integers(0, H, P), rng.integers(0, W, P)], 1).astype(np.int32)       # origins anywhere: footprints get clipped
        if P > 10:
            coords[3:6] = coords[2]                                                                # duplicated origins
        logits = (rng.standard_normal((P, n)) * 2).astype(np.float32)
        want, wcnt, wam = ostitch.stitch(logits, coords, H, W, ps, d)
        dh, dw = H // d, W // d
        if dh == 0 or dw == 0:
            continue
        lg, cd = torch.from_numpy(logits).cuda(), torch.from_numpy(coords).cuda()
        s, c, am = ops.stitch_binned(lg, cd, ps, d, dh, dw, want_count=True, want_argmax=True)
        tag = (trial, H, W, ps, d, n, P)
        assert np.array_equal(bits(s), want.view(np.int32)), tag
        assert np.array_equal(c.cpu().numpy().astype(np.int64), wcnt), tag
        assert np.array_equal(am.cpu().numpy().astype(np.int64), wam), tag
        r0 = int(rng.integers(0, dh))
        r1 = int(rng.integers(r0, dh)) + 1
        sb, cb, ab = ops.stitch_binned(lg, cd, ps, d, r1 - r0, dw, row_offset=r0, want_count=True, want_argmax=True)
        assert torch.equal(sb, s[r0:r1]) and torch.equal(cb, c[r0:r1]) and torch.equal(ab, am[r0:r1]), tag


def test_stitch_binned_and_upload_error_reporting(ops):
    from deephisto_b200 import _lib
    from deephisto_b200.slide import PinnedSlide, upload_rects

    lib = _lib.require_device()
    lg = torch.zeros((4, 11), device="cuda")
    cd = torch.zeros((4, 2), dtype=torch.int32, device="cuda")
    with pytest.raises(_lib.DeepHistoError, match="needs the sum map"):
        ops.stitch_binned(lg, cd, 64, 16, 8, 8, want_sum=False, want_argmax=True)            # 11 classes: class map through the sum map only
    out = torch.zeros((8, 8, 11), device="cuda")
    tiny = torch.zeros(256, dtype=torch.uint8, device="cuda")
    rc = lib.dh_stitch_binned(lg.data_ptr(), cd.data_ptr(), 4, 64, 16, 11, out.data_ptr(), None, None, 8, 8, 0, tiny.data_ptr(), 16,
                              torch.cuda.current_stream().cuda_stream)
    assert rc != 0 and "scratch too small" in _lib.last_error()
    with pytest.raises(ValueError):
        ops.stitch_binned(lg, cd[:3], 64, 16, 8, 8)
    # dh_upload_rects: rectangles must lie inside the slide; an empty list is a no-op
    host = PinnedSlide.from_numpy(synth.synth_slide(64, 48, 1))
    dev, n = upload_rects(host, [])
    assert n == 0 and int(dev.storage.sum()) == 0
    dev, n = upload_rects(host, [(0, 64, 0, 48)], tile=16)
    assert n == host.nbytes and np.array_equal(dev.to_numpy(), synth.synth_slide(64, 48, 1))
    bad = np.asarray([[0, 65, 0, 16]], dtype=np.int64)
    rc = lib.dh_upload_rects(dev.storage.data_ptr(), 64, host.pitch, host.host.data_ptr(), 1, bad.ctypes.data, torch.cuda.current_stream().cuda_stream)
    assert rc != 0 and "outside the slide" in _lib.last_error()
