"""Row-band partition of whole-slide prediction (deephisto_b200/bands.py): host logic against the CPU oracle, and the
exchange step (all-gather of band maps) over gloo with world_size 2 -- no GPU."""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from deephisto_b200 import bands
from oracle import dense as odense
from oracle import stitch as ostitch

ROOT = Path(__file__).resolve().parent.parent

CASES = [  # (H, W, ps, stride, d, batch)
    (2048, 2048, 224, 112, 16, 64),
    (1000, 777, 224, 100, 3, 7),
    (1000, 777, 224, 100, 16, 7),
    (300, 260, 32, 48, 4, 5),      # stride > ps: gaps between patches
    (224, 500, 224, 64, 8, 3),     # h == ps: no main-grid rows
    (448, 448, 224, 224, 1, 4),
    (32, 32, 32, 8, 16, 1),        # one patch, dh = 2: with world 3 / 8 some ranks own no map rows
    (40, 64, 32, 8, 32, 3),        # dh = 1
    (500, 224, 224, 64, 8, 3),     # w == ps: the last column is the whole grid
]


@pytest.mark.parametrize("case", CASES)
def test_patch_origin_matches_oracle(case):
    H, W, ps, stride, d, B = case
    g = bands.dense_grid(H, W, ps, stride, B)
    coords, n = odense.dense_coords(H, W, ps, stride, B)
    assert (g.N, g.n_padded) == (n, len(coords))
    got = np.array([bands.patch_origin(g, i) for i in range(g.n_padded)], dtype=np.int32)
    assert np.array_equal(got, coords)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("case", CASES)
def test_band_plan_reproduces_full_map(case, world):
    """Stitching only the patches a band's plan lists gives exactly the full map's rows of that band (bit-exact sums:
    within a band the accumulation order is the reference's), and the band's slide rows contain every listed patch."""
    H, W, ps, stride, d, B = case
    coords, _ = odense.dense_coords(H, W, ps, stride, B)
    rng = np.random.default_rng(5)
    logits = rng.standard_normal((len(coords), 4)).astype(np.float32)
    full, fcnt, famax = ostitch.stitch(logits, coords, H, W, ps, d)
    dh = H // d
    covered = 0
    for rank in range(world):
        plan = bands.plan_band(H, W, ps, stride, d, B, rank, world)
        assert (plan.row_begin, plan.row_end) == bands.band_rows(dh, rank, world)
        idx = bands.patch_indices(plan)
        assert len(idx) == len(set(idx)) == plan.n_patches
        if plan.row_end == plan.row_begin:                       # world > dh: an empty band computes nothing
            assert plan.n_patches == 0 and plan.patch_ranges == []
        masked = np.zeros_like(logits)
        masked[idx] = logits[idx]
        # zeroed logits still ADD 0.0 in the oracle loop; a patch outside the plan must not touch the band at all
        keep = np.zeros(len(coords), bool)
        keep[idx] = True
        band, bcnt, bamax = ostitch.stitch(logits[keep], coords[keep], H, W, ps, d, plan.row_begin, plan.row_end)
        assert np.array_equal(band.view(np.uint32), full[plan.row_begin:plan.row_end].view(np.uint32))
        assert np.array_equal(bcnt, fcnt[plan.row_begin:plan.row_end])
        for i in idx:
            y, _ = coords[i]
            assert plan.slide_y0 <= y and y + ps <= plan.slide_y1
        assert plan.rows_max >= plan.row_end - plan.row_begin
        covered += plan.row_end - plan.row_begin
    assert covered == dh


def test_band_halo_overhead_is_small_at_full_size():
    """100k x 100k, stride 112, 8 ranks (BASELINE config 4): the recomputed halo patch rows stay under 3 %."""
    H = W = 100000
    g = bands.dense_grid(H, W, 224, 112, 64)
    total = sum(bands.plan_band(H, W, 224, 112, 16, 64, r, 8).n_patches for r in range(8))
    assert g.n_padded <= total < 1.03 * g.n_padded
    p = bands.plan_band(H, W, 224, 112, 16, 64, 3, 8)
    assert p.slide_y1 - p.slide_y0 < H // 8 + 3 * 224


WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from deephisto_b200 import bands
from deephisto_b200.examples.predict_full_patched import assemble_bands
from oracle import dense as odense, stitch as ostitch
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
for (H, W, ps, stride, d, B) in [(1000, 777, 224, 100, 3, 7), (2048, 2048, 224, 112, 16, 64), (40, 64, 32, 8, 32, 3)]:  # last: dh = 1 < world
    coords, _ = odense.dense_coords(H, W, ps, stride, B)
    logits = np.random.default_rng(9).standard_normal((len(coords), 5)).astype(np.float32)
    full, _, famax = ostitch.stitch(logits, coords, H, W, ps, d)
    plan = bands.plan_band(H, W, ps, stride, d, B, rank, world)
    if H // d < world and plan.row_end == plan.row_begin:
        assert plan.n_patches == 0 and plan.rows_max == 1
    keep = np.zeros(len(coords), bool); keep[bands.patch_indices(plan)] = True
    band, _, bamax = ostitch.stitch(logits[keep], coords[keep], H, W, ps, d, plan.row_begin, plan.row_end)
    pad = torch.zeros((plan.rows_max,) + band.shape[1:]); pad[: len(band)] = torch.from_numpy(band)
    apad = torch.zeros((plan.rows_max, band.shape[1]), dtype=torch.uint8); apad[: len(band)] = torch.from_numpy(bamax.astype(np.uint8))
    got = assemble_bands(pad, H // d, world, dist)
    agot = assemble_bands(apad, H // d, world, dist)
    assert got.shape == full.shape, (got.shape, full.shape)
    assert np.array_equal(got.numpy().view(np.uint32), full.view(np.uint32))
    assert np.array_equal(agot.numpy(), famax.astype(np.uint8))
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_band_assembly_gloo_world2(tmp_path):
    import subprocess

    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), str(ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in out, out[-2000:]


SHARD_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from deephisto_b200.slide import PinnedSlide, sharded_upload
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for H, W in ((100, 37), (7, 5), (2, 64)):                     # ragged shares: 100 rows over 3 ranks, fewer rows than ranks
    a = np.random.default_rng(H).integers(0, 256, (H, W, 3), dtype=np.uint8)
    host = PinnedSlide.from_numpy(a)
    full, copied = sharded_upload(host, device="cpu")
    per = -(-H // world)
    assert copied == max(0, min(H, (rank + 1) * per) - min(H, rank * per)) * host.pitch
    got = full.view(H, host.pitch)[:, : 3 * W].numpy().reshape(H, W, 3)
    assert np.array_equal(got, a)
    t = torch.tensor([copied]); dist.all_reduce(t)
    assert int(t.item()) == host.nbytes                       # every byte travelled exactly once
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


def test_sharded_slide_upload_gloo_world3(tmp_path):
    """slide.sharded_upload: each rank copies 1/world of the rows from its host copy, one all-gather replicates them (the NCCL path
    of the multi-GPU sampling e2e); here over gloo with CPU tensors, including shares that are ragged or empty."""
    import subprocess

    script = tmp_path / "shard_worker.py"
    script.write_text(SHARD_WORKER)
    port = _free_port()
    procs = []
    for r in range(3):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="3", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), str(ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in out, out[-2000:]


@pytest.mark.parametrize("case", [(1500, 1000, 224, 112, 16), (8192, 8192, 224, 224, 16), (1000, 777, 224, 100, 7), (448, 448, 224, 224, 4),
                                  (224, 900, 224, 112, 8), (900, 224, 224, 112, 8), (224, 224, 224, 112, 8)])
@pytest.mark.parametrize("world", [1, 3])
def test_stream_jobs_cover_every_patch_once_with_its_rows(case, world):
    """bands.stream_jobs (row chunks of a slide streamed through HBM): over all ranks' bands every entry of the padded enumeration a
    band needs appears in exactly one job of that band, and the job's slide rows contain the patch; chunks respect the byte budget."""
    H, W, ps, stride, B = case
    g = bands.dense_grid(H, W, ps, stride, B)
    row_bytes = (3 * W + 15) // 16 * 16
    for budget_rows in (ps, ps + 2 * stride, 10 * ps):
        for rank in range(world):
            plan = bands.plan_band(H, W, ps, stride, 8, B, rank, world)
            jobs = bands.stream_jobs(g, plan.patch_ranges, row_bytes, budget_rows * row_bytes)
            seen = []
            for y0, y1, ranges in jobs:
                assert 0 <= y0 < y1 <= H
                assert y1 - y0 <= max(budget_rows, ps) or len(jobs) == 1 or y0 == H - ps
                for first, count in ranges:
                    for i in range(first, first + count):
                        y, _ = bands.patch_origin(g, i)
                        assert y0 <= y and y + ps <= y1, (i, y, y0, y1)
                        seen.append(i)
            assert sorted(seen) == sorted(bands.patch_indices(plan))
            # a bounded first chunk (its upload cannot hide behind the CNN): same coverage, the first job is not larger than the others
            ramp = bands.stream_jobs(g, plan.patch_ranges, row_bytes, budget_rows * row_bytes, first_budget_bytes=ps * row_bytes)
            assert sorted(i for _, _, rs in ramp for f, c in rs for i in range(f, f + c)) == sorted(bands.patch_indices(plan))
            if ramp:
                assert ramp[0][1] - ramp[0][0] <= max(j[1] - j[0] for j in ramp)
    if g.ny > 0 and g.nx > 1:
        with pytest.raises(ValueError):
            bands.stream_jobs(g, [(1, g.nx)], row_bytes, 1 << 30)          # main-grid ranges must be whole grid rows


@pytest.mark.parametrize("case", [(3000, 224, 16, 16), (100000, 224, 16, 16), (1777, 224, 4, 16), (1000, 224, 3, 16), (500, 224, 16, 16),
                                  (224, 224, 16, 16), (4096, 64, 8, 16)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_rnd_band_rows(case, world):
    """bands.rnd_band (row bands of the coverage-driven random sampler): map rows partition [0, dh); every band's slide rows hold
    at least one patch, start on a coarse-cell boundary, contain the band's own pixel rows, and together cover the slide."""
    H, ps, d, sp = case
    dh = H // d
    rows, covered = 0, np.zeros(H, bool)
    for rank in range(world):
        p = bands.rnd_band(H, ps, d, sp, rank, world)
        assert (p.row_begin, p.row_end) == bands.band_rows(dh, rank, world) and p.rows_max >= p.row_end - p.row_begin
        rows += p.row_end - p.row_begin
        if p.row_end == p.row_begin:
            assert (p.slide_y0, p.slide_y1) == (0, 0)
            continue
        assert 0 <= p.slide_y0 < p.slide_y1 <= H and p.slide_y1 - p.slide_y0 >= ps and p.slide_y0 % sp == 0
        assert p.slide_y0 <= p.row_begin * d and p.row_end * d <= p.slide_y1
        assert p.patch_ranges == []
        covered[p.slide_y0 : p.slide_y1] = True
    assert rows == dh and covered.all()


LIST_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from deephisto_b200.examples.predict_full_patched import gather_patch_lists
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for lens in ([5, 0, 3], [0, 0, 0], [1, 7, 2]):
    n = 5
    def mk(r):
        g = torch.Generator().manual_seed(100 + r)
        return torch.randn((lens[r], n), generator=g), torch.randint(0, 1 << 20, (lens[r], 2), generator=g, dtype=torch.int32)
    lg, co = mk(rank)
    lg_all, co_all, counts = gather_patch_lists(lg, co, world, dist)
    assert counts == lens
    want_lg = torch.cat([mk(r)[0] for r in range(world)]); want_co = torch.cat([mk(r)[1] for r in range(world)])
    assert lg_all.dtype == torch.float32 and co_all.dtype == torch.int32
    assert torch.equal(lg_all.view(torch.int32), want_lg.view(torch.int32)) and torch.equal(co_all, want_co)
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


def test_gather_patch_lists_gloo_world3(tmp_path):
    """The exchange step of the random sampler's banded prediction: ragged per-rank (logits, coords) lists all-gathered in rank
    order, bit patterns preserved, empty lists included."""
    import subprocess

    script = tmp_path / "list_worker.py"
    script.write_text(LIST_WORKER)
    port = _free_port()
    procs = []
    for r in range(3):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="3", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), str(ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in out, out[-2000:]
