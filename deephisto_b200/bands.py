"""Row-band partition of whole-slide patched prediction across the GPUs of one box (BASELINE config 4).

Pure integer host logic (no torch, no CUDA) so that it is testable on CPU; the enumeration it indexes is the
reference's FullImageDenseSampler._create_batched_coords (patch_samplers/full_samplers.py:374-404):
    [0, ny*nx)              main grid, row-major (y outer)
    [ny*nx, ny*nx+ny)       last column, one per grid row
    [ny*nx+ny, ny*nx+ny+nx) last row, one per grid column
    ny*nx+ny+nx             corner, followed by (n_padded - N) padding copies of the corner
and the stitched map is ImagePredictorPatched.process's (examples/predict_full_patched.py:40-63):
    prediction[y//d:(y+ps)//d, x//d:(x+ps)//d, :] += logits_i.

Rank r of G owns map rows [r*dh//G, (r+1)*dh//G). It computes the logits of exactly the patches that cover those
rows -- patch rows that straddle a band edge are recomputed by both neighbours (<= ceil(ps/stride) patch rows per
edge) -- so no logits cross the wire; the only exchange is the assembly of the finished band maps (NCCL all-gather).
Inside a band dh_stitch_dense sums in the reference's order, so the assembled sum map is bit-identical to the
single-GPU one."""

from __future__ import annotations

from dataclasses import dataclass, field


def range_len(stop: int, step: int) -> int:
    """len(range(0, stop, step))"""
    return 0 if stop <= 0 else (stop + step - 1) // step


@dataclass(frozen=True)
class DenseGrid:
    H: int
    W: int
    ps: int
    stride: int
    batch_size: int
    ny: int
    nx: int
    N: int
    n_padded: int

    @property
    def main_n(self) -> int:
        return self.ny * self.nx


def dense_grid(H: int, W: int, ps: int, stride: int, batch_size: int) -> DenseGrid:
    if H < ps or W < ps:
        raise ValueError(f"slide {H}x{W} smaller than patch {ps}")
    ny, nx = range_len(H - ps, stride), range_len(W - ps, stride)
    N = ny * nx + ny + nx + 1
    npad = (N + batch_size - 1) // batch_size * batch_size if batch_size > 0 else N
    return DenseGrid(H, W, ps, stride, batch_size, ny, nx, N, npad)


def cover_1d(i: int, cnt: int, ps: int, stride: int, d: int, last_cell: int) -> tuple[int, int, bool]:
    """Main-grid indices [lo, hi] (inclusive, empty if lo > hi) whose patch covers map cell i along one axis, and
    whether the last row/column patch covers it. A patch at origin y covers cell i  <=>  y//d <= i < (y+ps)//d."""
    e = (i + 1) * d
    hi = min((e - 1) // stride, cnt - 1)
    t = e - ps
    lo = 0 if t <= 0 else (t + stride - 1) // stride
    return lo, hi, i >= last_cell


@dataclass
class BandPlan:
    rank: int
    world: int
    row_begin: int                     # map rows owned: [row_begin, row_end)
    row_end: int
    rows_max: int                      # largest band height over all ranks (all-gather padding)
    patch_ranges: list = field(default_factory=list)   # [(first, count)] into the padded dense enumeration
    slide_y0: int = 0                  # slide rows this band reads: [slide_y0, slide_y1)
    slide_y1: int = 0

    @property
    def n_patches(self) -> int:
        return sum(c for _, c in self.patch_ranges)


def band_rows(dh: int, rank: int, world: int) -> tuple[int, int]:
    return rank * dh // world, (rank + 1) * dh // world


def plan_band(H: int, W: int, ps: int, stride: int, d: int, batch_size: int, rank: int, world: int) -> BandPlan:
    g = dense_grid(H, W, ps, stride, batch_size)
    dh = H // d
    r0, r1 = band_rows(dh, rank, world)
    rows_max = max(band_rows(dh, r, world)[1] - band_rows(dh, r, world)[0] for r in range(world))
    plan = BandPlan(rank, world, r0, r1, rows_max)
    if r1 <= r0:
        return plan
    last_cell = (H - ps) // d
    lo, _, _ = cover_1d(r0, g.ny, ps, stride, d, last_cell)
    _, hi, needs_last = cover_1d(r1 - 1, g.ny, ps, stride, d, last_cell)
    y0, y1 = None, None
    if hi >= lo:
        plan.patch_ranges.append((lo * g.nx, (hi - lo + 1) * g.nx))            # main-grid rows lo..hi
        plan.patch_ranges.append((g.main_n + lo, hi - lo + 1))                 # their last-column patches
        y0, y1 = lo * stride, hi * stride + ps
    if needs_last:
        first = g.main_n + g.ny
        plan.patch_ranges.append((first, g.n_padded - first))                  # last row, corner, padding copies
        y0 = H - ps if y0 is None else min(y0, H - ps)
        y1 = H
    if y0 is not None:
        plan.slide_y0, plan.slide_y1 = y0, y1
    return plan


def rnd_band(H: int, ps: int, d: int, speedup: int, rank: int, world: int) -> BandPlan:
    """Row band of rank `rank` for prediction with the coverage-driven RANDOM sampler (the reference's default sampler,
    examples/predict_full_patched.py:156-163): map rows as in plan_band; slide rows [slide_y0, slide_y1) = the band's pixel rows
    widened to whole coarse cells of the sampler's 1/speedup accumulator (full_samplers.py:81-94), and to at least one patch.
    The rank runs its own coverage sampler over exactly these rows, so every coarse cell of the band gets covered by the rank that
    owns it; patches never leave [slide_y0, slide_y1), and where neighbouring bands had to be widened into each other the stitch
    still sees every patch (the ranks exchange their (coords, logits) lists, 28 bytes per patch). patch_ranges stays empty: the
    patches are drawn, not enumerated."""
    if H < ps:
        raise ValueError(f"slide height {H} smaller than patch {ps}")
    dh = H // d
    r0, r1 = band_rows(dh, rank, world)
    rows_max = max(band_rows(dh, r, world)[1] - band_rows(dh, r, world)[0] for r in range(world))
    plan = BandPlan(rank, world, r0, r1, rows_max)
    if r1 <= r0:
        return plan
    y0 = (r0 * d) // speedup * speedup
    y1 = min(H, -(-(r1 * d) // speedup) * speedup)
    if r1 == dh:
        y1 = H                                       # the last band also owns the slide rows below the last whole map row
    if y1 - y0 < ps:                                 # thin band: widen (downwards first) to hold one patch
        y1 = min(H, y0 + ps)
        y0 = max(0, y1 - ps) // speedup * speedup
    plan.slide_y0, plan.slide_y1 = y0, y1
    return plan


def patch_indices(plan: BandPlan) -> list[int]:
    out: list[int] = []
    for first, count in plan.patch_ranges:
        out.extend(range(first, first + count))
    return out


def patch_origin(g: DenseGrid, i: int) -> tuple[int, int]:
    """(y, x) of entry i of the padded enumeration (same arithmetic as dh_dense_coords)."""
    if i < g.main_n:
        gy = i // g.nx
        return gy * g.stride, (i - gy * g.nx) * g.stride
    if i < g.main_n + g.ny:
        return (i - g.main_n) * g.stride, g.W - g.ps
    if i < g.main_n + g.ny + g.nx:
        return g.H - g.ps, (i - g.main_n - g.ny) * g.stride
    return g.H - g.ps, g.W - g.ps


def stream_jobs(g: DenseGrid, patch_ranges, row_bytes: int, budget_bytes: int,
                first_budget_bytes: int = None) -> list[tuple[int, int, list[tuple[int, int]]]]:
    """Row chunks for predicting a slide that is NOT resident in HBM: [(slide_y0, slide_y1, [(first, count)])] -- each job uploads
    slide rows [slide_y0, slide_y1) (at most ~budget_bytes of `row_bytes`-wide rows, never less than one patch row) and computes
    the listed entries of the padded dense enumeration. `patch_ranges` is what plan_band produces: whole main-grid rows, their
    last-column entries (these ride with their grid rows) and optionally the tail (last row, corner, padding copies). Consecutive
    chunks re-upload the ps - stride rows they share. `first_budget_bytes` (optional) bounds the FIRST chunk alone: nothing can hide
    its upload, so a small first chunk lets the CNN start early."""
    ps, stride = g.ps, g.stride

    def rows_for(budget):
        return max(1, (max(budget // max(row_bytes, 1), ps) - ps) // stride + 1)   # main-grid rows per chunk

    per = rows_for(budget_bytes)
    first_per = per if first_budget_bytes is None else min(per, rows_for(first_budget_bytes))
    jobs = []
    for first, count in patch_ranges:
        if count <= 0:
            continue
        if first < g.main_n:                                              # whole main-grid rows [lo, hi)
            if first % g.nx or count % g.nx:
                raise ValueError("main-grid patch ranges must cover whole grid rows")
            lo, hi = first // g.nx, (first + count) // g.nx
            a = lo
            while a < hi:
                b = min(a + (first_per if not jobs else per), hi)
                jobs.append((a * stride, (b - 1) * stride + ps, [(a * g.nx, (b - a) * g.nx), (g.main_n + a, b - a)]))
                a = b
        elif first >= g.main_n + g.ny:                                     # last row, corner, padding copies
            jobs.append((g.H - ps, g.H, [(first, count)]))
        elif g.nx == 0:                                                    # W == ps: the last column IS the grid, nothing to ride with
            lo, hi = first - g.main_n, first - g.main_n + count
            a = lo
            while a < hi:
                b = min(a + (first_per if not jobs else per), hi)
                jobs.append((a * stride, (b - 1) * stride + ps, [(g.main_n + a, b - a)]))
                a = b
        # otherwise entries [main_n + lo, main_n + hi) (last column) ride with their grid rows above
    covered = sum(c for _, _, rs in jobs for _, c in rs)
    wanted = sum(c for _, c in patch_ranges if c > 0)
    if covered != wanted:
        raise ValueError(f"stream_jobs covers {covered} of {wanted} patch entries (last-column ranges must match their grid rows)")
    return jobs
