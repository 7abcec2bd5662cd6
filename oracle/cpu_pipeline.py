"""CPU restatements of the reference's end-to-end sampler pipelines, used ONLY as the timed CPU baseline
(bench.py `cpu_baseline` and `--impl reference`) and by tests. The reference itself cannot travel to the GPU
box (pure Python importing the absent psimage / shapely), so these follow its code path step by step:

AnnotatedRndCPU  AnnoRegionRndSampler.torch_generator (patch_samplers/region_samplers.py:685-738):
    a `spawn` ProcessPoolExecutor (:314-323, :721) over chunks of `batches_per_worker` batches (:722-728); each worker
    runs _gen_single_proc_torch (:593-622): image / class / region draws (:544-591), rejection sampling with the
    polygon ∩ square area test (:114-143, GEOS replaced by the oracle's C clip area), per-patch slice of the slide
    (:513-520, psimage replaced by a numpy memmap in /dev/shm), torch.tensor(data, float32) / 255 (:616); results are
    pickled back and stacked per batch (:729-735). The pool is created once and reused (the reference re-creates it
    per generator call), so worker start-up and imports are NOT part of the timed region.
dense_batches    FullImageDenseSampler.generator_torch (patch_samplers/full_samplers.py:437-452)."""

from __future__ import annotations

import multiprocessing as mp
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import torch

from . import dense, region

_G = {}


def _init_worker(slide_path, shape, images, layer, one_image):
    # One intra-op thread per worker process: with the default (all cores per worker, as the reference would run) the
    # OpenMP pools of the workers oversubscribe the host and throughput drops ~10x; this is the setting that favours the CPU path.
    torch.set_num_threads(1)
    _G["slide"] = np.asarray(np.memmap(slide_path, dtype=np.uint8, mode="r", shape=tuple(shape)))
    _G["rs"] = region.RegionSet(images, layer=layer, one_image_for_batch=one_image)


def _worker(args):
    n_slots, k, ps, ri, seed, slot_offset, spt = args
    rs, slide = _G["rs"], _G["slide"]
    coords, labels, images, status = region.sample(rs, n_slots, k, ps, ri, seed=seed, slot_offset=slot_offset, slots_per_table_draw=spt)
    feats = []
    for (y, x) in coords.tolist():
        data = slide[y : y + ps, x : x + ps, :]                                   # :513-520
        feats.append((torch.tensor(data, dtype=torch.float32) / 255).numpy())     # :616
    # Transport: plain numpy through the result pipe. (The reference returns torch tensors, which torch.multiprocessing moves
    # through /dev/shm file descriptors; containers with a small /dev/shm kill the workers, and it is not faster.)
    return np.stack(feats), labels.astype(np.int64), coords.astype(np.float32)


def _ping(_):
    return os.getpid()


class AnnotatedRndCPU:
    def __init__(self, slide_path: str, shape, images, layer: int = 1, one_image_for_batch: bool = True, max_workers: int | None = None):
        self.workers = max_workers or os.cpu_count()
        init = (slide_path, tuple(shape), images, layer, one_image_for_batch)
        if self.workers == 1:
            _init_worker(*init)
            self.pool = None
        else:
            self.pool = ProcessPoolExecutor(max_workers=self.workers, mp_context=mp.get_context("spawn"), initializer=_init_worker,
                                            initargs=init)
            list(self.pool.map(_ping, range(4 * self.workers)))                   # start every worker before anything is timed

    def batches(self, ps: int, batch_size: int, n_batches: int, batches_per_worker: int = 2, k: int = 4, ri: float = 0.75, seed: int = 0):
        """Yields (features [B,ps,ps,3] f32, labels [B] i64, coords [B,2] f32) CPU tensors like the reference."""
        q = [batches_per_worker] * (n_batches // batches_per_worker)
        if n_batches % batches_per_worker:
            q.append(n_batches % batches_per_worker)
        jobs, off = [], 0
        for nb in q:
            jobs.append((batch_size * nb, k, ps, ri, seed, off, batch_size * batches_per_worker))
            off += batch_size * nb
        results = map(_worker, jobs) if self.pool is None else self.pool.map(_worker, jobs)
        for feats, labels, coords in results:
            for i in range(0, len(feats), batch_size):                             # :729-735 stack per batch
                yield (torch.from_numpy(feats[i : i + batch_size]), torch.from_numpy(labels[i : i + batch_size]),
                       torch.from_numpy(coords[i : i + batch_size]))

    def close(self):
        if self.pool is not None:
            self.pool.shutdown(wait=True, cancel_futures=True)
            self.pool = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def dense_batches(slide: np.ndarray, ps: int, stride: int, batch_size: int):
    """FullImageDenseSampler.generator_torch on a numpy slide (full_samplers.py:437-452): np.stack -> astype -> /255 -> torch.tensor."""
    coords, _ = dense.dense_coords(slide.shape[0], slide.shape[1], ps, stride, batch_size)
    n_batches = len(coords) // batch_size
    for i in range(n_batches):
        c = coords[i * batch_size : (i + 1) * batch_size]
        patches = [slide[y : y + ps, x : x + ps, :] for y, x in c.tolist()]
        features = torch.tensor(np.stack(patches).astype(np.float32) / 255)
        yield features, torch.tensor(c.astype(np.float32)), i / n_batches
