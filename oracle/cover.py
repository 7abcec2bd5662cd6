"""Coverage-driven random whole-slide sampler (CPU oracle).

Follows FullImageRndSampler (patch_samplers/full_samplers.py):
  _calc_probmap_sp :105-114  eligible = accum < dense_level, topped up with random cells when < B
  _prepare_indices :125-162  B distinct eligible cells, uniform; jitter randint(speedup); clamp
  _update_accum_sp :81-94    accum[y//s:(y+ps)//s, x//s:(x+ps)//s] += 1; filled = count_nonzero/size
The reference draws from the unseeded global numpy RNG (parity is distributional only); the draws
here restate deephisto_b200/csrc/dh_cover.cu: Philox4x32-10 keyed by seed with counters
(index, batch_lo, batch_hi, stream)."""

import numpy as np

from .philox import STREAM_COVER_JIT, STREAM_COVER_PICK, STREAM_COVER_TOP, bounded, philox4x32_10


class CoverSampler:
    def __init__(self, h: int, w: int, ps: int, batch_size: int, seed: int = 0, dense_level: int = 2, speedup: int = 16):
        self.h, self.w, self.ps, self.B, self.seed = h, w, ps, batch_size, seed
        self.dense_level, self.speedup = dense_level, speedup
        self.dh, self.dw = h // speedup, w // speedup
        self.accum = np.zeros([self.dh, self.dw], dtype=np.int64)
        self.batch_index = 0

    def next_coords(self):
        k0, k1 = self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF
        b_lo, b_hi = self.batch_index & 0xFFFFFFFF, (self.batch_index >> 32) & 0xFFFFFFFF
        cells = self.dh * self.dw
        flat = self.accum.reshape(-1)
        elig_mask = flat < self.dense_level
        elig = np.flatnonzero(elig_mask)                                     # index order
        M = len(elig)
        extra = []
        t = 0
        while M + len(extra) < self.B:                                        # :107-112 top-up
            r = philox4x32_10(t, b_lo, b_hi, STREAM_COVER_TOP, k0, k1)[0]
            t += 1
            cell = int(bounded(r, cells))
            if elig_mask[cell] or cell in extra:
                continue
            extra.append(cell)
        Mt = M + len(extra)
        # partial Fisher-Yates of the virtual array a[i] = i
        swaps = {}
        r = philox4x32_10(np.arange(self.B), b_lo, b_hi, STREAM_COVER_PICK, k0, k1)[0]
        ranks = []
        for i in range(self.B):
            j = i + int(bounded(r[i], Mt - i))
            aj, ai = swaps.get(j, j), swaps.get(i, i)
            ranks.append(aj)
            swaps[j] = ai
        cell_of = [int(elig[k]) if k < M else extra[k - M] for k in ranks]
        jit = philox4x32_10(np.arange(self.B), b_lo, b_hi, STREAM_COVER_JIT, k0, k1)
        jy, jx = bounded(jit[0], self.speedup), bounded(jit[1], self.speedup)
        pd2 = self.ps // self.speedup // 2                                    # :144
        coords = []
        for i, cell in enumerate(cell_of):
            y = (cell // self.dw - pd2) * self.speedup + int(jy[i])           # :146-151
            x = (cell % self.dw - pd2) * self.speedup + int(jx[i])
            y = max(min(y, self.h - self.ps), 0)                              # :128-132
            x = max(min(x, self.w - self.ps), 0)
            coords.append((y, x))
        s, p = self.speedup, self.ps
        for y, x in coords:                                                   # :86-92
            self.accum[y // s : (y + p) // s, x // s : (x + p) // s] += 1
        self.batch_index += 1
        filled = np.count_nonzero(self.accum) / self.accum.size              # :93
        return np.asarray(coords, dtype=np.int32), filled
