"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11),
vectorised in numpy. Restates deephisto_b200/csrc/dh_common.cuh::philox4x32_10 -- the reference itself
uses the unseeded global numpy RNG (full_samplers.py:110-111,137-150; region_samplers.py:123-124),
so this stream is the new build's own contract (DESIGN.md "Philox contract")."""

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_TABLE, STREAM_GROUP, STREAM_ATTEMPT, STREAM_COVER_TOP, STREAM_COVER_PICK, STREAM_COVER_JIT = 1, 2, 3, 4, 5, 6


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable unsigned ints; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)))
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def bounded(r, n):
    """uniform int in [0, n) from a uint32 word: (r * n) >> 32 (dh_common.cuh::bounded_u32)."""
    return ((np.asarray(r, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def u01(r):
    """(r + 0.5) / 2^32 as float64 (dh_region.cu::u01)."""
    return (np.asarray(r, dtype=np.float64) + 0.5) * (1.0 / 4294967296.0)
