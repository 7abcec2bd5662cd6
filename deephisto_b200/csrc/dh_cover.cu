// K6: coverage-driven random whole-slide sampler (coarse accumulator at 1/speedup resolution).
//
// Reference path replaced:
//   patch_samplers/full_samplers.py:105-114  _calc_probmap_sp   (uniform over cells with accum < dense_level,
//                                                                topped up with random cells when fewer than B)
//   patch_samplers/full_samplers.py:125-162  _prepare_indices   (np.random.choice(dh*dw, B, replace=False, p) = uniform
//                                                                random B-subset in random order; jitter; clamp)
//   patch_samplers/full_samplers.py:81-94    _update_accum_sp   (accum[y//16:(y+ps)//16, x//16:(x+ps)//16] += 1,
//                                                                filled_ratio = count_nonzero / size)
// The reference draws from the unseeded global numpy RNG, so parity is distributional; the stream
// defined here (Philox4x32-10, counters below) is restated on the CPU in oracle/cover.py and the
// coordinates are bit-identical to that restatement.
//
// The reference rescans the whole coarse grid three times per batch (probmap, choice, count_nonzero). Here the eligibility
// bitmask, the per-block eligible counts and the non-zero count are STATE kept in the scratch buffer and maintained
// incrementally by the accumulator update (a cell leaves the eligible set when its count reaches dense_level, joins the
// non-zero set when it leaves 0), so a batch costs O(B * footprint + number of blocks), independent of the slide size:
//   init (batch_index == 0, or dh_cover_init after restoring an accumulator): bitmask + block counts + non-zero count
//   group of batches: ONE persistent launch of one block (cover_group_kernel): the block counts live in shared memory as a three-level
//          count tree (block / 32 blocks / 1024 blocks) for the whole group, so a batch needs no scan of the grid at all: Philox draws,
//          duplicate check (the Fisher-Yates chain is the identity unless two draws collide), rank -> cell by four warp scans,
//          accumulator update by all 1024 threads, one counter store per batch
//   (cover_batch_kernel, one launch per batch with a full scan of the block counts, remains for coarse grids beyond 67 M cells)
#include "dh_common.cuh"

namespace dh {

constexpr int kCellsPerBlock = 2048;
constexpr int kMaxBatch = 2048;
static int g_cover_variant = 0;  // profiling / tests: 1 = one launch per batch (cover_batch_kernel) instead of the persistent group kernel

struct CoverScratch {
    uint32_t* block_cnt;  // [nb] eligible cells per block (state)
    uint32_t* block_off;  // [nb + 1] exclusive offsets of this batch, [nb] = M
    uint32_t* mask;       // [words] eligibility bits
    uint32_t* ranks;      // [B]
    uint32_t* extra;      // [B] top-up cells
    uint32_t* meta;       // [0] = M (eligible), [1] = number of top-up cells, [2] = non-zero accumulator cells (state)
};

__global__ void __launch_bounds__(256) cover_mask_kernel(const uint32_t* __restrict__ accum, int64_t cells, uint32_t dense_level,
                                                         CoverScratch s) {
    __shared__ uint32_t wsum[8];
    const int64_t base = (int64_t)blockIdx.x * kCellsPerBlock;
    uint32_t cnt = 0;
#pragma unroll
    for (int it = 0; it < kCellsPerBlock / 256; ++it) {
        int64_t c = base + it * 256 + threadIdx.x;
        bool el = (c < cells) && (accum[c] < dense_level);
        unsigned bal = __ballot_sync(0xffffffffu, el);
        if ((threadIdx.x & 31) == 0) s.mask[(base + it * 256 + threadIdx.x) >> 5] = bal;  // mask is padded to whole blocks
        cnt += el ? 1u : 0u;
    }
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        s.block_cnt[blockIdx.x] = t;
    }
}

__device__ __forceinline__ bool mask_test(const uint32_t* mask, uint32_t cell) { return (mask[cell >> 5] >> (cell & 31)) & 1u; }

// One batch in ONE launch of a single 1024-thread block (the serial part -- the hash-chained Fisher-Yates -- is the critical path
// anyway; what can run in parallel around it does: scan of the block counts, Philox draws, placement and accumulator update by 32
// warps). Phases, separated by block barriers:
//   scan   exclusive scan of the per-block eligible counts -> block_off, M
//   top-up (thread 0, rare) random non-eligible cells until M + extra >= B                      full_samplers.py:107-112
//   draw   all B Fisher-Yates draws j_i = i + randint(Mt - i) in parallel                        (Philox is counter based)
//   chain  (thread 0) the swaps, kept in a shared-memory hash map -> ranks                       full_samplers.py:135-143
//   place  one warp per pick: rank -> k-th eligible cell -> jitter -> clamp -> coords            full_samplers.py:144-153
//   update one warp per pick: accumulator += footprint, state transitions                        full_samplers.py:86-92
//   publish non-zero count
// stop_when_full: when every coarse cell is already covered the launch changes nothing and reports the count, so a host may enqueue
// several batches ahead of reading the count back (the reference stops at filled_ratio >= 1, full_samplers.py:263-274).
__global__ void __launch_bounds__(1024) cover_batch_kernel(uint32_t* __restrict__ accum, CoverScratch s, int nb, int64_t cells, int64_t dw,
                                                           int64_t H, int64_t W, int ps, int speedup, uint32_t dense_level, int B, uint32_t key0,
                                                           uint32_t key1, uint32_t batch_lo, uint32_t batch_hi, int32_t* __restrict__ coords,
                                                           uint32_t* __restrict__ nonzero_out, int stop_when_full, int hsize, int off_in_smem) {
    __shared__ uint32_t carry;
    __shared__ uint32_t wtot[32];
    extern __shared__ __align__(16) uint32_t cover_smem[];
    uint32_t* hkey = cover_smem;               // [hsize] open-addressing hash of the Fisher-Yates swaps, hsize = pow2 >= 2 B
    uint32_t* hval = hkey + hsize;             // [hsize]
    uint32_t* s_rank = hval + hsize;           // [B] first the draws j_i, then (in place) the ranks
    uint32_t* s_off = s_rank + B;              // [nb + 1] copy of block_off for the placement's binary search (when off_in_smem)
    const uint32_t* boff = off_in_smem ? s_off : s.block_off;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (stop_when_full && s.meta[2] >= (uint32_t)cells) {  // uniform: the state is only written by earlier launches
        if (threadIdx.x == 0) *nonzero_out = s.meta[2];
        return;
    }
    if (threadIdx.x == 0) carry = 0;
    for (int i = threadIdx.x; i < hsize; i += blockDim.x) hkey[i] = 0xffffffffu;
    __syncthreads();
    // ---- scan: exclusive scan of block counts, 1024 at a time
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        int b = b0 + threadIdx.x;
        uint32_t v = b < nb ? s.block_cnt[b] : 0u;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wtot[wid] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < wid; ++w) woff += wtot[w];
        uint32_t excl = carry + woff + inc - v;
        __syncthreads();
        if (b < nb) {
            s.block_off[b] = excl;
            if (off_in_smem) s_off[b] = excl;
        }
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    const uint32_t M = carry;
    // ---- top-up: add distinct random non-eligible cells until M + extra >= B
    if (threadIdx.x == 0) {
        s.block_off[nb] = M;
        if (off_in_smem) s_off[nb] = M;
        uint32_t n_extra = 0;
        if (M < (uint32_t)B) {
            uint32_t need = (uint32_t)B - M;
            uint32_t t = 0;
            while (n_extra < need) {
                Philox4 p = philox4x32_10(t, batch_lo, batch_hi, kStreamCoverTop, key0, key1);
                ++t;
                uint32_t cell = (uint32_t)(((uint64_t)p.v[0] * (uint64_t)cells) >> 32);
                if (mask_test(s.mask, cell)) continue;
                bool dup = false;
                for (uint32_t q = 0; q < n_extra; ++q) dup |= (s.extra[q] == cell);
                if (dup) continue;
                s.extra[n_extra++] = cell;
            }
        }
        s.meta[0] = M;
        s.meta[1] = n_extra;
        wtot[0] = n_extra;
    }
    __syncthreads();
    const uint32_t Mt = M + wtot[0];
    // ---- draw: j_i = i + randint(Mt - i) for every step of the partial Fisher-Yates over the virtual array a[i] = i
    for (uint32_t i = threadIdx.x; i < (uint32_t)B; i += blockDim.x) {
        Philox4 p = philox4x32_10(i, batch_lo, batch_hi, kStreamCoverPick, key0, key1);
        s_rank[i] = i + bounded_u32(p.v[0], Mt - i);
    }
    __syncthreads();
    // ---- chain: the swaps depend on each other (hash map of the touched entries)
    if (threadIdx.x == 0) {
        const uint32_t hmask = (uint32_t)hsize - 1;
        auto hget = [&](uint32_t k) -> uint32_t {
            uint32_t h = (k * 0x9E3779B1u) & hmask;
            while (hkey[h] != 0xffffffffu) {
                if (hkey[h] == k) return hval[h];
                h = (h + 1) & hmask;
            }
            return k;
        };
        auto hset = [&](uint32_t k, uint32_t v) {
            uint32_t h = (k * 0x9E3779B1u) & hmask;
            while (hkey[h] != 0xffffffffu && hkey[h] != k) h = (h + 1) & hmask;
            hkey[h] = k;
            hval[h] = v;
        };
        for (uint32_t i = 0; i < (uint32_t)B; ++i) {
            const uint32_t j = s_rank[i];
            const uint32_t aj = hget(j), ai = hget(i);
            s_rank[i] = aj;
            hset(j, ai);
        }
    }
    __syncthreads();
    // ---- place: one warp per pick
    for (int slot = wid; slot < B; slot += 32) {
        const uint32_t rank = s_rank[slot];
        uint32_t cell;
        if (rank >= M) {
            cell = s.extra[rank - M];
        } else {
            // block containing the rank: last b with block_off[b] <= rank
            int lo = 0, hi = nb - 1;
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if (boff[mid] <= rank) lo = mid; else hi = mid - 1;
            }
            uint32_t local = rank - boff[lo];
            const uint32_t w0 = (uint32_t)lo * (kCellsPerBlock / 32);
            // each lane owns 2 mask words of the block's 64; warp prefix over popcounts
            uint32_t m0 = s.mask[w0 + 2 * lane], m1 = s.mask[w0 + 2 * lane + 1];
            uint32_t pc = __popc(m0) + __popc(m1);
            uint32_t inc = pc;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            uint32_t excl = inc - pc;
            bool mine = (local >= excl) && (local < inc);
            unsigned bal = __ballot_sync(0xffffffffu, mine);
            int src = __ffs(bal) - 1;
            uint32_t found = 0;
            if (lane == src) {
                uint32_t r = local - excl;
                uint32_t word = m0, widx = w0 + 2 * lane;
                if (r >= (uint32_t)__popc(m0)) { r -= __popc(m0); word = m1; widx += 1; }
                for (uint32_t q = 0; q < r; ++q) word &= word - 1;  // r-th set bit of word
                found = widx * 32 + (__ffs(word) - 1);
            }
            cell = __shfl_sync(0xffffffffu, found, src);
        }
        // full_samplers.py:144-153 jitter + clamp
        Philox4 pj = philox4x32_10((uint32_t)slot, batch_lo, batch_hi, kStreamCoverJit, key0, key1);
        const int64_t pd2 = ps / speedup / 2;
        int64_t cy = cell / dw, cx = cell - cy * dw;
        int64_t y = (cy - pd2) * speedup + (int64_t)bounded_u32(pj.v[0], (uint32_t)speedup);
        int64_t x = (cx - pd2) * speedup + (int64_t)bounded_u32(pj.v[1], (uint32_t)speedup);
        y = y > H - ps ? H - ps : y; y = y < 0 ? 0 : y;
        x = x > W - ps ? W - ps : x; x = x < 0 ? 0 : x;
        if (lane == 0) { coords[2 * slot] = (int32_t)y; coords[2 * slot + 1] = (int32_t)x; }
    }
    __syncthreads();  // every pick has been located against the state of the PREVIOUS batch before the state changes
    // ---- update: accumulator footprint (full_samplers.py:86-92). State transitions: count reaches dense_level -> the cell leaves
    // the eligible set; count leaves 0 -> one more non-zero cell.
    for (int slot = wid; slot < B; slot += 32) {
        const int64_t y = coords[2 * slot], x = coords[2 * slot + 1];
        const int64_t r0 = y / speedup, r1 = (y + ps) / speedup, c0 = x / speedup, c1 = (x + ps) / speedup;
        const int fw = (int)(c1 - c0);
        const int tot = (int)(r1 - r0) * fw;
        uint32_t became_nonzero = 0;
        for (int f0 = lane; f0 < tot; f0 += 32 * 8) {  // 8 independent atomics in flight per lane (a 14 x 14 footprint is 7 per lane)
            int64_t cell[8];
            uint32_t old[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int f = f0 + 32 * u;
                const int rr = f / fw, cc = f - rr * fw;
                cell[u] = f < tot ? (r0 + rr) * dw + c0 + cc : -1;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) old[u] = cell[u] >= 0 ? atomicAdd(accum + cell[u], 1u) : 1u;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (cell[u] < 0) continue;
                became_nonzero += old[u] == 0u;
                if (old[u] + 1u == dense_level) {
                    atomicAnd(s.mask + (cell[u] >> 5), ~(1u << (cell[u] & 31)));
                    atomicSub(s.block_cnt + cell[u] / kCellsPerBlock, 1u);
                }
            }
        }
        for (int o = 16; o; o >>= 1) became_nonzero += __shfl_xor_sync(0xffffffffu, became_nonzero, o);
        if (lane == 0 && became_nonzero) atomicAdd(s.meta + 2, became_nonzero);
    }
    __syncthreads();
    if (threadIdx.x == 0) *nonzero_out = atomicAdd(s.meta + 2, 0u);
}

// ---- persistent group kernel -----------------------------------------------------------------------------------------------
// Shared-memory count tree over the eligibility bitmask: cnt[b] = eligible cells of block b (2048 cells = 64 mask words),
// sup[s] = sum of 32 blocks, top[t] = sum of 32 supers (<= 32 entries), M = total. Rank r -> cell = four warp scans
// (top, sup, block, mask words) + a bit select; the accumulator update keeps the tree current with shared-memory atomics.
struct CoverTree {
    uint32_t* cnt;
    uint32_t* sup;
    uint32_t* top;
    int ntop;
};

// lane `lane` holds v (count of child `lane`); finds the child containing rank r: returns its index, r becomes the rank inside it
__device__ __forceinline__ int tree_step(uint32_t v, uint32_t& r, int lane) {
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t excl = inc - v;
    const unsigned bal = __ballot_sync(0xffffffffu, r >= excl && r < inc);
    const int src = __ffs(bal) - 1;          // bal != 0: r < total by construction
    r -= __shfl_sync(0xffffffffu, excl, src);
    return src;
}

__device__ __forceinline__ uint32_t tree_find(const CoverTree& t, const uint32_t* __restrict__ mask, uint32_t r, int lane) {
    const int it = tree_step(lane < t.ntop ? t.top[lane] : 0u, r, lane);
    const int is = it * 32 + tree_step(t.sup[it * 32 + lane], r, lane);
    const int ib = is * 32 + tree_step(t.cnt[is * 32 + lane], r, lane);
    const uint32_t w0 = (uint32_t)ib * (kCellsPerBlock / 32);
    // each lane owns 2 of the block's 64 mask words (L2 reads: the words are updated by atomics of this very launch)
    uint2 mw;                                                 // (the mask is only 4-byte aligned inside the scratch buffer)
    mw.x = __ldcg(mask + w0 + 2 * lane);
    mw.y = __ldcg(mask + w0 + 2 * lane + 1);
    const uint32_t p0 = __popc(mw.x);
    const int src = tree_step(p0 + __popc(mw.y), r, lane);
    uint32_t found = 0;
    if (lane == src) {
        uint32_t word = mw.x, widx = w0 + 2 * lane, q = r;
        if (q >= p0) { q -= p0; word = mw.y; widx += 1; }
        for (uint32_t i = 0; i < q; ++i) word &= word - 1;      // q-th set bit of the word
        found = widx * 32 + (__ffs(word) - 1);
    }
    return __shfl_sync(0xffffffffu, found, src);
}

// n_batches consecutive batches, each with stop_when_full semantics. coords [n_batches][B][2], nonzero_out [n_batches].
__global__ void __launch_bounds__(1024) cover_group_kernel(uint32_t* __restrict__ accum, CoverScratch s, int nb, int64_t cells, int64_t dw,
                                                           int64_t H, int64_t W, int ps, int speedup, uint32_t dense_level, int B, uint32_t key0,
                                                           uint32_t key1, uint64_t first_batch, int n_batches, int32_t* __restrict__ coords,
                                                           uint32_t* __restrict__ nonzero_out, int stop_when_full, int hsize) {
    extern __shared__ __align__(16) uint32_t cover_smem[];
    __shared__ uint32_t s_M, s_nz, s_extra;
    const int nsup = (nb + 31) / 32, ntop = (nsup + 31) / 32;
    CoverTree tr;
    tr.cnt = cover_smem;                        // [ntop * 1024] (zero padded)
    tr.sup = tr.cnt + ntop * 1024;              // [ntop * 32]
    tr.top = tr.sup + ntop * 32;                // [32]
    tr.ntop = ntop;
    uint32_t* hkey = tr.top + 32;               // [hsize] Fisher-Yates swap hash (only when two draws of a batch collide)
    uint32_t* hval = hkey + hsize;              // [hsize]
    uint32_t* s_rank = hval + hsize;            // [B]
    int4* s_fp = reinterpret_cast<int4*>(s_rank + ((B + 3) & ~3));   // [B] footprint of every pick: first coarse cell, rows, columns
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // ---- build the tree from the block counts of the state
    for (int i = tid; i < ntop * 1024; i += 1024) tr.cnt[i] = i < nb ? s.block_cnt[i] : 0u;
    __syncthreads();
    for (int i = wid; i < ntop * 32; i += 32) {
        uint32_t v = tr.cnt[i * 32 + lane];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) tr.sup[i] = v;
    }
    __syncthreads();
    if (wid == 0) {
        uint32_t tot = 0;
        for (int i = 0; i < 32; ++i) {
            uint32_t v = i < ntop ? tr.sup[i * 32 + lane] : 0u;
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) tr.top[i] = v;
            tot += v;
        }
        if (lane == 0) { s_M = tot; s_nz = s.meta[2]; }
    }
    __syncthreads();

    const int fmax = ps / speedup + 1;                      // largest footprint side in coarse cells
    const int fcells = fmax * fmax;
    const int64_t pd2 = ps / speedup / 2;
    // accumulator update: thread -> (group of picks, footprint offset), fixed for the launch
    const int fpad = (fcells + 31) & ~31;
    const int fgroups = fpad <= 1024 ? 1024 / fpad : 0;
    const int fgrp = fgroups ? tid / fpad : 0, foff = fgroups ? tid - fgrp * fpad : 0;
    const bool fvalid = fgroups && fgrp < fgroups && foff < fcells;
    const int frow = foff / fmax, fcol = foff - frow * fmax;
    const int64_t frowoff = (int64_t)frow * dw;
    for (int g = 0; g < n_batches; ++g) {
        const uint64_t batch = first_batch + (uint64_t)g;
        const uint32_t batch_lo = (uint32_t)batch, batch_hi = (uint32_t)(batch >> 32);
        int32_t* const out = coords + (int64_t)g * B * 2;
        if (stop_when_full && s_nz >= (uint32_t)cells) {       // uniform (shared state, barriers below keep it stable)
            if (tid == 0) nonzero_out[g] = s_nz;
            continue;
        }
        const uint32_t M = s_M;
        // ---- top-up (full_samplers.py:107-112): distinct random non-eligible cells until M + extra >= B; rare (the last batches)
        if (M < (uint32_t)B) {
            if (tid == 0) {
                uint32_t n_extra = 0, t = 0;
                const uint32_t need = (uint32_t)B - M;
                while (n_extra < need) {
                    const Philox4 p = philox4x32_10(t, batch_lo, batch_hi, kStreamCoverTop, key0, key1);
                    ++t;
                    const uint32_t cell = (uint32_t)(((uint64_t)p.v[0] * (uint64_t)cells) >> 32);
                    if ((__ldcg(s.mask + (cell >> 5)) >> (cell & 31)) & 1u) continue;
                    bool dup = false;
                    for (uint32_t q = 0; q < n_extra; ++q) dup |= (s.extra[q] == cell);
                    if (dup) continue;
                    s.extra[n_extra++] = cell;
                }
                s_extra = n_extra;
            }
        } else if (tid == 0) {
            s_extra = 0;
        }
        __syncthreads();
        const uint32_t Mt = M + s_extra;
        // ---- draws j_i = i + randint(Mt - i) of the partial Fisher-Yates over the virtual array a[i] = i (full_samplers.py:135-143)
        for (uint32_t i = tid; i < (uint32_t)B; i += 1024) {
            const Philox4 p = philox4x32_10(i, batch_lo, batch_hi, kStreamCoverPick, key0, key1);
            s_rank[i] = i + bounded_u32(p.v[0], Mt - i);
        }
        __syncthreads();
        // pick i reads position j_i, which still holds j_i unless an EARLIER draw hit the same position: with all draws distinct the
        // chain of swaps is the identity on the picks (rank_i = j_i). Duplicates (probability ~ B^2 / 2 Mt) take the serial chain.
        int dup = 0;
        for (uint32_t i = tid; i < (uint32_t)B; i += 1024) {
            const uint32_t j = s_rank[i];
            for (uint32_t k = 0; k < i; ++k) dup |= (s_rank[k] == j);
        }
        if (__syncthreads_or(dup)) {
            for (int i = tid; i < hsize; i += 1024) hkey[i] = 0xffffffffu;
            __syncthreads();
            if (tid == 0) {
                const uint32_t hmask = (uint32_t)hsize - 1;
                auto hget = [&](uint32_t k) -> uint32_t {
                    uint32_t h = (k * 0x9E3779B1u) & hmask;
                    while (hkey[h] != 0xffffffffu) {
                        if (hkey[h] == k) return hval[h];
                        h = (h + 1) & hmask;
                    }
                    return k;
                };
                auto hset = [&](uint32_t k, uint32_t v) {
                    uint32_t h = (k * 0x9E3779B1u) & hmask;
                    while (hkey[h] != 0xffffffffu && hkey[h] != k) h = (h + 1) & hmask;
                    hkey[h] = k;
                    hval[h] = v;
                };
                for (uint32_t i = 0; i < (uint32_t)B; ++i) {
                    const uint32_t j = s_rank[i];
                    const uint32_t aj = hget(j), ai = hget(i);
                    s_rank[i] = aj;
                    hset(j, ai);
                }
            }
            __syncthreads();
        }
        // ---- place: one warp per pick: rank -> cell (count tree) -> jitter -> clamp (full_samplers.py:144-153)
        for (int slot = wid; slot < B; slot += 32) {
            const uint32_t rank = s_rank[slot];
            const uint32_t cell = rank >= M ? s.extra[rank - M] : tree_find(tr, s.mask, rank, lane);
            if (lane == 0) {
                const Philox4 pj = philox4x32_10((uint32_t)slot, batch_lo, batch_hi, kStreamCoverJit, key0, key1);
                const uint32_t ucy = cell / (uint32_t)dw;                          // cells < 2^31 (host check): 32-bit division
                const int64_t cy = ucy, cx = cell - ucy * (uint32_t)dw;
                int64_t y = (cy - pd2) * speedup + (int64_t)bounded_u32(pj.v[0], (uint32_t)speedup);
                int64_t x = (cx - pd2) * speedup + (int64_t)bounded_u32(pj.v[1], (uint32_t)speedup);
                y = y > H - ps ? H - ps : y; y = y < 0 ? 0 : y;
                x = x > W - ps ? W - ps : x; x = x < 0 ? 0 : x;
                const uint32_t uy = (uint32_t)y, ux = (uint32_t)x, us = (uint32_t)speedup;
                const uint32_t r0 = uy / us, r1 = (uy + (uint32_t)ps) / us, c0 = ux / us, c1 = (ux + (uint32_t)ps) / us;
                s_fp[slot] = make_int4((int)(r0 * (uint32_t)dw + c0), (int)(r1 - r0), (int)(c1 - c0), 0);
                reinterpret_cast<int2*>(out)[slot] = make_int2((int32_t)y, (int32_t)x);
            }
        }
        __syncthreads();  // every pick has been located against the state of the PREVIOUS batch before the state changes
        // ---- update (full_samplers.py:86-92): accum[y//s:(y+ps)//s, x//s:(x+ps)//s] += 1 over all B footprints. A thread owns ONE
        // footprint offset (frow, fcol) for the whole launch (computed once, above) and walks the picks in steps of `fgroups`; the
        // pick's footprint (first cell, height, width) was stored by the placement. No division in this loop.
        uint32_t became_nonzero = 0;
        auto touch = [&](int64_t cell, uint32_t old) {
            became_nonzero += old == 0u;
            if (old + 1u == dense_level) {          // the cell leaves the eligible set
                atomicAnd(s.mask + (cell >> 5), ~(1u << (cell & 31)));
                const int b = (int)(cell / kCellsPerBlock);
                atomicSub(tr.cnt + b, 1u);
                atomicSub(tr.sup + (b >> 5), 1u);
                atomicSub(tr.top + (b >> 10), 1u);
                atomicSub(&s_M, 1u);
            }
        };
        if (fgroups > 0) {
            for (int s0 = fgrp; s0 < B; s0 += fgroups * 8) {
                int64_t cell[8];
                uint32_t old[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int slot = s0 + u * fgroups;
                    cell[u] = -1;
                    if (fvalid && slot < B) {
                        const int4 fp = s_fp[slot];                                  // first cell, rows, columns of the footprint
                        if (frow < fp.y && fcol < fp.z) cell[u] = (int64_t)fp.x + frowoff + fcol;
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) old[u] = cell[u] >= 0 ? atomicAdd(accum + cell[u], 1u) : 1u;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (cell[u] >= 0) touch(cell[u], old[u]);
            }
        } else {                                       // footprints of more than 1024 coarse cells: flat index with divisions
            const int total = B * fcells;
            for (int w = tid; w < total; w += 1024) {
                const int slot = w / fcells, f = w - slot * fcells;
                const int rr = f / fmax, cc = f - rr * fmax;
                const int4 fp = s_fp[slot];
                if (rr < fp.y && cc < fp.z) {
                    const int64_t cell = (int64_t)fp.x + (int64_t)rr * dw + cc;
                    touch(cell, atomicAdd(accum + cell, 1u));
                }
            }
        }
        for (int o = 16; o; o >>= 1) became_nonzero += __shfl_xor_sync(0xffffffffu, became_nonzero, o);
        if (lane == 0 && became_nonzero) atomicAdd(&s_nz, became_nonzero);
        __syncthreads();
        if (tid == 0) nonzero_out[g] = s_nz;
    }
    // ---- the state goes back to global memory for the next launch
    __syncthreads();
    for (int i = tid; i < nb; i += 1024) s.block_cnt[i] = tr.cnt[i];
    if (tid == 0) { s.meta[2] = s_nz; s.meta[0] = s_M; }
}

__global__ void __launch_bounds__(256) cover_nonzero_kernel(const uint32_t* __restrict__ accum, int64_t cells, uint32_t* __restrict__ out) {
    uint32_t cnt = 0;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (int64_t)gridDim.x * blockDim.x) cnt += accum[c] != 0u;
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, cnt);
}

}  // namespace dh

using namespace dh;

static CoverScratch carve(uint32_t* scratch, int nb) {
    CoverScratch s;
    s.block_cnt = scratch;
    s.block_off = s.block_cnt + nb;
    s.mask = s.block_off + nb + 1;
    s.ranks = s.mask + (int64_t)nb * (kCellsPerBlock / 32);
    s.extra = s.ranks + kMaxBatch;
    s.meta = s.extra + kMaxBatch;
    return s;
}

extern "C" DH_API int64_t dh_cover_scratch_words(int64_t dh_, int64_t dw_) {
    int64_t cells = dh_ * dw_;
    int64_t nb = (cells + kCellsPerBlock - 1) / kCellsPerBlock;
    return nb + (nb + 1) + nb * (kCellsPerBlock / 32) + 2 * kMaxBatch + 8;
}

extern "C" DH_API int dh_cover_init(const uint32_t* accum, int64_t dh_, int64_t dw_, int dense_level, uint32_t* scratch, void* stream) {
    DH_REQUIRE(accum && scratch, "dh_cover_init: null pointer");
    DH_REQUIRE(dh_ > 0 && dw_ > 0 && dense_level > 0, "dh_cover_init: bad parameters");
    const int64_t cells = dh_ * dw_;
    DH_REQUIRE(cells < (1ll << 31), "dh_cover_init: coarse grid too large");
    const int nb = (int)((cells + kCellsPerBlock - 1) / kCellsPerBlock);
    CoverScratch s = carve(scratch, nb);
    cudaStream_t st = as_stream(stream);
    cover_mask_kernel<<<nb, 256, 0, st>>>(accum, cells, (uint32_t)dense_level, s);
    DH_CHECK_LAUNCH("cover_mask_kernel");
    cudaError_t e = cudaMemsetAsync(s.meta + 2, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    int64_t blocks = (cells + 255) / 256;
    int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
    cover_nonzero_kernel<<<grid, 256, 0, st>>>(accum, cells, s.meta + 2);
    DH_CHECK_LAUNCH("cover_nonzero_kernel");
    return DH_OK;
}

static int cover_launch(uint32_t* accum, int64_t dh_, int64_t dw_, int64_t H, int64_t W, int ps, int speedup, int dense_level, int B,
                        uint64_t seed, uint64_t first_batch, int n_batches, int32_t* coords_out, uint32_t* nonzero_out, uint32_t* scratch,
                        int stop_when_full, void* stream, const char* who) {
    DH_REQUIRE(accum && coords_out && nonzero_out && scratch, "%s: null pointer", who);
    DH_REQUIRE(ps > 0 && speedup > 0 && dense_level > 0, "%s: bad parameters", who);
    DH_REQUIRE(dh_ == H / speedup && dw_ == W / speedup && dh_ > 0 && dw_ > 0, "%s: coarse grid must be (H//speedup, W//speedup)", who);
    DH_REQUIRE(H >= ps && W >= ps, "%s: slide smaller than a patch", who);
    DH_REQUIRE(B >= 1 && B <= kMaxBatch, "%s: batch size %d outside 1..%d", who, B, kMaxBatch);
    const int64_t cells = dh_ * dw_;
    DH_REQUIRE(cells >= B, "%s: fewer coarse cells (%lld) than the batch size", who, (long long)cells);
    DH_REQUIRE(cells < (1ll << 31), "%s: coarse grid too large", who);
    if (n_batches == 0) return DH_OK;
    const int nb = (int)((cells + kCellsPerBlock - 1) / kCellsPerBlock);
    CoverScratch s = carve(scratch, nb);
    cudaStream_t st = as_stream(stream);
    if (first_batch == 0) {  // first batch of a run: build the state from the accumulator as it is
        int rc = dh_cover_init(accum, dh_, dw_, dense_level, scratch, stream);
        if (rc != DH_OK) return rc;
    }
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    int hsize = 64;
    while (hsize < 2 * B) hsize *= 2;
    const int nsup = (nb + 31) / 32, ntop = (nsup + 31) / 32;
    const size_t tree_smem = (size_t)(ntop * 1024 + ntop * 32 + 32 + 2 * hsize + ((B + 3) & ~3) + 4 * B) * sizeof(uint32_t);
    if (ntop <= 32 && tree_smem <= 200 * 1024 && g_cover_variant != 1) {
        // persistent kernel: the whole group in one launch (up to 32 768 blocks = 67 M coarse cells = a 131k x 131k slide at speedup 16)
        static size_t configured = 0;
        if (tree_smem > 48 * 1024 && tree_smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(cover_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tree_smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(cover_group_kernel)");
            configured = tree_smem;
        }
        cover_group_kernel<<<1, 1024, tree_smem, st>>>(accum, s, nb, cells, dw_, H, W, ps, speedup, (uint32_t)dense_level, B, k0, k1, first_batch,
                                                      n_batches, coords_out, nonzero_out, stop_when_full, hsize);
        DH_CHECK_LAUNCH("cover_group_kernel");
        return DH_OK;
    }
    // one launch per batch with a full scan of the block counts
    const int off_in_smem = nb + 1 <= 24 * 1024 ? 1 : 0;
    const size_t smem = (size_t)(2 * hsize + B + (off_in_smem ? nb + 1 : 0)) * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(cover_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(cover_batch_kernel)");
    }
    for (int g = 0; g < n_batches; ++g) {
        const uint64_t batch = first_batch + (uint64_t)g;
        cover_batch_kernel<<<1, 1024, smem, st>>>(accum, s, nb, cells, dw_, H, W, ps, speedup, (uint32_t)dense_level, B, k0, k1, (uint32_t)batch,
                                                  (uint32_t)(batch >> 32), coords_out + (int64_t)g * B * 2, nonzero_out + g, stop_when_full, hsize,
                                                  off_in_smem);
        DH_CHECK_LAUNCH("cover_batch_kernel");
    }
    return DH_OK;
}

extern "C" DH_API int dh_cover_set_variant(int variant) {
    if (variant < 0 || variant > 1) { set_error("dh_cover_set_variant: variant must be 0 (auto) or 1 (one launch per batch)"); return DH_ERR_INVALID; }
    g_cover_variant = variant;
    return DH_OK;
}

extern "C" DH_API int dh_cover_sample(uint32_t* accum, int64_t dh_, int64_t dw_, int64_t H, int64_t W, int ps, int speedup, int dense_level,
                               int B, uint64_t seed, uint64_t batch_index, int32_t* coords_out, uint32_t* nonzero_out,
                               uint32_t* scratch, int stop_when_full, void* stream) {
    return cover_launch(accum, dh_, dw_, H, W, ps, speedup, dense_level, B, seed, batch_index, 1, coords_out, nonzero_out, scratch, stop_when_full,
                        stream, "dh_cover_sample");
}

extern "C" DH_API int dh_cover_sample_group(uint32_t* accum, int64_t dh_, int64_t dw_, int64_t H, int64_t W, int ps, int speedup, int dense_level,
                                            int B, uint64_t seed, uint64_t first_batch_index, int n_batches, int32_t* coords_out,
                                            uint32_t* nonzero_out, uint32_t* scratch, void* stream) {
    DH_REQUIRE(n_batches >= 0, "dh_cover_sample_group: negative batch count");
    return cover_launch(accum, dh_, dw_, H, W, ps, speedup, dense_level, B, seed, first_batch_index, n_batches, coords_out, nonzero_out, scratch, 1,
                        stream, "dh_cover_sample_group");
}
