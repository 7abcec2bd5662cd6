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
//   batch: ONE launch of one block (cover_batch_kernel below): scan, top-up, Fisher-Yates, placement, accumulator update, count
#include "dh_common.cuh"

namespace dh {

constexpr int kCellsPerBlock = 2048;
constexpr int kMaxBatch = 2048;

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
    extern __shared__ uint32_t cover_smem[];
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

extern "C" DH_API int dh_cover_sample(uint32_t* accum, int64_t dh_, int64_t dw_, int64_t H, int64_t W, int ps, int speedup, int dense_level,
                               int B, uint64_t seed, uint64_t batch_index, int32_t* coords_out, uint32_t* nonzero_out,
                               uint32_t* scratch, int stop_when_full, void* stream) {
    DH_REQUIRE(accum && coords_out && nonzero_out && scratch, "dh_cover_sample: null pointer");
    DH_REQUIRE(ps > 0 && speedup > 0 && dense_level > 0, "dh_cover_sample: bad parameters");
    DH_REQUIRE(dh_ == H / speedup && dw_ == W / speedup && dh_ > 0 && dw_ > 0, "dh_cover_sample: coarse grid must be (H//speedup, W//speedup)");
    DH_REQUIRE(H >= ps && W >= ps, "dh_cover_sample: slide smaller than a patch");
    DH_REQUIRE(B >= 1 && B <= kMaxBatch, "dh_cover_sample: batch size %d outside 1..%d", B, kMaxBatch);
    const int64_t cells = dh_ * dw_;
    DH_REQUIRE(cells >= B, "dh_cover_sample: fewer coarse cells (%lld) than the batch size", (long long)cells);
    DH_REQUIRE(cells < (1ll << 31), "dh_cover_sample: coarse grid too large");
    const int nb = (int)((cells + kCellsPerBlock - 1) / kCellsPerBlock);
    CoverScratch s = carve(scratch, nb);
    cudaStream_t st = as_stream(stream);
    if (batch_index == 0) {  // first batch of a run: build the state from the accumulator as it is
        int rc = dh_cover_init(accum, dh_, dw_, dense_level, scratch, stream);
        if (rc != DH_OK) return rc;
    }
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const uint32_t b_lo = (uint32_t)batch_index, b_hi = (uint32_t)(batch_index >> 32);
    int hsize = 64;
    while (hsize < 2 * B) hsize *= 2;
    const int off_in_smem = nb + 1 <= 24 * 1024 ? 1 : 0;  // up to 96 KB of block offsets (a 112k x 112k slide at speedup 16)
    const size_t smem = (size_t)(2 * hsize + B + (off_in_smem ? nb + 1 : 0)) * sizeof(uint32_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(cover_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(cover_batch_kernel)");
    }
    cover_batch_kernel<<<1, 1024, smem, st>>>(accum, s, nb, cells, dw_, H, W, ps, speedup, (uint32_t)dense_level, B, k0, k1, b_lo, b_hi, coords_out,
                                              nonzero_out, stop_when_full, hsize, off_in_smem);
    DH_CHECK_LAUNCH("cover_batch_kernel");
    return DH_OK;
}

extern "C" DH_API int dh_cover_sample_group(uint32_t* accum, int64_t dh_, int64_t dw_, int64_t H, int64_t W, int ps, int speedup, int dense_level,
                                            int B, uint64_t seed, uint64_t first_batch_index, int n_batches, int32_t* coords_out,
                                            uint32_t* nonzero_out, uint32_t* scratch, void* stream) {
    DH_REQUIRE(n_batches >= 0, "dh_cover_sample_group: negative batch count");
    for (int g = 0; g < n_batches; ++g) {
        int rc = dh_cover_sample(accum, dh_, dw_, H, W, ps, speedup, dense_level, B, seed, first_batch_index + (uint64_t)g,
                                 coords_out + (int64_t)g * B * 2, nonzero_out + g, scratch, 1, stream);
        if (rc != DH_OK) return rc;
    }
    return DH_OK;
}
