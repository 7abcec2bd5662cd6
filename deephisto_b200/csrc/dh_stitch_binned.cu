// K5b: deterministic stitching for ARBITRARY patch coordinates (random samplers, any iterator of patches).
//
// Reference path replaced:
//   examples/predict_full_patched.py:47-54   for each batch, for each patch in sampler order:
//       prediction[y//d:(y+ps)//d, x//d:(x+ps)//d, :] += logits_i
//
// dh_stitch_scatter adds with atomics (order not preserved, 1e-5 parity, read-modify-write traffic). This file is the GATHER
// formulation for a coordinate LIST: the map is cut into warp-sized tiles, every patch is binned into the tiles its footprint
// touches, and one warp per tile adds the covering patches in ascending patch index -- the reference's order -- with plain
// fp32 adds starting from 0. The result is bit-identical to the numpy loop, has no atomics on the map, and writes every
// output exactly once (HBM write roofline: dh*dw*n*4 bytes for the sum map, dh*dw for the class map) with streaming stores
// (st.global.cs: a map line is never touched again; default-policy stores cost 8-12 % at d <= 4, profiles/r02_stitch.md).
//
//   count  : one thread per patch, atomicAdd on the per-tile counters                  (P * tiles-per-patch integer atomics)
//   alloc  : one thread per tile: list segment = warp-aggregated atomicAdd on one cursor (segment placement is arbitrary)
//   fill   : one thread per patch, patch index into the tile's list segment           (arbitrary order inside a segment)
//   tile   : one warp per tile: rank-sort the segment by patch index (shared memory), clip the footprints to tile-relative
//            ranges, then walk the tile's rows. Rows between two footprint boundaries have identical values (run-length): the
//            lane's values are recomputed only at a boundary and stored to every row of the run.
// Tile kernels (they share the binning):
//   bin_seg_kernel (n <= 8 classes, the product path): inside a row run the tile's columns fall into SEGMENTS of cells with the same
//       covering patches (boundaries = the column edges of the tile's patches, found once per tile). Per run one lane per segment adds
//       the covering patches' logits (n adds per patch and SEGMENT instead of per patch and output float), the sums go to shared memory,
//       every lane fetches the values of its output units from its precomputed segment slots and streams them to all rows of the run.
//       Store flavours: 16-byte stores (aligned rows), 16-byte stores at a per-row shift (rows not 16-byte aligned), scalar, and the
//       class-map / count outputs (1 or 4 cells per lane).
//   bin_tile_kernel / bin_tile_phased_kernel (any n; round-1 formulation: every lane re-sums its own floats per run): more than 8
//       classes, and the A/B reference for the segment kernel (dh_stitch_binned_set_variant(1)).
#include "dh_common.cuh"
#include <cstdlib>

namespace dh {

constexpr int kBinWarps = 8;    // warps per CTA, one tile each (no block-level synchronisation anywhere)
constexpr int kBinCap = 128;    // patches per tile staged in shared memory; longer lists take the global-memory path
constexpr int kBinMaxN = 8;     // classes the cell kernel keeps in registers

static int g_bin_tile_rows = 0;  // 0 = heuristic; profiling override (dh_stitch_binned_set_tile_rows)
static int g_bin_sparse = -1;    // cell-lane kernel, per-lane patch sets: -1 = footprints under 24 cells, 0 / 1 = profiling override (DH_BIN_SPARSE)
static int g_bin_variant = 0;    // 0 = auto (sum maps of n <= 8 classes: cell-lane kernel for footprints under 2048 floats on aligned rows and
                                 // under 24 cells on unaligned rows, segment kernel for wider footprints on unaligned rows; row-run kernels otherwise),
                                 // 1 = row-run kernels only, 2 = segment kernel whenever n <= 8, 3 = cell-lane kernel whenever n <= 8
                                 // (dh_stitch_binned_set_variant)

struct BinGeom {
    int64_t rows, row_offset, dw;
    int64_t units_per_row;  // sum mode: dw * n floats; cell mode: dw cells
    int64_t nty, ntx;
    int ps, d, n;
    int scale;              // units per cell: n (sum mode) or 1 (cell mode)
    int TH, TW;             // tile = TH rows x TW units (TW = 32 * VEC * G)
    int G;                  // groups of 32 * VEC units per tile: a lane owns VEC consecutive units in each group
    int sparse;             // bin_cell_sum_kernel: footprints much narrower than the 32-cell tile -> per-lane patch sets (see the kernel)
    int phased;             // sum map rows not 16-byte aligned (dw * n % 4 != 0): in row r a tile covers units [t*TW + s_r, (t+1)*TW + s_r),
                            // s_r = (4 - r * dw * n) mod 4, so that every 4-unit vector is aligned; lists are binned 3 units wider
};

struct BinRec {
    int r0, r1;      // band-relative rows [r0, r1)
    int u0, u1;      // units of the row [u0, u1)
};

// Programmatic dependent launch between the four kernels of one call (count -> alloc -> fill -> tile): a kernel signals at its start that
// its successor may be scheduled, the successor waits (before it touches anything its predecessors wrote) until they have completed
// and flushed. The launch latency and block ramp of three kernel boundaries then overlap with the predecessor's tail
// (~2-3 us each: 2 % of a d = 4 call, 7 % at d = 16). Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// numpy slice semantics of prediction[y//d:(y+ps)//d, x//d:(x+ps)//d]: stops clipped to the array, empty when start >= stop
__device__ __forceinline__ bool bin_footprint(const BinGeom& g, int y, int x, BinRec& f) {
    // 32-bit unsigned divisions for the usual non-negative origins ((unsigned)y + ps cannot wrap: both < 2^31)
    const unsigned ud = (unsigned)g.d, ups = (unsigned)g.ps;
    int64_t r0, r1, c0, c1;
    if (y >= 0) { r0 = (unsigned)y / ud; r1 = ((unsigned)y + ups) / ud; }
    else { r0 = 0; const int64_t e = (int64_t)y + g.ps; r1 = e > 0 ? e / g.d : 0; }  // callers never pass negative origins: keep the footprint inside the map (same as dh_stitch_scatter)
    if (x >= 0) { c0 = (unsigned)x / ud; c1 = ((unsigned)x + ups) / ud; }
    else { c0 = 0; const int64_t e = (int64_t)x + g.ps; c1 = e > 0 ? e / g.d : 0; }
    if (c1 > g.dw) c1 = g.dw;
    r0 = r0 > g.row_offset ? r0 : g.row_offset;
    r1 = r1 < g.row_offset + g.rows ? r1 : g.row_offset + g.rows;
    if (r1 <= r0 || c1 <= c0) return false;
    f.r0 = (int)(r0 - g.row_offset);
    f.r1 = (int)(r1 - g.row_offset);
    f.u0 = (int)(c0 * g.scale);
    f.u1 = (int)(c1 * g.scale);
    return true;
}

// cnt[ntiles] is followed by the list cursor cnt[ntiles] (zeroed together)
__global__ void __launch_bounds__(256) bin_alloc_kernel(uint32_t* __restrict__ cnt, int64_t ntiles, uint32_t* __restrict__ off, uint32_t* __restrict__ len) {
    pdl_trigger();
    pdl_wait();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint32_t c = t < ntiles ? cnt[t] : 0u;
    uint32_t inc = c;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    uint32_t base = 0;
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 31 && total) base = atomicAdd(cnt + ntiles, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (t < ntiles) { off[t] = base + inc - c; len[t] = c; }
}

template <bool FILL>
__global__ void __launch_bounds__(256) bin_patches_kernel(const int32_t* __restrict__ coords, int64_t P, BinGeom g, uint32_t* __restrict__ cnt,
                                                          const uint32_t* __restrict__ off, uint32_t* __restrict__ list) {
    pdl_trigger();
    if (FILL) pdl_wait();                                   // the segment offsets; the count pass follows a memset (plain stream order)
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        BinRec f;
        if (!bin_footprint(g, __ldg(coords + 2 * p), __ldg(coords + 2 * p + 1), f)) continue;
        const int ty0 = f.r0 / g.TH, ty1 = (f.r1 - 1) / g.TH, tx1 = (f.u1 - 1) / g.TW;
        const int tx0 = g.phased ? (f.u0 >= 3 ? (f.u0 - 3) / g.TW : 0) : f.u0 / g.TW;   // phased tiles reach 3 units to the right
        if constexpr (FILL) {
            // four tiles per trip: the atomics (whose results place the entries) are issued back to back instead of one round trip each
            const int nx = tx1 - tx0 + 1, nt = (ty1 - ty0 + 1) * nx;
            int ty = ty0, tx = tx0;
            for (int k0 = 0; k0 < nt; k0 += 4) {
                uint32_t slot[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    slot[u] = 0xffffffffu;
                    if (k0 + u < nt) {
                        const int64_t t = (int64_t)ty * g.ntx + tx;
                        slot[u] = off[t] + atomicSub(cnt + t, 1u) - 1u;                  // the counters run back to zero
                        if (++tx > tx1) { tx = tx0; ++ty; }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (k0 + u < nt) list[slot[u]] = (uint32_t)p;
            }
        } else {
            for (int ty = ty0; ty <= ty1; ++ty)
                for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(cnt + (int64_t)ty * g.ntx + tx, 1u);
        }
    }
}

// np.argmax: first maximum; NaN counts as the maximum (first NaN wins). Same rule as stitch_finalize_kernel.
__device__ __forceinline__ uint8_t first_argmax(const float* v, int n) {
    float best = v[0];
    int best_c = 0;
#pragma unroll
    for (int c = 1; c < kBinMaxN; ++c)
        if (c < n && (v[c] > best || (v[c] != v[c] && best == best))) { best = v[c]; best_c = c; }
    return (uint8_t)best_c;
}

// ---- tile kernels ------------------------------------------------------------------------------------------------------
// VEC units per lane. CELL = false: units are floats of a map row, output = sum map. CELL = true: units are cells (VEC = 1 or 4
// cells per lane), outputs = argmax byte and / or count. STAGED: the logits of the tile's patches sit in shared memory (n <= kBinMaxN).
//
// Rare path: more than kBinCap patches over one tile -> sorted list in global memory, footprints recomputed on the fly.
template <int VEC, bool CELL>
__device__ __noinline__ void bin_tile_slow(const float* __restrict__ logits, const int32_t* __restrict__ coords, const BinGeom& g, uint32_t beg, int L,
                                           const uint32_t* __restrict__ list, uint32_t* __restrict__ sorted, float* __restrict__ sum_map,
                                           uint32_t* __restrict__ count_map, uint8_t* __restrict__ argmax_map, int64_t ty, int64_t tx) {
    const int lane = threadIdx.x & 31;
    const int n = g.n;
    for (int j = lane; j < L; j += 32) {
        const uint32_t v = list[beg + j];
        int rank = 0;
        for (int i = 0; i < L; ++i) rank += list[beg + i] < v;
        sorted[beg + rank] = v;
    }
    __syncwarp();
  constexpr int W1 = CELL ? 1 : VEC;                       // units a lane handles per pass (the fast kernel's lane-to-unit map does not matter here)
  // phased tiles also own the 3 units right of the tile (a neighbour's vectors start 0..3 units late); writing them from both sides
  // stores identical values
  const int64_t tile_end = (tx + 1) * g.TW + (g.phased ? 3 : 0);
  const int64_t ulimit = tile_end < g.units_per_row ? tile_end : g.units_per_row;
  const int npass = (g.TW + (g.phased ? 3 : 0) + 32 * W1 - 1) / (32 * W1);   // TW need not be a multiple of 32 * W1 (cell-lane tiles)
  for (int gi = 0; gi < npass; ++gi) {
    const int64_t ubase = tx * g.TW + (int64_t)gi * 32 * W1 + (int64_t)lane * W1;
    const bool active = ubase < ulimit;
    int cls[W1];
    int c = (int)(ubase % n);
#pragma unroll
    for (int k = 0; k < W1; ++k) { cls[k] = c; c = c + 1 == n ? 0 : c + 1; }
    const int R0 = (int)(ty * g.TH);
    const int R1 = (int)((int64_t)R0 + g.TH < g.rows ? R0 + g.TH : g.rows);
    const bool want_vals = CELL ? argmax_map != nullptr : true;
    int r = R0;
    while (r < R1) {
        float acc[CELL ? kBinMaxN : W1];
#pragma unroll
        for (int k = 0; k < (CELL ? kBinMaxN : W1); ++k) acc[k] = 0.f;
        uint32_t hits = 0;
        int next = R1;
        for (int j = 0; j < L; ++j) {
            const uint32_t id = sorted[beg + j];
            BinRec f;
            bin_footprint(g, __ldg(coords + 2 * (int64_t)id), __ldg(coords + 2 * (int64_t)id + 1), f);
            if (f.r0 > r) { next = f.r0 < next ? f.r0 : next; continue; }  // starts below this row (warp-uniform)
            if (f.r1 <= r) continue;                                        // ended above
            next = f.r1 < next ? f.r1 : next;
            const int64_t lo = f.u0 - ubase, hi = f.u1 - ubase;             // covered units relative to the lane's first unit
            if (hi <= 0 || lo >= W1) continue;
            const float* lg = logits + (int64_t)id * n;
            if constexpr (CELL) {
                ++hits;
                if (want_vals) {
#pragma unroll
                    for (int q = 0; q < kBinMaxN; ++q)
                        if (q < n) acc[q] += __ldg(lg + q);
                }
            } else {
#pragma unroll
                for (int k = 0; k < W1; ++k)
                    if (k >= lo && k < hi) acc[k] += __ldg(lg + cls[k]);
            }
        }
        if (active) {
            if constexpr (CELL) {
                const uint8_t am = want_vals ? first_argmax(acc, n) : (uint8_t)0;
                int64_t o = (int64_t)r * g.units_per_row + ubase;
                for (int rr = r; rr < next; ++rr, o += g.units_per_row) {
                    if (argmax_map) argmax_map[o] = am;
                    if (count_map) count_map[o] = hits;
                }
            } else {
                float* o = sum_map + (int64_t)r * g.units_per_row + ubase;
                for (int rr = r; rr < next; ++rr, o += g.units_per_row) {
#pragma unroll
                    for (int k = 0; k < W1; ++k)
                        if (ubase + k < ulimit) o[k] = acc[k];
                }
            }
        }
        r = next;
    }
  }
}

__host__ __device__ inline int bin_warp_smem_bytes(int n, bool staged) { return kBinCap * 20 + kBinCap * 4 * (staged ? n : 1); }

struct BinStage {
    uint32_t* ids;   // [L] patch indices, ascending
    int2* rows;      // [L] band-relative rows [r0, r1)
    int2* cols;      // [L] units [u0, u1)
    float* lg;       // [L * n] logits (STAGED)
};

// per-warp staging: sorted patch ids, their row and unit ranges, their logits (the unsorted ids alias the logits area)
template <bool STAGED>
__device__ __forceinline__ BinStage bin_stage(unsigned char* base, const float* __restrict__ logits, const int32_t* __restrict__ coords,
                                              const BinGeom& g, const uint32_t* __restrict__ list, uint32_t beg, int L, int lane, bool load_logits) {
    BinStage st;
    const int n = g.n;
    st.ids = reinterpret_cast<uint32_t*>(base);
    st.rows = reinterpret_cast<int2*>(base + kBinCap * 4);
    st.cols = reinterpret_cast<int2*>(base + kBinCap * 12);
    st.lg = reinterpret_cast<float*>(base + kBinCap * 20);
    uint32_t* s_raw = reinterpret_cast<uint32_t*>(st.lg);
    for (int j = lane; j < L; j += 32) s_raw[j] = list[beg + j];
    __syncwarp();
    for (int j = lane; j < L; j += 32) {  // rank sort: patch indices are distinct within a tile
        const uint32_t v = s_raw[j];
        int rank = 0;
        for (int i = 0; i < L; ++i) rank += s_raw[i] < v;
        st.ids[rank] = v;
    }
    __syncwarp();
    for (int j = lane; j < L; j += 32) {
        const uint32_t id = st.ids[j];
        BinRec f;
        bin_footprint(g, __ldg(coords + 2 * (int64_t)id), __ldg(coords + 2 * (int64_t)id + 1), f);
        st.rows[j] = make_int2(f.r0, f.r1);
        st.cols[j] = make_int2(f.u0, f.u1);
    }
    if (STAGED && load_logits) {
        for (int e = lane; e < L * n; e += 32) {
            const int j = e / n;
            st.lg[e] = __ldg(logits + (int64_t)st.ids[j] * n + (e - j * n));
        }
    }
    __syncwarp();
    return st;
}


template <int VEC, int G, bool CELL, bool STAGED>
__global__ void __launch_bounds__(kBinWarps * 32) bin_tile_kernel(const float* __restrict__ logits, const int32_t* __restrict__ coords, BinGeom g,
                                                                  const uint32_t* __restrict__ off, const uint32_t* __restrict__ len,
                                                                  const uint32_t* __restrict__ list, uint32_t* __restrict__ sorted, float* __restrict__ sum_map,
                                                                  uint32_t* __restrict__ count_map, uint8_t* __restrict__ argmax_map) {
    extern __shared__ __align__(16) unsigned char bin_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t ctas_x = (g.ntx + kBinWarps - 1) / kBinWarps;
    const int64_t ty = blockIdx.x / ctas_x;
    const int64_t tx = (blockIdx.x % ctas_x) * kBinWarps + w;
    if (tx >= g.ntx) return;
    const int64_t t = ty * g.ntx + tx;
    pdl_wait();                                             // the fill pass (and through it count / alloc) has completed
    const uint32_t beg = off[t];
    const int L = (int)len[t];
    const int n = g.n;
    if (L > kBinCap) {
        bin_tile_slow<VEC, CELL>(logits, coords, g, beg, L, list, sorted, sum_map, count_map, argmax_map, ty, tx);
        return;
    }
    const BinStage stg = bin_stage<STAGED>(bin_smem + (size_t)w * bin_warp_smem_bytes(n, STAGED), logits, coords, g, list, beg, L, lane,
                                           !CELL || argmax_map != nullptr);
    uint32_t* s_ids = stg.ids;
    int2* s_rows = stg.rows;
    int2* s_cols = stg.cols;
    float* s_lg = stg.lg;

    constexpr int GS = 32 * VEC;                             // units per group: group gi of the tile starts GS * gi units further right
    const int ub = (int)(tx * g.TW) + lane * VEC;            // first unit owned by this lane in group 0 (units_per_row + TW < 2^31, host check)
    int cls[G][VEC];                                         // units_per_row % VEC == 0 (host check): a lane's VEC units are all inside or all outside
    if constexpr (!CELL) {
#pragma unroll
        for (int gi = 0; gi < G; ++gi) {
            int c = (ub + gi * GS) % n;
#pragma unroll
            for (int k = 0; k < VEC; ++k) { cls[gi][k] = c; c = c + 1 == n ? 0 : c + 1; }
        }
    }
    const int R0 = (int)(ty * g.TH);
    const int R1 = (int)((int64_t)R0 + g.TH < g.rows ? R0 + g.TH : g.rows);
    const bool want_vals = CELL ? argmax_map != nullptr : true;

    int r = R0;
    while (r < R1) {
        float acc[CELL ? VEC : G][CELL ? kBinMaxN : VEC];   // sum mode: [group][float]; cell mode: [cell][class]
#pragma unroll
        for (int gi = 0; gi < (CELL ? VEC : G); ++gi)
#pragma unroll
            for (int k = 0; k < (CELL ? kBinMaxN : VEC); ++k) acc[gi][k] = 0.f;
        uint32_t hits[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) hits[k] = 0;
        int next = R1;
        for (int c0 = 0; c0 < L; c0 += 32) {
            // lane j looks at patch c0 + j: does it cover row r, and where is its next row boundary? (ballot + warp minimum)
            const int j = c0 + lane;
            const int2 rw = j < L ? s_rows[j] : make_int2(0x7fffffff, 0x7fffffff);
            const int cand = rw.x > r ? rw.x : (rw.y > r ? rw.y : 0x7fffffff);
            const int nb = __reduce_min_sync(0xffffffffu, cand);
            next = nb < next ? nb : next;
            unsigned m = __ballot_sync(0xffffffffu, rw.x <= r && r < rw.y);
            while (m) {  // covering patches in ascending list index = the reference's order; warp-uniform trip count
                const int jj = c0 + __ffs(m) - 1;
                m &= m - 1;
                const int2 cu = s_cols[jj];
                const int lo = cu.x - ub, hi = cu.y - ub;     // covered units relative to the lane's first unit
                // adding +0.0f is the identity here: a sum that starts at +0.0f can never be -0.0f
                if constexpr (CELL) {
                    bool cov[VEC];
#pragma unroll
                    for (int k = 0; k < VEC; ++k) { cov[k] = k >= lo && k < hi; hits[k] += cov[k]; }
                    if (want_vals) {
#pragma unroll
                        for (int q = 0; q < kBinMaxN; ++q)
                            if (q < n) {
                                const float v = s_lg[jj * n + q];     // one broadcast read per class, used by all cells of the lane
#pragma unroll
                                for (int k = 0; k < VEC; ++k) acc[k][q] += cov[k] ? v : 0.f;
                            }
                    }
                } else {
                    const float* lg = STAGED ? s_lg + jj * n : logits + (int64_t)s_ids[jj] * n;
#pragma unroll
                    for (int gi = 0; gi < G; ++gi)
#pragma unroll
                        for (int k = 0; k < VEC; ++k) {
                            const float v = STAGED ? lg[cls[gi][k]] : __ldg(lg + cls[gi][k]);
                            acc[gi][k] += (k + gi * GS >= lo && k + gi * GS < hi) ? v : 0.f;
                        }
                }
            }
        }
        if constexpr (CELL) {
            if (ub < g.units_per_row) {                         // VEC == 4: dw % 4 == 0 and aligned maps (host check) -> one vector store per row
                uint8_t am[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) am[k] = want_vals ? first_argmax(acc[k], n) : (uint8_t)0;
                int64_t o = (int64_t)r * g.units_per_row + ub;
                for (int rr = r; rr < next; ++rr, o += g.units_per_row) {
                    if constexpr (VEC == 4) {
                        if (argmax_map) __stcs(reinterpret_cast<uchar4*>(argmax_map + o), make_uchar4(am[0], am[1], am[2], am[3]));
                        if (count_map) __stcs(reinterpret_cast<uint4*>(count_map + o), make_uint4(hits[0], hits[1], hits[2], hits[3]));
                    } else {
                        if (argmax_map) __stcs(argmax_map + o, am[0]);
                        if (count_map) __stcs(count_map + o, hits[0]);
                    }
                }
            }
        } else {
            float* o = sum_map + (int64_t)r * g.units_per_row + ub;
            for (int rr = r; rr < next; ++rr, o += g.units_per_row) {
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    if (ub + gi * GS >= g.units_per_row) continue;
                    if constexpr (VEC == 4) __stcs(reinterpret_cast<float4*>(o + gi * GS), make_float4(acc[gi][0], acc[gi][1], acc[gi][2], acc[gi][3]));
                    else {
#pragma unroll
                        for (int k = 0; k < VEC; ++k) __stcs(o + gi * GS + k, acc[gi][k]);
                    }
                }
            }
        }
        r = next;
    }
}

// Sum map whose rows are NOT 16-byte aligned (dw * n % 4 != 0; BinGeom::phased): in row r every lane's vector starts s_r units
// later, s_r = (4 - r * dw * n) mod 4, which makes its address a multiple of 16 bytes. A lane therefore sums the 7 units
// [ub, ub + 7) of a row run once and stores the window [ub + s_r, ub + s_r + 4) of each row; the first s_r units of a row are
// scalar stores of tile 0, the last vector of a row is clipped to the row end.
template <bool STAGED>
__global__ void __launch_bounds__(kBinWarps * 32) bin_tile_phased_kernel(const float* __restrict__ logits, const int32_t* __restrict__ coords, BinGeom g,
                                                                         const uint32_t* __restrict__ off, const uint32_t* __restrict__ len,
                                                                         const uint32_t* __restrict__ list, uint32_t* __restrict__ sorted,
                                                                         float* __restrict__ sum_map) {
    extern __shared__ __align__(16) unsigned char bin_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t ctas_x = (g.ntx + kBinWarps - 1) / kBinWarps;
    const int64_t ty = blockIdx.x / ctas_x;
    const int64_t tx = (blockIdx.x % ctas_x) * kBinWarps + w;
    if (tx >= g.ntx) return;
    const int64_t t = ty * g.ntx + tx;
    pdl_wait();                                             // the fill pass (and through it count / alloc) has completed
    const uint32_t beg = off[t];
    const int L = (int)len[t];
    const int n = g.n;
    if (L > kBinCap) {
        bin_tile_slow<4, false>(logits, coords, g, beg, L, list, sorted, sum_map, nullptr, nullptr, ty, tx);
        return;
    }
    const BinStage stg = bin_stage<STAGED>(bin_smem + (size_t)w * bin_warp_smem_bytes(n, STAGED), logits, coords, g, list, beg, L, lane, true);
    constexpr int NU = 7;                                    // units summed per lane: a 4-unit window at shift 0..3
    const int RF = (int)g.units_per_row;
    const int ub = (int)(tx * g.TW) + lane * 4;
    int cls[NU];
    {
        int c = ub % n;
#pragma unroll
        for (int k = 0; k < NU; ++k) { cls[k] = c; c = c + 1 == n ? 0 : c + 1; }
    }
    const int R0 = (int)(ty * g.TH);
    const int R1 = (int)((int64_t)R0 + g.TH < g.rows ? R0 + g.TH : g.rows);
    const int pstep = RF & 3;
    int r = R0;
    while (r < R1) {
        float acc[NU];
#pragma unroll
        for (int k = 0; k < NU; ++k) acc[k] = 0.f;
        int next = R1;
        for (int c0 = 0; c0 < L; c0 += 32) {
            const int j = c0 + lane;
            const int2 rw = j < L ? stg.rows[j] : make_int2(0x7fffffff, 0x7fffffff);
            const int cand = rw.x > r ? rw.x : (rw.y > r ? rw.y : 0x7fffffff);
            const int nb = __reduce_min_sync(0xffffffffu, cand);
            next = nb < next ? nb : next;
            unsigned m = __ballot_sync(0xffffffffu, rw.x <= r && r < rw.y);
            while (m) {  // covering patches in ascending list index = the reference's order
                const int jj = c0 + __ffs(m) - 1;
                m &= m - 1;
                const int2 cu = stg.cols[jj];
                const int lo = cu.x - ub, hi = cu.y - ub;
                const float* lg = STAGED ? stg.lg + jj * n : logits + (int64_t)stg.ids[jj] * n;
#pragma unroll
                for (int k = 0; k < NU; ++k) {
                    const float v = STAGED ? lg[cls[k]] : __ldg(lg + cls[k]);
                    acc[k] += (k >= lo && k < hi) ? v : 0.f;   // adding +0.0f is the identity: a sum that starts at +0.0f is never -0.0f
                }
            }
        }
        float* rowp = sum_map + (int64_t)r * RF;
        int p = (int)(((int64_t)r * RF) & 3);                 // sum_map is 16-byte aligned (host check)
        for (int rr = r; rr < next; ++rr) {
            const int s = (4 - p) & 3;
            float w0, w1, w2, w3;
            switch (s) {
                case 0: w0 = acc[0]; w1 = acc[1]; w2 = acc[2]; w3 = acc[3]; break;
                case 1: w0 = acc[1]; w1 = acc[2]; w2 = acc[3]; w3 = acc[4]; break;
                case 2: w0 = acc[2]; w1 = acc[3]; w2 = acc[4]; w3 = acc[5]; break;
                default: w0 = acc[3]; w1 = acc[4]; w2 = acc[5]; w3 = acc[6]; break;
            }
            const int u = ub + s;
            if (u + 4 <= RF) {
                __stcs(reinterpret_cast<float4*>(rowp + u), make_float4(w0, w1, w2, w3));
            } else if (u < RF) {                                 // the last vector of the row, clipped
                rowp[u] = w0;
                if (u + 1 < RF) rowp[u + 1] = w1;
                if (u + 2 < RF) rowp[u + 2] = w2;
            }
            if (tx == 0 && lane == 0) {                          // the s units before the row's first aligned vector
                if (s > 0) rowp[0] = acc[0];
                if (s > 1) rowp[1] = acc[1];
                if (s > 2) rowp[2] = acc[2];
            }
            rowp += RF;
            p = (p + pstep) & 3;
        }
        r = next;
    }
}

// ---- segment formulation ---------------------------------------------------------------------------------------------------
// KIND: 0 = sum map, 16-byte stores, G groups of 128 floats per tile; 1 = sum map, rows not 16-byte aligned (per-row shift, 7 units
// gathered, 4 stored); 2 = sum map, scalar stores (32 floats per tile); 3 / 4 = class map and / or count, 1 / 4 cells per lane.
constexpr int kSegPatches = 64;     // patches per tile staged by the segment kernel (longer lists: the global-memory path); 5.6 KB per warp -> 4 CTAs per SM
constexpr int kSegCap = 264;        // segments per tile: <= cells of the tile + 1
constexpr int kSegVals = 288;       // floats per value buffer: segments * n <= tile units + 3 n
constexpr int kSegWords = 16;       // boundary bitmap words (<= 512 tile cells)

__host__ __device__ inline int seg_warp_smem_bytes() {
    return kSegPatches * 8 /* r0 r1 c0 c1 (u16) */ + kSegPatches * 32 /* logits, stride 8 */ + (2 * kSegWords + 4) * 4 /* bitmaps */ + kSegCap * 2 /* segments */ +
           144 /* runs */ + 2 * kSegVals * 4 /* values */;
}

template <int KIND, int G>
__global__ void __launch_bounds__(kBinWarps * 32) bin_seg_kernel(const float* __restrict__ logits, const int32_t* __restrict__ coords, BinGeom g,
                                                                 const uint32_t* __restrict__ off, const uint32_t* __restrict__ len,
                                                                 const uint32_t* __restrict__ list, uint32_t* __restrict__ sorted,
                                                                 float* __restrict__ sum_map, uint32_t* __restrict__ count_map,
                                                                 uint8_t* __restrict__ argmax_map) {
    constexpr bool CELL = KIND >= 3;
    constexpr int VEC = (KIND == 0 || KIND == 1 || KIND == 4) ? 4 : 1;
    constexpr int K = KIND == 0 ? 4 * G : (KIND == 1 ? 7 : (KIND == 4 ? 4 : 1));   // units a lane fetches per run
    extern __shared__ __align__(16) unsigned char bin_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t ctas_x = (g.ntx + kBinWarps - 1) / kBinWarps;
    const int64_t ty = blockIdx.x / ctas_x;
    const int64_t tx = (blockIdx.x % ctas_x) * kBinWarps + w;
    if (tx >= g.ntx) return;
    const int64_t t = ty * g.ntx + tx;
    pdl_wait();                                             // the fill pass (and through it count / alloc) has completed
    const uint32_t beg = off[t];
    const int L = (int)len[t];
    const int n = g.n;
    if (L > kSegPatches) {
        bin_tile_slow<VEC, CELL>(logits, coords, g, beg, L, list, sorted, sum_map, count_map, argmax_map, ty, tx);
        return;
    }
    unsigned char* base = bin_smem + (size_t)w * seg_warp_smem_bytes();
    uint16_t* s_r0 = reinterpret_cast<uint16_t*>(base);
    uint16_t* s_r1 = s_r0 + kSegPatches;
    uint16_t* s_c0 = s_r1 + kSegPatches;
    uint16_t* s_c1 = s_c0 + kSegPatches;
    float* s_lg = reinterpret_cast<float*>(base + kSegPatches * 8);                      // [L][8]
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(base + kSegPatches * 40);            // [kSegWords] segment-start bitmap over the tile's cells
    uint32_t* s_wpre = s_bits + kSegWords;                                          // [kSegWords] set bits before each word
    uint32_t* s_rbits = s_wpre + kSegWords;                                         // [4] run-start bitmap over the tile's rows (TH <= 128)
    uint16_t* s_segc = reinterpret_cast<uint16_t*>(s_rbits + 4);                    // [S] first cell of each segment (tile relative)
    uint8_t* s_run = reinterpret_cast<uint8_t*>(s_segc + kSegCap);                  // [NR + 1] first row of each run, then TH
    float* s_val = reinterpret_cast<float*>(base + kSegPatches * 40 + (2 * kSegWords + 4) * 4 + kSegCap * 2 + 144);  // [2][kSegVals] double-buffered
    uint32_t* s_raw = reinterpret_cast<uint32_t*>(s_val);                            // staging only: the tile's patch indices (unordered) alias the values

    const int R0 = (int)(ty * g.TH);
    const int TH = (int)((int64_t)R0 + g.TH < g.rows ? g.TH : g.rows - R0);         // rows of this tile
    const int RF = (int)g.units_per_row;
    const int U0 = (int)(tx * g.TW);
    const int U1 = (KIND == 1) ? (U0 + g.TW + 3 < RF ? U0 + g.TW + 3 : RF) : (U0 + g.TW < RF ? U0 + g.TW : RF);   // phased tiles reach 3 units further
    const int scale = g.scale;                                                      // units per cell: n (sum) or 1 (cell)
    const int C0 = U0 / scale;
    const int ncells = (U1 + scale - 1) / scale - C0;

    // ---- stage: footprints clipped to the tile (rows / cells relative to the tile) and logits, filed in ascending patch index. A lane
    // owns the list entries lane and lane + 32: it loads their origins and logits straight after the index (dependent chain off/len ->
    // list -> coords, logits: one level shorter than sorting first) and ranks the indices while those loads are in flight.
    const bool want_lg = !CELL || argmax_map != nullptr;
    uint32_t myid[2];
    int py[2] = {0, 0}, px[2] = {0, 0};
    float lgv[2][8];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const int j = lane + 32 * sl;
        myid[sl] = j < L ? list[beg + j] : 0xffffffffu;
        if (j < L) s_raw[j] = myid[sl];
    }
    if (lane < kSegWords) s_bits[lane] = lane == 0 ? 1u : 0u;                       // cell 0 starts segment 0
    if (lane < 4) s_rbits[lane] = lane == 0 ? 1u : 0u;                              // row 0 starts run 0
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        if (lane + 32 * sl < L) {
            py[sl] = __ldg(coords + 2 * (int64_t)myid[sl]);
            px[sl] = __ldg(coords + 2 * (int64_t)myid[sl] + 1);
            if (want_lg) {
#pragma unroll
                for (int q = 0; q < 8; ++q) lgv[sl][q] = q < n ? __ldg(logits + (int64_t)myid[sl] * n + q) : 0.f;   // zero padded: the adds below are unconditional in q
            }
        }
    }
    __syncwarp();
    int rank[2] = {0, 0};
    for (int i = 0; i < L; ++i) {                                                   // patch indices are distinct within a tile
        const uint32_t o = s_raw[i];
        rank[0] += o < myid[0];
        rank[1] += o < myid[1];
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        if (lane + 32 * sl < L) {
            const int j = rank[sl];
            BinRec f;
            f.r0 = f.r1 = f.u0 = f.u1 = 0;
            bin_footprint(g, py[sl], px[sl], f);
            const int a = f.r0 - R0, b = f.r1 - R0;
            s_r0[j] = (uint16_t)(a < 0 ? 0 : a);
            s_r1[j] = (uint16_t)(b > TH ? TH : (b < 0 ? 0 : b));
            if (a > 0 && a < TH) atomicOr(s_rbits + (a >> 5), 1u << (a & 31));         // the covering set changes at every footprint edge
            if (b > 0 && b < TH) atomicOr(s_rbits + (b >> 5), 1u << (b & 31));
            int c0 = f.u0 / scale - C0, c1 = f.u1 / scale - C0;                      // footprints are whole cells: u0, u1 are multiples of scale
            c0 = c0 < 0 ? 0 : (c0 > ncells ? ncells : c0);
            c1 = c1 < 0 ? 0 : (c1 > ncells ? ncells : c1);
            s_c0[j] = (uint16_t)c0;
            s_c1[j] = (uint16_t)c1;
            if (c0 > 0 && c0 < ncells) atomicOr(s_bits + (c0 >> 5), 1u << (c0 & 31));
            if (c1 > 0 && c1 < ncells) atomicOr(s_bits + (c1 >> 5), 1u << (c1 & 31));
            if (want_lg) {
                *reinterpret_cast<float4*>(s_lg + j * 8) = make_float4(lgv[sl][0], lgv[sl][1], lgv[sl][2], lgv[sl][3]);
                *reinterpret_cast<float4*>(s_lg + j * 8 + 4) = make_float4(lgv[sl][4], lgv[sl][5], lgv[sl][6], lgv[sl][7]);
            }
        }
    }
    __syncwarp();
    // ---- segments: prefix popcounts of the bitmap words, first cell of every segment
    int S;
    {
        const uint32_t word = lane < kSegWords ? s_bits[lane] : 0u;
        const uint32_t pc = __popc(word);
        uint32_t inc = pc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        S = (int)__shfl_sync(0xffffffffu, inc, 31);
        if (lane < kSegWords) {
            s_wpre[lane] = inc - pc;
            uint32_t rest = word;
            int k = (int)(inc - pc);
            while (rest) {
                s_segc[k++] = (uint16_t)(lane * 32 + __ffs(rest) - 1);
                rest &= rest - 1;
            }
        }
    }
    int NR;         // row runs: maximal row ranges with the same covering patches
    {
        const uint32_t rw = lane < 4 ? s_rbits[lane] : 0u;
        const int pc = __popc(rw);
        int before = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int v = __shfl_sync(0xffffffffu, pc, k);
            if (lane > k) before += v;
        }
        NR = before + pc;                                     // lane 3 holds the total
        NR = __shfl_sync(0xffffffffu, NR, 3);
        if (lane < 4) {
            uint32_t rest = rw;
            int k = before;
            while (rest) {
                s_run[k++] = (uint8_t)(lane * 32 + __ffs(rest) - 1);
                rest &= rest - 1;
            }
            if (lane == 3) s_run[NR] = (uint8_t)TH;
        }
    }
    __syncwarp();   // also: every lane is done with s_raw before the value buffers (which alias it) are written
    // ---- the lane's output units and their slots in the value buffer
    const int ub = U0 + lane * VEC;
    int slot[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int u = ub + (KIND == 0 ? (k >> 2) * (32 * VEC) + (k & 3) : k);
        int sl = 0;
        if (u < U1) {
            const int c = u / scale, rel = c - C0;
            const int seg = (int)s_wpre[rel >> 5] + __popc(s_bits[rel >> 5] & (0xffffffffu >> (31 - (rel & 31)))) - 1;
            sl = CELL ? seg : seg * n + (u - c * scale);
        }
        slot[k] = sl;
    }
    const bool want_vals = CELL ? argmax_map != nullptr : true;
    const int pstep = RF & 3;

    // fetch this lane's units of one run from `val` and stream them to every row of the run [r, next)
    auto emit = [&](int r, int next, const float* val) {
        if constexpr (CELL) {
            uint32_t pk[K];
#pragma unroll
            for (int k = 0; k < K; ++k) pk[k] = reinterpret_cast<const uint32_t*>(val)[slot[k]];
            if (ub < U1) {                                          // VEC == 4: dw % 4 == 0 and aligned maps (host check)
                int64_t o = (int64_t)(R0 + r) * RF + ub;
                for (int rr = r; rr < next; ++rr, o += RF) {
                    if constexpr (KIND == 4) {
                        if (argmax_map) __stcs(reinterpret_cast<uchar4*>(argmax_map + o), make_uchar4(pk[0] & 255u, pk[1] & 255u, pk[2] & 255u, pk[3] & 255u));
                        if (count_map) __stcs(reinterpret_cast<uint4*>(count_map + o), make_uint4(pk[0] >> 8, pk[1] >> 8, pk[2] >> 8, pk[3] >> 8));
                    } else {
                        if (argmax_map) __stcs(argmax_map + o, (uint8_t)(pk[0] & 255u));
                        if (count_map) __stcs(count_map + o, pk[0] >> 8);
                    }
                }
            }
        } else {
            float v[K];
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = val[slot[k]];
            if constexpr (KIND == 0) {
                float* o = sum_map + (int64_t)(R0 + r) * RF + ub;
                for (int rr = r; rr < next; ++rr, o += RF) {
#pragma unroll
                    for (int gi = 0; gi < G; ++gi)
                        if (ub + gi * 128 < U1) __stcs(reinterpret_cast<float4*>(o + gi * 128), make_float4(v[4 * gi], v[4 * gi + 1], v[4 * gi + 2], v[4 * gi + 3]));
                }
            } else if constexpr (KIND == 2) {
                float* o = sum_map + (int64_t)(R0 + r) * RF + ub;
                if (ub < U1)
                    for (int rr = r; rr < next; ++rr, o += RF) __stcs(o, v[0]);
            } else {
                // rows not 16-byte aligned: in row rr the lane's vector starts sft = (4 - rr * RF) mod 4 units later (then 16-byte aligned);
                // the first sft units of a row are scalar stores of tile 0, the last vector of a row is clipped
                float* rowp = sum_map + (int64_t)(R0 + r) * RF;
                int p = (int)(((int64_t)(R0 + r) * RF) & 3);         // sum_map is 16-byte aligned (host check)
                for (int rr = r; rr < next; ++rr) {
                    const int sft = (4 - p) & 3;
                    float w0, w1, w2, w3;
                    switch (sft) {
                        case 0: w0 = v[0]; w1 = v[1]; w2 = v[2]; w3 = v[3]; break;
                        case 1: w0 = v[1]; w1 = v[2]; w2 = v[3]; w3 = v[4]; break;
                        case 2: w0 = v[2]; w1 = v[3]; w2 = v[4]; w3 = v[5]; break;
                        default: w0 = v[3]; w1 = v[4]; w2 = v[5]; w3 = v[6]; break;
                    }
                    const int u = ub + sft;
                    if (u + 4 <= RF) {
                        __stcs(reinterpret_cast<float4*>(rowp + u), make_float4(w0, w1, w2, w3));
                    } else if (u < RF) {
                        rowp[u] = w0;
                        if (u + 1 < RF) rowp[u + 1] = w1;
                        if (u + 2 < RF) rowp[u + 2] = w2;
                    }
                    if (tx == 0 && lane == 0) {
                        if (sft > 0) rowp[0] = v[0];
                        if (sft > 1) rowp[1] = v[1];
                        if (sft > 2) rowp[2] = v[2];
                    }
                    rowp += RF;
                    p = (p + pstep) & 3;
                }
            }
        }
    };

    // ---- passes: a lane owns one REGION = (run, segment) of the pass -- RPP = 32 / S runs at a time (one run, segments in chunks of 32,
    // when a tile has more than 32 segments). The patch loop walks the patches that cover ANY run of the pass (consecutive runs differ
    // by one patch each), in ascending list index = the reference's order; a lane adds those that cover its own row and columns.
    const bool multi = S <= 32;
    const int RPP = multi ? 32 / S : 1;
    const int k_l = multi ? lane / S : 0;
    const int seg_l = multi ? lane - k_l * S : 0;
    const int vstride = CELL ? S : S * n;                         // value-buffer stride between the runs of a pass
    int buf = 0;
    for (int run0 = 0; run0 < NR; run0 += RPP) {
        float* const val = s_val + buf * kSegVals;
        const int nrun = NR - run0 < RPP ? NR - run0 : RPP;
        const int row_first = (int)s_run[run0], row_last = (int)s_run[run0 + nrun - 1];
        for (int s0 = 0; s0 < S; s0 += 32) {                      // one iteration when multi
            const int sg = multi ? seg_l : s0 + lane;
            const bool live = multi ? k_l < nrun : sg < S;
            const int myrow = live ? (int)s_run[run0 + k_l] : 0x7fff;
            const int sc = live ? (int)s_segc[sg] : 0xffff;
            float acc[kBinMaxN];
#pragma unroll
            for (int q = 0; q < kBinMaxN; ++q) acc[q] = 0.f;
            uint32_t hits = 0;
            for (int j0 = 0; j0 < L; j0 += 32) {
                const int j = j0 + lane;
                const int a = j < L ? (int)s_r0[j] : 0x7fff, b = j < L ? (int)s_r1[j] : 0;
                unsigned m = __ballot_sync(0xffffffffu, a <= row_last && b > row_first);
                while (m) {  // warp-uniform trip count
                    const int jj = j0 + __ffs(m) - 1;
                    m &= m - 1;
                    const bool cov = (int)s_r0[jj] <= myrow && myrow < (int)s_r1[jj] && (int)s_c0[jj] <= sc && sc < (int)s_c1[jj];
                    hits += cov;
                    if (want_vals) {
                        const float4 lo = *reinterpret_cast<const float4*>(s_lg + jj * 8);
                        // adding under the predicate only: a sum that starts at +0.0f and adds v is the reference's 0.0 + v
                        // (classes q >= n add the +0.0f padding to sums nobody reads)
                        if (cov) { acc[0] += lo.x; acc[1] += lo.y; acc[2] += lo.z; acc[3] += lo.w; }
                        if (n > 4) {
                            const float4 hi = *reinterpret_cast<const float4*>(s_lg + jj * 8 + 4);
                            if (cov) { acc[4] += hi.x; acc[5] += hi.y; acc[6] += hi.z; acc[7] += hi.w; }
                        }
                    }
                }
            }
            if (live) {
                const int region = multi ? lane : sg;             // (run k_l, segment seg_l) -> k_l * S + seg_l == lane
                if constexpr (CELL) {
                    const uint32_t am = want_vals ? (uint32_t)first_argmax(acc, n) : 0u;
                    reinterpret_cast<uint32_t*>(val)[region] = am | (hits << 8);
                } else {
#pragma unroll
                    for (int q = 0; q < kBinMaxN; ++q)
                        if (q < n) val[region * n + q] = acc[q];
                }
            }
        }
        __syncwarp();
        for (int k = 0; k < nrun; ++k) emit((int)s_run[run0 + k], (int)s_run[run0 + k + 1], val + k * vstride);
        buf ^= 1;   // the next pass writes the other buffer: one warp barrier per pass is enough
    }
}

// ---- cell-lane formulation (sum map, n <= 8 classes known at compile time) ----------------------------------------------------
// A lane owns one CELL of the tile (32 cells x TH rows) and keeps its N class sums in registers: per covering patch one coverage
// test (a precomputed 32-bit lane mask) and N adds, instead of a test and an add per output FLOAT as in the row-run kernels (whose
// 4 floats per lane straddle two cells). The run's row image (32 * N floats) goes through shared memory once per run, every lane
// picks up its 16-byte vectors and streams them to the rows of the run. Runs come from a row bitmap built while staging.
// PHASED (rows not 16-byte aligned): a tile is TW = 4 * floor((31 N - 2) / 4) units wide, so that the units [U0, U0 + TW + 3) any of
// its shifted vectors can touch lie inside 32 cells; the vectors of a row are read from the row image at the row's shift.
constexpr int kCellRunBytes = 160;   // [4] row bitmap words + [TH + 1 <= 129] run starts (u8)
constexpr int kCellCap = 128;        // patches per tile staged in shared memory (longer lists: the global-memory path); 6.7 KB per warp -> 4 CTAs per SM

__host__ __device__ inline int cell_warp_smem_bytes() {
    return 1024 /* ids while staging, then the run's row image */ + kCellCap * 8 /* rows */ + kCellCap * 4 /* lane masks */ + kCellCap * 32 /* logits, stride 8 */ +
           kCellRunBytes;
}

__host__ __device__ inline int cell_tile_units(int n, bool phased) { return phased ? (31 * n - 2) / 4 * 4 : 32 * n; }

template <int N>
__device__ __forceinline__ void cell_add(float (&acc)[N], bool cov, const float4& lo, const float4& hi) {
    if (cov) {
        acc[0] += lo.x;
        if (N > 1) acc[N > 1 ? 1 : 0] += lo.y;
        if (N > 2) acc[N > 2 ? 2 : 0] += lo.z;
        if (N > 3) acc[N > 3 ? 3 : 0] += lo.w;
        if (N > 4) acc[N > 4 ? 4 : 0] += hi.x;
        if (N > 5) acc[N > 5 ? 5 : 0] += hi.y;
        if (N > 6) acc[N > 6 ? 6 : 0] += hi.z;
        if (N > 7) acc[N > 7 ? 7 : 0] += hi.w;
    }
}

template <int N, bool PHASED, bool NOSTORE = false>
__global__ void __launch_bounds__(kBinWarps * 32) bin_cell_sum_kernel(const float* __restrict__ logits, const int32_t* __restrict__ coords, BinGeom g,
                                                                      const uint32_t* __restrict__ off, const uint32_t* __restrict__ len,
                                                                      const uint32_t* __restrict__ list, uint32_t* __restrict__ sorted,
                                                                      float* __restrict__ sum_map, uint32_t* __restrict__ count_map,
                                                                      uint8_t* __restrict__ argmax_map) {
    // outputs: the sum map (may be null), and / or the per-cell count and class (first arg max of the lane's own sums): a lane owns a
    // cell, so these cost one compare chain and one byte / word store per row -- no second binning + tile pass for them
    constexpr int LS = 8;                                   // logits row stride in shared memory
    constexpr int TWU = PHASED ? (31 * N - 2) / 4 * 4 : 32 * N;
    constexpr int NV = TWU / 4;                             // 16-byte vectors per tile row
    constexpr int V = (NV + 31) / 32;                       // vectors per lane and row
    extern __shared__ __align__(16) unsigned char bin_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t ctas_x = (g.ntx + kBinWarps - 1) / kBinWarps;
    const int64_t ty = blockIdx.x / ctas_x;
    const int64_t tx = (blockIdx.x % ctas_x) * kBinWarps + w;
    if (tx >= g.ntx) return;
    const int64_t t = ty * g.ntx + tx;
    pdl_wait();                                             // the fill pass (and through it count / alloc) has completed
    const uint32_t beg = off[t];
    const int L = (int)len[t];
    if (L > kCellCap) {
        if (sum_map) bin_tile_slow<1, false>(logits, coords, g, beg, L, list, sorted, sum_map, nullptr, nullptr, ty, tx);
        if (count_map || argmax_map) {     // the same tile in cell units: 32 cells per tile row (!PHASED: host check)
            BinGeom gc = g;
            gc.scale = 1; gc.units_per_row = g.dw; gc.TW = 32;
            bin_tile_slow<1, true>(logits, coords, gc, beg, L, list, sorted, nullptr, count_map, argmax_map, ty, tx);
        }
        return;
    }
    unsigned char* base = bin_smem + (size_t)w * cell_warp_smem_bytes();
    uint32_t* s_raw = reinterpret_cast<uint32_t*>(base);               // [kCellCap] the tile's patch indices, unordered (staging only)
    float* s_out = reinterpret_cast<float*>(base);                     // [32 * N <= 256] the run's row image (aliases the indices)
    int2* s_rows = reinterpret_cast<int2*>(base + 1024);               // [L] tile-relative rows [r0, r1), clipped
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(base + 1024 + kCellCap * 8);   // [L] lanes (cells) the patch covers
    float* s_lg = reinterpret_cast<float*>(base + 1024 + kCellCap * 12);           // [L][8]
    uint32_t* s_rbits = reinterpret_cast<uint32_t*>(base + 1024 + kCellCap * 44);  // [4] run-start bitmap over the tile's rows (TH <= 128)
    uint8_t* s_run = reinterpret_cast<uint8_t*>(s_rbits + 4);                     // [NR + 1] first row of each run, then TH

    const int R0 = (int)(ty * g.TH);
    const int TH = (int)((int64_t)R0 + g.TH < g.rows ? g.TH : g.rows - R0);
    const int RF = (int)g.units_per_row;
    const int U0 = (int)(tx * TWU);
    const int C0 = U0 / N;                                   // the tile's first cell; lane l owns cell C0 + l

    // ---- stage. In chunks of 64 list entries a lane owns the patches lane and lane + 32: it loads their origins and logits straight
    // after the index (dependent chain: off/len -> list -> coords, logits -- one level shorter than sorting first), ranks them by
    // patch index while those loads are in flight, and files footprint, lane mask and logits at the rank.
    for (int j = lane; j < L; j += 32) s_raw[j] = list[beg + j];
    if (lane < 4) s_rbits[lane] = lane == 0 ? 1u : 0u;       // row 0 starts run 0
    __syncwarp();
    for (int ch = 0; ch < L; ch += 64) {
        uint32_t myid[2];
        int py[2], px[2];
        float lgv[2][LS];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const int j = ch + lane + 32 * sl;
            myid[sl] = j < L ? s_raw[j] : 0xffffffffu;
            if (j < L) {
                py[sl] = __ldg(coords + 2 * (int64_t)myid[sl]);
                px[sl] = __ldg(coords + 2 * (int64_t)myid[sl] + 1);
#pragma unroll
                for (int q = 0; q < LS; ++q) lgv[sl][q] = q < N ? __ldg(logits + (int64_t)myid[sl] * N + q) : 0.f;
            }
        }
        int rank[2] = {0, 0};
        for (int i = 0; i < L; ++i) {                         // patch indices are distinct within a tile
            const uint32_t o = s_raw[i];
            rank[0] += o < myid[0];
            rank[1] += o < myid[1];
        }
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            if (ch + lane + 32 * sl < L) {
                const int k = rank[sl];
                BinRec f;
                f.r0 = f.r1 = f.u0 = f.u1 = 0;
                bin_footprint(g, py[sl], px[sl], f);
                int a = f.r0 - R0, b = f.r1 - R0;
                a = a < 0 ? 0 : a;
                b = b > TH ? TH : (b < 0 ? 0 : b);
                s_rows[k] = make_int2(a, b);
                if (a > 0 && a < TH) atomicOr(s_rbits + (a >> 5), 1u << (a & 31));  // the covering set changes at every footprint edge
                if (b > 0 && b < TH) atomicOr(s_rbits + (b >> 5), 1u << (b & 31));
                int c0 = f.u0 / N - C0, c1 = f.u1 / N - C0;                         // footprints are whole cells: u0, u1 are multiples of N
                c0 = c0 < 0 ? 0 : (c0 > 32 ? 32 : c0);
                c1 = c1 < 0 ? 0 : (c1 > 32 ? 32 : c1);
                const int wd = c1 - c0;
                s_mask[k] = wd <= 0 ? 0u : ((wd >= 32 ? 0xffffffffu : ((1u << wd) - 1u)) << c0);
                *reinterpret_cast<float4*>(s_lg + k * LS) = make_float4(lgv[sl][0], lgv[sl][1], lgv[sl][2], lgv[sl][3]);
                if constexpr (N > 4) *reinterpret_cast<float4*>(s_lg + k * LS + 4) = make_float4(lgv[sl][4], lgv[sl][5], lgv[sl][6], lgv[sl][7]);
            }
        }
    }
    __syncwarp();
    int NR;         // row runs: maximal row ranges with the same covering patches
    {
        const uint32_t rw = lane < 4 ? s_rbits[lane] : 0u;
        const int pc = __popc(rw);
        int before = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int v = __shfl_sync(0xffffffffu, pc, k);
            if (lane > k) before += v;
        }
        NR = __shfl_sync(0xffffffffu, before + pc, 3);
        if (lane < 4) {
            uint32_t rest = rw;
            int k = before;
            while (rest) {
                s_run[k++] = (uint8_t)(lane * 32 + __ffs(rest) - 1);
                rest &= rest - 1;
            }
            if (lane == 3) s_run[NR] = (uint8_t)TH;
        }
    }
    __syncwarp();   // also: every lane is done with s_raw before the row image (which aliases it) is written

    const uint32_t lanebit = 1u << lane;
    // SPARSE (footprints much narrower than the tile, e.g. 14 of 32 cells at d = 16): a lane is covered by a third of the patches that
    // cover its row. Each lane gets the set of patches that cover its COLUMN as bit words (a 32 x 32 bit transpose of the lane masks by
    // ballots, once per tile); per run it walks `row set & column set` -- its own covering patches only, in ascending index -- instead
    // of every patch of the row under a predicate. The trip count is the largest per-lane count of the warp (divergent loop).
    const bool sparse = g.sparse != 0;
    uint32_t colw[kCellCap / 32];
#pragma unroll
    for (int w = 0; w < kCellCap / 32; ++w) {
        colw[w] = 0u;
        if (sparse && w * 32 < L) {                          // warp-uniform
            const int j = w * 32 + lane;
            const uint32_t mk = j < L ? s_mask[j] : 0u;
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                const uint32_t b = __ballot_sync(0xffffffffu, (mk >> l) & 1u);
                if (lane == l) colw[w] = b;
            }
        }
    }
    const int img0 = U0 - C0 * N;                            // the tile's first unit inside the row image (0 when aligned)
    bool vfull[V];                                           // aligned rows: the lane's vectors that exist (RF % 4 == 0: whole or not at all)
#pragma unroll
    for (int i = 0; i < V; ++i) vfull[i] = lane + 32 * i < NV && U0 + 4 * (lane + 32 * i) < RF;
    const int pstep = RF & 3;

    for (int k = 0; k < NR; ++k) {
        const int r = (int)s_run[k];
        int next = (int)s_run[k + 1];
        float acc[N];
#pragma unroll
        for (int q = 0; q < N; ++q) acc[q] = 0.f;
        uint32_t hits = 0;
        if (sparse) {
#pragma unroll
            for (int w = 0; w < kCellCap / 32; ++w) {
                if (w * 32 >= L) break;                      // warp-uniform
                const int j = w * 32 + lane;
                const int2 rw = j < L ? s_rows[j] : make_int2(0x7fff, 0);
                unsigned mine = __ballot_sync(0xffffffffu, rw.x <= r && r < rw.y) & colw[w];
                while (mine) {  // this lane's covering patches in ascending list index = the reference's order
                    const int jj = w * 32 + __ffs(mine) - 1;
                    mine &= mine - 1;
                    const float4 lo = *reinterpret_cast<const float4*>(s_lg + jj * LS);
                    float4 hi = make_float4(0.f, 0.f, 0.f, 0.f);
                    if constexpr (N > 4) hi = *reinterpret_cast<const float4*>(s_lg + jj * LS + 4);
                    cell_add<N>(acc, true, lo, hi);
                    ++hits;
                }
            }
        } else
        for (int j0 = 0; j0 < L; j0 += 32) {
            const int j = j0 + lane;
            const int2 rw = j < L ? s_rows[j] : make_int2(0x7fff, 0);
            unsigned m = __ballot_sync(0xffffffffu, rw.x <= r && r < rw.y);
            while (m) {  // covering patches in ascending list index = the reference's order, two per trip (independent shared-memory
                         // loads); warp-uniform trip count
                const int ja = j0 + __ffs(m) - 1;
                m &= m - 1;
                const bool two = m != 0u;
                const int jb = two ? j0 + __ffs(m) - 1 : ja;
                m &= m - 1;
                const bool cova = (s_mask[ja] & lanebit) != 0u;
                const bool covb = two && (s_mask[jb] & lanebit) != 0u;
                // adding under the predicate only: a sum that starts at +0.0f and adds v is the reference's 0.0 + v
                const float4 loa = *reinterpret_cast<const float4*>(s_lg + ja * LS);
                const float4 lob = *reinterpret_cast<const float4*>(s_lg + jb * LS);
                float4 hia = make_float4(0.f, 0.f, 0.f, 0.f), hib = hia;
                if constexpr (N > 4) {
                    hia = *reinterpret_cast<const float4*>(s_lg + ja * LS + 4);
                    hib = *reinterpret_cast<const float4*>(s_lg + jb * LS + 4);
                }
                cell_add<N>(acc, cova, loa, hia);
                cell_add<N>(acc, covb, lob, hib);
                hits += (cova ? 1u : 0u) + (covb ? 1u : 0u);
            }
        }
        if (count_map != nullptr || argmax_map != nullptr) {  // per-cell outputs straight from the lane's registers
            int best_c = 0;
            float best = acc[0];
#pragma unroll
            for (int c = 1; c < N; ++c)                       // np.argmax: first maximum; NaN counts as the maximum (first NaN wins)
                if (acc[c] > best || (acc[c] != acc[c] && best == best)) { best = acc[c]; best_c = c; }
            const int cell = C0 + lane;
            if (cell < (int)g.dw) {
                int64_t o = (int64_t)(R0 + r) * g.dw + cell;
                for (int rr = r; rr < next; ++rr, o += g.dw) {
                    if (argmax_map) __stcs(argmax_map + o, (uint8_t)best_c);
                    if (count_map) __stcs(count_map + o, hits);
                }
            }
        }
        if (sum_map == nullptr) continue;                     // warp-uniform
        __syncwarp();                                        // the previous run's image has been read by every lane
#pragma unroll
        for (int q = 0; q < N; ++q) s_out[lane * N + q] = acc[q];
        __syncwarp();
        if constexpr (!PHASED) {
            float4 v[V];
#pragma unroll
            for (int i = 0; i < V; ++i)
                if (lane + 32 * i < NV) v[i] = reinterpret_cast<const float4*>(s_out)[lane + 32 * i];
            float* o = sum_map + (int64_t)(R0 + r) * RF + U0 + 4 * lane;
            if (NOSTORE && RF > 0) next = r + (v[0].x == 12345.f && v[V - 1].w == 54321.f);   // profiling variant 4: everything but the stores
            for (int rr = r; rr < next; ++rr, o += RF) {
#pragma unroll
                for (int i = 0; i < V; ++i)
                    if (vfull[i]) __stcs(reinterpret_cast<float4*>(o + 128 * i), v[i]);
            }
        } else {
            // in row rr the tile's vectors start sft = (4 - rr * RF) mod 4 units later (then 16-byte aligned); the first sft units of a
            // row are scalar stores of tile 0, the last vector of a row is clipped
            float* rowp = sum_map + (int64_t)(R0 + r) * RF;
            int p = (int)(((int64_t)(R0 + r) * RF) & 3);     // sum_map is 16-byte aligned (host check)
            for (int rr = r; rr < next; ++rr) {
                const int sft = (4 - p) & 3;
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const int f = lane + 32 * i;
                    const int u = U0 + sft + 4 * f;
                    if (f < NV && u < RF) {
                        const float* s = s_out + img0 + sft + 4 * f;
                        const float w0 = s[0], w1 = s[1], w2 = s[2], w3 = s[3];   // inside the 32-cell image (tile width rule above)
                        if (u + 4 <= RF) {
                            __stcs(reinterpret_cast<float4*>(rowp + u), make_float4(w0, w1, w2, w3));
                        } else {
                            rowp[u] = w0;
                            if (u + 1 < RF) rowp[u + 1] = w1;
                            if (u + 2 < RF) rowp[u + 2] = w2;
                        }
                    }
                }
                if (tx == 0 && lane < sft) rowp[lane] = s_out[lane];
                rowp += RF;
                p = (p + pstep) & 3;
            }
        }
    }
}

// profiling override: rows + 1000 * groups + 100000 * extra shared-memory KB per CTA (each field 0 = heuristic / none)
static int tile_rows_for(int ps, int d) {
    if (g_bin_tile_rows % 1000 > 0) return g_bin_tile_rows % 1000;
    const int fh = ps / d;
    return fh >= 48 ? 64 : fh >= 24 ? 32 : 16;
}

static int groups_for(bool cell, int vec, int ps, int d, int n) {
    if (cell || vec != 4) return 1;
    const int ov = (g_bin_tile_rows / 1000) % 10;
    if (ov > 0) return ov >= 2 ? 2 : 1;
    return (int64_t)(ps / d) * n >= 512 ? 2 : 1;  // wide footprints: 256-float tiles halve the per-tile staging work (measured: profiles/r01_stitch.md)
}

static BinGeom make_geom(bool cell, int vec, int ps, int d, int n, int64_t rows, int64_t dw, int64_t row_offset, bool phased = false,
                         bool cell_lane = false) {
    BinGeom g;
    g.phased = phased ? 1 : 0;
    g.sparse = 0;
    g.rows = rows; g.row_offset = row_offset; g.dw = dw;
    g.ps = ps; g.d = d; g.n = n;
    g.scale = cell ? 1 : n;
    g.units_per_row = dw * g.scale;
    g.TH = tile_rows_for(ps, d);
    g.G = phased ? 1 : groups_for(cell, vec, ps, d, n);
    g.TW = 32 * vec * g.G;
    if (cell_lane) {                                                  // bin_cell_sum_kernel: 32 cells per tile
        g.G = 1;
        g.TW = cell_tile_units(n, phased);
        g.sparse = (g_bin_sparse < 0 ? ps / d < 24 : g_bin_sparse) ? 1 : 0;
    }
    g.nty = (rows + g.TH - 1) / g.TH;
    g.ntx = (g.units_per_row + g.TW - 1) / g.TW;
    return g;
}

static int64_t entries_cap(const BinGeom& g, int64_t P) {
    const int64_t fh = g.ps / g.d + 1, fw = (int64_t)(g.ps / g.d + 1) * g.scale + (g.phased ? 3 : 0);  // largest footprint, rows / units
    const int64_t tyx = (fh - 1 + g.TH - 1) / g.TH + 1, txx = (fw - 1 + g.TW - 1) / g.TW + 1;
    return P * tyx * txx;
}

struct BinScratch {
    uint32_t *cnt, *off, *len, *list, *sorted;
    int64_t total_bytes;
};

static size_t roundup256(size_t v) { return (v + 255) / 256 * 256; }

static void carve_bin(const BinGeom& g, int64_t P, void* base, BinScratch& s) {
    const int64_t ntiles = g.nty * g.ntx, cap = entries_cap(g, P);
    uintptr_t p = reinterpret_cast<uintptr_t>(base);
    size_t o = 0;
    s.cnt = reinterpret_cast<uint32_t*>(p + o); o += roundup256((ntiles + 1) * 4);  // [ntiles] counters + the list cursor
    s.off = reinterpret_cast<uint32_t*>(p + o); o += roundup256(ntiles * 4);
    s.len = reinterpret_cast<uint32_t*>(p + o); o += roundup256(ntiles * 4);
    s.list = reinterpret_cast<uint32_t*>(p + o); o += roundup256(cap * 4);
    s.sorted = reinterpret_cast<uint32_t*>(p + o); o += roundup256(cap * 4);
    s.total_bytes = (int64_t)o;
}

static bool sum_vec4(const float* sum_map, int64_t dw, int n) { return (dw * n) % 4 == 0 && reinterpret_cast<uintptr_t>(sum_map) % 16 == 0; }

static int g_bin_pdl = -1;   // programmatic dependent launch between the kernels of a call: -1 = read DH_BIN_PDL once (default on)

template <typename... KArgs, typename... Args>
static cudaError_t launch_dependent(bool pdl, void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args) {
    if (g_bin_pdl < 0) {
        const char* e = getenv("DH_BIN_PDL");
        g_bin_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_bin_pdl && pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// SEGK >= 0: bin_seg_kernel<SEGK, G> (n <= kBinMaxN); -1: the row-run kernels
template <int VEC, int G, bool CELL, bool STAGED, bool PHASED = false, int SEGK = -1, int CELLN = 0>
static int run_binned(const float* logits, const int32_t* coords, int64_t P, const BinGeom& g, float* sum_map, uint32_t* count_map,
                      uint8_t* argmax_u8, void* scratch, int64_t scratch_bytes, cudaStream_t st) {
    BinScratch s;
    carve_bin(g, P, scratch, s);
    DH_REQUIRE(s.total_bytes <= scratch_bytes, "dh_stitch_binned: scratch too small (%lld bytes, %lld needed)", (long long)scratch_bytes,
               (long long)s.total_bytes);
    const int64_t ntiles = g.nty * g.ntx;
    DH_REQUIRE(ntiles < (1ll << 31) - 1 && entries_cap(g, P) < (1ll << 32), "dh_stitch_binned: map or patch list too large");
    // dependent launches pay off while the tile kernel is short (40 000^2 list: d = 16 +5 %, d = 4 +1.5 %, d = 2 +-0, d = 1 -0.6 %)
    const bool pdl = g.rows * g.units_per_row * (g.scale == 1 ? 1 : 4) <= (4ll << 30);
    cudaError_t e = cudaMemsetAsync(s.cnt, 0, (ntiles + 1) * 4, st);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
    const int64_t pb = (P + 255) / 256;
    const int pgrid = (int)(pb < (int64_t)kNumSMs * 16 ? pb : (int64_t)kNumSMs * 16);
    bin_patches_kernel<false><<<pgrid, 256, 0, st>>>(coords, P, g, s.cnt, nullptr, nullptr);
    DH_CHECK_LAUNCH("bin_patches_kernel<count>");
    e = launch_dependent(pdl, bin_alloc_kernel, (unsigned)((ntiles + 255) / 256), 256, 0, st, s.cnt, ntiles, s.off, s.len);
    if (e != cudaSuccess) return cuda_fail(e, "bin_alloc_kernel");
    DH_CHECK_LAUNCH("bin_alloc_kernel");
    e = launch_dependent(pdl, bin_patches_kernel<true>, (unsigned)pgrid, 256, 0, st, coords, P, g, s.cnt, (const uint32_t*)s.off, s.list);
    if (e != cudaSuccess) return cuda_fail(e, "bin_patches_kernel<fill>");
    DH_CHECK_LAUNCH("bin_patches_kernel<fill>");
    const int64_t ctas = g.nty * ((g.ntx + kBinWarps - 1) / kBinWarps);
    DH_REQUIRE(ctas < (1ll << 31), "dh_stitch_binned: too many tiles");
    // at least 50 KB per CTA = at most 4 resident CTAs per SM: a fifth one only adds HBM write interleaving (measured 1.5-3 % slower)
    int smem = kBinWarps * bin_warp_smem_bytes(g.n, STAGED) + (g_bin_tile_rows / 100000) * 1024;
    if (!CELL && smem < 50 * 1024) smem = 50 * 1024;
    if constexpr (CELLN == 5 && !PHASED) {
        if (g_bin_variant == 4) {   // profiling only: the cell-lane kernel without its stores (the map is left untouched)
            auto kern = bin_cell_sum_kernel<5, false, true>;
            smem = kBinWarps * cell_warp_smem_bytes() + (g_bin_tile_rows / 100000) * 1024;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(bin_cell_sum_kernel)");
            e = launch_dependent(pdl, kern, (unsigned)ctas, kBinWarps * 32, (size_t)smem, st, logits, coords, g, s.off, s.len, s.list, s.sorted, sum_map, (uint32_t*)nullptr, (uint8_t*)nullptr);
        if (e != cudaSuccess) return cuda_fail(e, "tile kernel");
            DH_CHECK_LAUNCH("bin_cell_sum_kernel<nostore>");
            return DH_OK;
        }
    }
    if constexpr (CELLN > 0) {
        auto kern = bin_cell_sum_kernel<CELLN, PHASED>;
        smem = kBinWarps * cell_warp_smem_bytes();
        // Resident CTAs per SM through the shared-memory request. Tiled writes reach a higher share of the HBM write rate with FEWER
        // concurrent writers (an all-zero map written by this kernel: 0.80-0.88 of the copy peak at 4 CTAs per SM, 0.87-0.90 at 3,
        // 0.90-0.94 at 2), the staging and summing want more warps to overlap with. Measured optimum (profiles/r02_stitch.md):
        // 3 CTAs for footprints under 512 floats on aligned rows, 2 for wider ones, 4 on unaligned rows.
        if (!PHASED && sum_map != nullptr && g_bin_tile_rows / 100000 == 0) {
            const int per_sm = (int64_t)(g.ps / g.d) * g.n < 512 ? 3 : 2;
            const int want = (227 * 1024 / per_sm - 1024) / 1024 * 1024;            // the largest request that still fits per_sm CTAs
            const int cap = 227 * 1024 / (per_sm + 1) - 1024;                        // anything above this excludes per_sm + 1 CTAs
            if (smem <= cap) smem = cap + 1024 <= want ? cap + 1024 : want;
        }
        smem += (g_bin_tile_rows / 100000) * 1024;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(bin_cell_sum_kernel)");
        e = launch_dependent(pdl, kern, (unsigned)ctas, kBinWarps * 32, (size_t)smem, st, logits, coords, g, s.off, s.len, s.list, s.sorted, sum_map,
                             PHASED ? nullptr : count_map, PHASED ? nullptr : argmax_u8);
        if (e != cudaSuccess) return cuda_fail(e, "tile kernel");
    } else if constexpr (SEGK >= 0) {
        auto kern = bin_seg_kernel<SEGK, G>;
        smem = kBinWarps * seg_warp_smem_bytes();
        // wide footprints on unaligned rows: three resident CTAs per SM instead of four (fewer concurrent writers, see bin_cell_sum_kernel's
        // launch): 39 999^2, d = 2 / 1: 0.80 / 0.86 -> 0.83 / 0.90 of the HBM peak (two CTAs: 0.65 / 0.71); d = 4 wants all four (0.65 vs 0.58)
        if (SEGK == 1 && g_bin_tile_rows / 100000 == 0 && (int64_t)(g.ps / g.d) * g.n >= 512) {
            const int cap = 227 * 1024 / 4 - 1024;              // anything above this excludes a fourth CTA
            if (smem <= cap) smem = cap + 1024;
        }
        smem += (g_bin_tile_rows / 100000) * 1024;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(bin_seg_kernel)");
        e = launch_dependent(pdl, kern, (unsigned)ctas, kBinWarps * 32, (size_t)smem, st, logits, coords, g, s.off, s.len, s.list, s.sorted, sum_map, count_map, argmax_u8);
        if (e != cudaSuccess) return cuda_fail(e, "tile kernel");
    } else if constexpr (PHASED) {
        auto kern = bin_tile_phased_kernel<STAGED>;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(bin_tile_phased_kernel)");
        }
        e = launch_dependent(pdl, kern, (unsigned)ctas, kBinWarps * 32, (size_t)smem, st, logits, coords, g, s.off, s.len, s.list, s.sorted, sum_map);
        if (e != cudaSuccess) return cuda_fail(e, "tile kernel");
    } else {
        auto kern = bin_tile_kernel<VEC, G, CELL, STAGED>;
        if (smem > 48 * 1024) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(bin_tile_kernel)");
        }
        e = launch_dependent(pdl, kern, (unsigned)ctas, kBinWarps * 32, (size_t)smem, st, logits, coords, g, s.off, s.len, s.list, s.sorted, sum_map, count_map, argmax_u8);
        if (e != cudaSuccess) return cuda_fail(e, "tile kernel");
    }
    DH_CHECK_LAUNCH("bin_tile_kernel");
    return DH_OK;
}

}  // namespace dh

using namespace dh;

extern "C" DH_API int dh_stitch_binned_set_tile_rows(int rows) {
    g_bin_tile_rows = rows > 0 ? rows : 0;
    return DH_OK;
}

extern "C" DH_API int dh_stitch_binned_set_variant(int variant) {
    if (variant < 0 || variant > 4) { set_error("dh_stitch_binned_set_variant: variant must be 0 (auto), 1 (row-run kernels), 2 (segment kernel), 3 (cell-lane kernel) or 4 (profiling: cell-lane kernel without stores)"); return DH_ERR_INVALID; }
    g_bin_variant = variant;
    return DH_OK;
}

extern "C" DH_API int64_t dh_stitch_binned_scratch_bytes(int64_t P, int ps, int d, int n, int64_t rows, int64_t dw) {
    if (P <= 0 || ps <= 0 || d <= 0 || n <= 0 || rows <= 0 || dw <= 0) return 256;
    int64_t need = 0;
    for (int mode = 0; mode < 7; ++mode) {  // sum (16-byte stores), sum (scalar stores), cell, cell x 4, sum (phased 16-byte stores), cell-lane sum (aligned, phased)
        BinGeom g = mode >= 5 ? make_geom(false, 4, ps, d, n, rows, dw, 0, mode == 6, true)
                              : make_geom(mode == 2 || mode == 3, mode == 1 || mode == 2 ? 1 : 4, ps, d, n, rows, dw, 0, mode == 4);
        BinScratch s;
        carve_bin(g, P, nullptr, s);
        need = s.total_bytes > need ? s.total_bytes : need;
    }
    return need;
}

extern "C" DH_API int dh_stitch_binned(const float* logits, const int32_t* coords, int64_t P, int ps, int d, int n, float* sum_map,
                                       uint32_t* count_map, uint8_t* argmax_u8, int64_t rows, int64_t dw, int64_t row_offset, void* scratch,
                                       int64_t scratch_bytes, void* stream) {
    DH_REQUIRE(sum_map || count_map || argmax_u8, "dh_stitch_binned: no output requested");
    DH_REQUIRE(ps > 0 && d > 0 && n > 0 && rows >= 0 && dw >= 0 && P >= 0 && row_offset >= 0, "dh_stitch_binned: bad sizes");
    DH_REQUIRE(P < (1ll << 32) && rows < (1ll << 31) - 4096 && dw * (int64_t)n < (1ll << 31) - 4096, "dh_stitch_binned: sizes exceed 32-bit tile coordinates");
    DH_REQUIRE(!(argmax_u8 && n > 256), "dh_stitch_binned: at most 256 classes fit the u8 class map");
    DH_REQUIRE(!(argmax_u8 && n > kBinMaxN && !sum_map), "dh_stitch_binned: the class map of more than %d classes needs the sum map", kBinMaxN);
    if (rows == 0 || dw == 0) return DH_OK;
    cudaStream_t st = as_stream(stream);
    if (P == 0) {  // no patches: the reference's map stays zero
        cudaError_t e = cudaSuccess;
        if (sum_map) e = cudaMemsetAsync(sum_map, 0, (size_t)rows * dw * n * 4, st);
        if (e == cudaSuccess && count_map) e = cudaMemsetAsync(count_map, 0, (size_t)rows * dw * 4, st);
        if (e == cudaSuccess && argmax_u8) e = cudaMemsetAsync(argmax_u8, 0, (size_t)rows * dw, st);
        return e == cudaSuccess ? DH_OK : cuda_fail(e, "cudaMemsetAsync");
    }
    DH_REQUIRE(logits && coords && scratch, "dh_stitch_binned: null input");
    if (const char* e = getenv("DH_BIN_SPARSE")) g_bin_sparse = e[0] == '0' ? 0 : (e[0] == '1' ? 1 : -1);   // profiling override
    int rc = DH_OK;
    const bool staged = n <= kBinMaxN;
    uint8_t* cell_argmax = argmax_u8 && staged ? argmax_u8 : nullptr;
    bool cells_done = false;     // count / class map already written by the cell-lane launch
#define DH_CELL_CASE(NN, PH, SUM, CNT, AMX)                                                                                                   \
    case NN:                                                                                                                                  \
        rc = PH ? run_binned<4, 1, false, true, true, -1, NN>(logits, coords, P, gc, SUM, nullptr, nullptr, scratch, scratch_bytes, st)        \
                : run_binned<4, 1, false, true, false, -1, NN>(logits, coords, P, gc, SUM, CNT, AMX, scratch, scratch_bytes, st);              \
        break;
#define DH_CELL_SWITCH(PH, SUM, CNT, AMX)                                                                                                     \
    switch (n) {                                                                                                                              \
        DH_CELL_CASE(1, PH, SUM, CNT, AMX) DH_CELL_CASE(2, PH, SUM, CNT, AMX) DH_CELL_CASE(3, PH, SUM, CNT, AMX) DH_CELL_CASE(4, PH, SUM, CNT, AMX) \
        DH_CELL_CASE(5, PH, SUM, CNT, AMX) DH_CELL_CASE(6, PH, SUM, CNT, AMX) DH_CELL_CASE(7, PH, SUM, CNT, AMX) DH_CELL_CASE(8, PH, SUM, CNT, AMX) \
        default: rc = DH_ERR_INVALID; break;                                                                                                  \
    }
    if (sum_map) {
        const bool v4 = sum_vec4(sum_map, dw, n);
        // rows not 16-byte aligned (dw * n % 4 != 0) but an aligned base: 16-byte stores at a per-row shift (bin_tile_phased_kernel)
        // (not for small footprints: 7 instead of 4 sums per lane cost more than the vector stores save -- measured at d = 16)
        const bool shiftable = !v4 && reinterpret_cast<uintptr_t>(sum_map) % 16 == 0 && dw * (int64_t)n >= 8;
        const bool phased = shiftable && ps / d >= 24;
        const BinGeom g = make_geom(false, v4 || phased ? 4 : 1, ps, d, n, rows, dw, row_offset, phased);
        // measured (profiles/r02_stitch.md): the segment kernel wins where the row-run kernel has to sum 7 floats per lane (rows not
        // 16-byte aligned: 0.61 vs 0.51 of the HBM peak at d = 4), the row-run kernel wins on aligned rows (0.71 vs 0.68)
        // the cell-lane kernel (one cell per lane, issue-active 46 % where the row-run kernel has 72 %), measured on the 40 000^2 coverage
        // list with streaming stores in every kernel: aligned rows d = 16 / 4 / 2 / 1: 0.14 / 0.83 / 0.97 / 1.03 of the HBM copy peak vs
        // 0.10 / 0.77 / 0.93 / 1.02 with the row-run kernel. Unaligned rows: small footprints only (d = 16: 0.13 vs 0.09); the segment
        // kernel keeps the wide ones (d = 4: 0.68 vs 0.58).
        const bool cell_phased = shiftable;
        const bool cell_lane = staged && tile_rows_for(ps, d) <= 128 &&
                               ((g_bin_variant >= 3 && (v4 || cell_phased)) || (g_bin_variant == 0 && ((v4 && (int64_t)(ps / d) * n < 2048) || (shiftable && ps / d < 24))));
        if (cell_lane) {
            const BinGeom gc = make_geom(false, 4, ps, d, n, rows, dw, row_offset, !v4, true);
            // aligned rows: the count and the class map come out of the same launch (a lane owns a cell) instead of a second binning + tile pass
            // (footprints under 1024 floats: at d = 1 the per-lane byte / word stores of a 0.4 + 1.6 GB map pair cost as much as the
            // segment kernel's separate pass with 4 cells per lane: 6.06 ms fused vs ~5.5 ms, profiles/r02_stitch.md)
            const bool fuse = v4 && (count_map || cell_argmax) && (int64_t)(ps / d) * n < 1024;
            uint32_t* fc = fuse ? count_map : nullptr;
            uint8_t* fa = fuse ? cell_argmax : nullptr;
            DH_CELL_SWITCH(!v4, sum_map, fc, fa)
            if (rc != DH_OK) return rc;
            cells_done = fuse;
        }
        const bool seg = !cell_lane && staged && (g_bin_variant == 2 || (g_bin_variant == 0 && phased)) && g.TH <= 128 && (int64_t)g.TW / n + 4 <= kSegCap - 2 && g.TW + 3 + 3 * n <= kSegVals;
        if (cell_lane) {
            // done above
        } else if (seg && phased) rc = run_binned<4, 1, false, true, true, 1>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st);
        else if (seg && v4 && g.G == 2) rc = run_binned<4, 2, false, true, false, 0>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st);
        else if (seg && v4) rc = run_binned<4, 1, false, true, false, 0>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st);
        else if (seg) rc = run_binned<1, 1, false, true, false, 2>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st);
        else if (phased) rc = staged ? run_binned<4, 1, false, true, true>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st)
                                : run_binned<4, 1, false, false, true>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st);
        else if (v4 && g.G == 2) rc = staged ? run_binned<4, 2, false, true>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st)
                                             : run_binned<4, 2, false, false>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st);
        else if (v4) rc = staged ? run_binned<4, 1, false, true>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st)
                                 : run_binned<4, 1, false, false>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st);
        else rc = staged ? run_binned<1, 1, false, true>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st)
                         : run_binned<1, 1, false, false>(logits, coords, P, g, sum_map, nullptr, nullptr, scratch, scratch_bytes, st);
        if (rc != DH_OK) return rc;
    }
    int64_t cellout_max = 512;                                      // floats; DH_BIN_CELLOUT_MAX: profiling override
    if (const char* e = getenv("DH_BIN_CELLOUT_MAX")) cellout_max = atoll(e);
    if (!sum_map && cell_argmax && (g_bin_variant == 0 || g_bin_variant == 3) && tile_rows_for(ps, d) <= 128 && (int64_t)(ps / d) * n < cellout_max) {
        // class map (and count) without a sum map, footprints under 512 floats: the cell-lane kernel with its sums left in registers
        // (40 000^2 list, class map only: d = 16 / 4 see profiles/r02_stitch.md); wider footprints keep the segment kernel below
        const BinGeom gc = make_geom(false, 4, ps, d, n, rows, dw, row_offset, false, true);
        float* const no_sum = nullptr;
        DH_CELL_SWITCH(false, no_sum, count_map, cell_argmax)
        if (rc != DH_OK) return rc;
        cells_done = true;
    }
#undef DH_CELL_SWITCH
#undef DH_CELL_CASE
    if (!cells_done && (count_map || cell_argmax)) {
        // 4 cells per lane (uchar4 / uint4 stores) when the rows keep the vectors aligned
        // -- for footprints at least ~100 cells wide; narrower ones make a 128-cell tile mostly foreign patches (measured: d = 1 / 2 are
        // 1.9x / 1.3x faster with 4 cells per lane, d = 4 / 16 1.3x / 1.7x slower; profiles/r01_stitch.md)
        const bool c4 = ps / d >= 96 && dw % 4 == 0 && reinterpret_cast<uintptr_t>(cell_argmax) % 4 == 0 &&
                        reinterpret_cast<uintptr_t>(count_map) % 16 == 0;
        const BinGeom g = make_geom(true, c4 ? 4 : 1, ps, d, cell_argmax ? n : 1, rows, dw, row_offset);   // count only: no logits staged
        // class map / count outputs: the segment kernel for footprints of 24 cells and more (40 000^2 class map, d = 4 / 2 / 1: 0.33 / 0.43 /
        // 0.67 ms vs 0.43 / 0.69 / 1.13 ms with the row-run kernel; d = 16: 0.25 vs 0.20 ms, the row-run kernel stays)
        if ((g_bin_variant == 2 || (g_bin_variant == 0 && ps / d >= 24)) && g.TH <= 128)
            rc = c4 ? run_binned<4, 1, true, true, false, 4>(logits, coords, P, g, nullptr, count_map, cell_argmax, scratch, scratch_bytes, st)
                    : run_binned<1, 1, true, true, false, 3>(logits, coords, P, g, nullptr, count_map, cell_argmax, scratch, scratch_bytes, st);
        else
            rc = c4 ? run_binned<4, 1, true, true>(logits, coords, P, g, nullptr, count_map, cell_argmax, scratch, scratch_bytes, st)
                    : run_binned<1, 1, true, true>(logits, coords, P, g, nullptr, count_map, cell_argmax, scratch, scratch_bytes, st);
        if (rc != DH_OK) return rc;
    }
    if (argmax_u8 && !cell_argmax) rc = dh_stitch_finalize(sum_map, nullptr, rows * dw, n, nullptr, argmax_u8, stream);
    return rc;
}
