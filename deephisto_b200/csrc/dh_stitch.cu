// K5: overlap-aware stitching of per-patch logits into the downscaled whole-slide map.
//
// Reference path replaced:
//   examples/predict_full_patched.py:40-63  ImagePredictorPatched.process
//     prediction[y//d:(y+ps)//d, x//d:(x+ps)//d, :] += logits_i   (sampler order)   then argmax(axis=2)
//
// dh_stitch_dense is a GATHER formulation for the dense enumeration of full_samplers.py:374-404:
// every output cell sums the logits of the patches that cover it, visited in the reference's order
// (main grid row-major, last column, last row, corner, padding copies of the corner), with plain
// fp32 adds, so the sum map is bit-identical to the numpy loop and needs no atomics.
// Rows that are covered by the same set of patch rows have identical values: a block computes one
// column tile once per such row class into shared memory and streams it out for every row of the
// class, so the kernel is a pure coalesced store stream (16-byte vectors) -> HBM write roofline.
#include "dh_common.cuh"
#include <cstdlib>

namespace dh {

struct StitchGrid {
    int64_t H, W;
    int ps, stride, d, n;
    int64_t ny, nx, N, pads;  // pads = number of padding copies of the corner patch
    int64_t dh, dw;
    int64_t lastrow_cell, lastcol_cell;  // (H-ps)//d, (W-ps)//d
};

struct Cover {
    int64_t lo, hi;  // main-grid index range (inclusive), empty if lo > hi
    bool last;       // covered by the last row / column patch
};

// patches covering cell i of one axis; 32-bit arithmetic ((i + 1) * d <= H + d < 2^31, host check): a 64-bit division is ~100 instructions
__device__ __forceinline__ Cover cover_1d32(int i, int cnt, int ps, int stride, int d, int last_cell) {
    Cover c;
    const int e = (i + 1) * d;
    int hi = (e - 1) / stride;
    if (hi > cnt - 1) hi = cnt - 1;
    const int t = e - ps;
    c.hi = hi;
    c.lo = t <= 0 ? 0 : (t + stride - 1) / stride;
    c.last = i >= last_cell;
    return c;
}

constexpr int kStitchThreads = 256;
static int g_stitch_stage = 1;   // 0: profiling A/B, logits read from HBM per row class (dh_stitch_dense_set_variant)
#ifndef DH_STITCH_MINB
#define DH_STITCH_MINB 3   // resident CTAs per SM the dense kernel is compiled for (80 registers, no spills; 4 = 64 registers, 120 B of spills)
#endif

// ---- register-resident row classes ----------------------------------------------------------------------
// When dw*n is a multiple of 4 floats and the column tile is a multiple of 4 cells, every row segment of a tile starts on
// a 16-byte boundary: a thread keeps its float4s (and count / argmax words) of the current row class IN REGISTERS and the
// per-row work is just the stores (other widths: PHASED below). The row class (which patch rows cover map row i) is advanced incrementally -- no division
// in the row loop (the first version of this kernel spent its time on two 64-bit divisions per row per thread:
// profiles/r01_stitch.md).
constexpr int kStitchV = 2;  // float4 per thread per row
constexpr int kStitchNC = 8; // classes summed in registers by the pipelined per-cell loop (more classes: one class at a time)

struct RowClass {            // 32-bit: (i + 1) * d <= H + d < 2^31 (host check)
    int lo_raw, hi_raw;      // unclamped main-grid row range covering the current map row
    int rem_lo, rem_hi;
    __device__ __forceinline__ void init(int64_t i, int ps, int stride, int d) {
        const int e = (int)(i + 1) * d;
        hi_raw = (e - 1) / stride;
        rem_hi = (e - 1) - hi_raw * stride;
        const int u = e - ps + stride - 1;  // lo = max(0, floor(u / stride))
        if (u >= 0) { lo_raw = u / stride; rem_lo = u - lo_raw * stride; }
        else        { lo_raw = 0; rem_lo = u; }
    }
    // number of consecutive map rows, starting at the current one, with the same (lo_raw, hi_raw)
    __device__ __forceinline__ int run_length(int stride, int d) const {
        const int a = (stride - rem_hi + d - 1) / d;  // rows until hi_raw changes (rem_hi < stride)
        const int b = (stride - rem_lo + d - 1) / d;  // rows until lo_raw changes (rem_lo may be negative)
        return a < b ? a : b;
    }
    __device__ __forceinline__ void advance(int rows, int stride, int d) {
        rem_hi += rows * d;
        if (rem_hi >= stride) { const int q = rem_hi / stride; hi_raw += q; rem_hi -= q * stride; }
        rem_lo += rows * d;
        if (rem_lo >= stride) { const int q = rem_lo / stride; lo_raw += q; rem_lo -= q * stride; }
    }
};

// PHASED: dw * n is not a multiple of 4, so the 16-byte phase of a row segment changes from row to row (p = row * dw * n mod 4).
// A thread then keeps one register set per phase -- the same class values cut into vectors at the four possible offsets -- and the
// row loop picks the set of the row's phase; the <= 3 floats before the first and after the last aligned vector of a row are
// stored as scalars from shared memory.
// Register budget: 4 resident CTAs per SM (64 registers) unless PHASED keeps four register sets. ncu at d = 16 (profiles/
// r02_dense_d16_ncu.txt): 101 registers -> 2 CTAs per SM, 24 % warps active, the short row classes (7 rows of stores per L2 round trip
// for the logits) left the kernel latency-bound at 0.29 of the HBM peak.
template <bool WITH_SUM, bool WITH_ARGMAX, bool WITH_COUNT, bool PHASED>
__global__ void __launch_bounds__(kStitchThreads, PHASED ? 2 : DH_STITCH_MINB) stitch_dense_aligned_kernel(const float* __restrict__ logits, StitchGrid g,
                                                                              float* __restrict__ sum_map,
                                                                              uint32_t* __restrict__ count_map,
                                                                              uint8_t* __restrict__ argmax_map,
                                                                              int64_t row_begin, int64_t row_end, int tj_max,
                                                                              int rows_per_block, int amax_vec, int cls_off, int stage_off) {
    extern __shared__ __align__(16) float smem[];
    const int n = g.n;
    float* vals = smem;                                                       // [tj_max * n]
    uint32_t* cnts = reinterpret_cast<uint32_t*>(smem + tj_max * n);          // [tj_max]
    uint8_t* amax = reinterpret_cast<uint8_t*>(cnts + tj_max);                // [tj_max] (tj_max % 16 == 0)
    // column classes (below): per-class sums / count / arg max, the class covers, the scan's per-warp totals
    const int cw = n > 2 ? n : 2;
    float* cvals = smem + cls_off;                                            // [tj_max * max(n, 2)] (the cell covers alias it during set-up)
    int2* ccov = reinterpret_cast<int2*>(cvals + tj_max * cw);                // [tj_max]
    uint32_t* ccnt = reinterpret_cast<uint32_t*>(ccov + tj_max);              // [tj_max]
    uint32_t* wsum = ccnt + tj_max;                                           // [16]
    uint8_t* camax = reinterpret_cast<uint8_t*>(wsum + 16);                   // [tj_max]

    const int64_t j0 = (int64_t)blockIdx.x * tj_max;
    const int tj = (int)((g.dw - j0) < tj_max ? (g.dw - j0) : tj_max);
    const int64_t i0 = row_begin + (int64_t)blockIdx.y * rows_per_block;
    const int64_t i1 = (i0 + rows_per_block) < row_end ? (i0 + rows_per_block) : row_end;
    const int tile_floats = tj * n;
    const int nvec = tile_floats >> 2;  // !PHASED: (tj * n) % 4 == 0 by construction
    const int tid = threadIdx.x;

    RowClass rc;
    rc.init(i0, g.ps, g.stride, g.d);
    // the column cover of a thread's cells (<= 2: tj <= 2 * kStitchThreads) does not depend on the row class: computed once
    // ---- stage the logits this block can touch (stage_off > 0: they fit the shared-memory budget, host check). The rows [i0, i1) are
    // covered by the main-grid patch rows [gy0, gy1], the tile's cells by the patch columns [gx0, gx1] (cover_1d32 is monotone), plus the
    // last-column / last-row / corner patches: a handful of CONTIGUOUS pieces of the logits array, fetched with one round of
    // asynchronous 4-byte copies. Every row class of the block then sums from shared memory -- without this a class of 7 rows
    // (d = 16) cost its own chain of L2 round trips (profiles/r02_stitch.md).
    const bool staged = stage_off > 0 && (WITH_SUM || WITH_ARGMAX);
    int gy0 = 0, gx0 = 0, nc = 0;                       // patch grid indices fit 32 bits (host check)
    int o_lc = 0, o_lr = 0, o_cn = 0;                   // float offsets of the last-column / last-row / corner pieces behind the main piece
    const float* const s_main = smem + stage_off;
    if (staged) {
        float* base = smem + stage_off;
        const Cover ra = cover_1d32((int)i0, (int)g.ny, g.ps, g.stride, g.d, (int)g.lastrow_cell), rb = cover_1d32((int)i1 - 1, (int)g.ny, g.ps, g.stride, g.d, (int)g.lastrow_cell);
        const Cover ca = cover_1d32((int)j0, (int)g.nx, g.ps, g.stride, g.d, (int)g.lastcol_cell), cb = cover_1d32((int)j0 + tj - 1, (int)g.nx, g.ps, g.stride, g.d, (int)g.lastcol_cell);
        gy0 = (int)ra.lo; gx0 = (int)ca.lo;
        const int nr = rb.hi >= ra.lo ? (int)(rb.hi - ra.lo + 1) : 0;
        nc = cb.hi >= ca.lo ? (int)(cb.hi - ca.lo + 1) : 0;
        const int64_t main_n = g.ny * g.nx;
        float* d_main = base;
        float* d_lc = d_main + (int64_t)nr * nc * n;
        float* d_lr = d_lc + (cb.last ? nr * n : 0);
        float* d_cn = d_lr + (rb.last ? nc * n : 0);
        o_lc = (int)(d_lc - d_main); o_lr = (int)(d_lr - d_main); o_cn = (int)(d_cn - d_main);
        auto fetch = [&](float* dst, const float* src, int len) {
            for (int k = tid; k < len; k += kStitchThreads) {
                const uint32_t a = (uint32_t)__cvta_generic_to_shared(dst + k);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(a), "l"(src + k) : "memory");
            }
        };
        for (int r = 0; r < nr; ++r) fetch(d_main + (int64_t)r * nc * n, logits + ((int64_t)(gy0 + r) * g.nx + gx0) * n, nc * n);
        if (cb.last) fetch(d_lc, logits + (main_n + gy0) * n, nr * n);
        if (rb.last) fetch(d_lr, logits + (main_n + g.ny + gx0) * n, nc * n);
        if (cb.last && rb.last) fetch(d_cn, logits + (g.N - 1) * n, (int)(g.pads + 1) * n);
    }
    // COLUMN CLASSES: neighbouring cells are covered by the same patch columns (stride / d cells in a row, 7 at d = 16), so their
    // values are equal in every row. The tile's cells are cut into classes once per block (cover per cell -> class starts -> block scan);
    // per row class one thread per COLUMN class sums the covering patches, every thread then copies the values of its own cells
    // (ncu at d = 16 before this: 60 % issue-active, 21 % of the instructions in the per-cell logit loads, profiles/r02_stitch.md).
    int my_cls[2] = {0, 0};
    int ncls = 0;
    {
        int2* cov = reinterpret_cast<int2*>(cvals);
        int2 mine[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int t = tid + k * kStitchThreads;
            const Cover c = cover_1d32((int)j0 + (t < tj ? t : 0), (int)g.nx, g.ps, g.stride, g.d, (int)g.lastcol_cell);
            mine[k] = make_int2((int)c.lo, (int)c.hi * 2 + (c.last ? 1 : 0));      // patch grid indices fit 31 bits (host check)
            if (t < tj) cov[t] = mine[k];
        }
        __syncthreads();
        const int lane = tid & 31, warp = tid >> 5;
        int rank[2];
        bool start[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int t = tid + k * kStitchThreads;
            bool st = false;
            if (t < tj) {
                st = t == 0;
                if (t > 0) { const int2 prev = cov[t - 1]; st = prev.x != mine[k].x || prev.y != mine[k].y; }
            }
            start[k] = st;
            const unsigned b = __ballot_sync(0xffffffffu, st);
            rank[k] = __popc(b & (0xffffffffu >> (31 - lane)));                   // class starts up to and including this cell, within the warp
            if (lane == 0) wsum[k * 8 + warp] = __popc(b);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            int before = 0;
            for (int q = 0; q < 16; ++q) {
                const int v = (int)wsum[q];
                if (q < k * 8 + warp) before += v;
                if (k == 0) ncls += v;
            }
            my_cls[k] = before + rank[k] - 1;
            if (start[k]) ccov[my_cls[k]] = mine[k];
        }
        // the barrier that opens the first row class publishes ccov and retires the reads of cov (which aliases cvals)
    }
    if (staged) asm volatile("cp.async.wait_all;" ::: "memory");   // issued before the class set-up; the same barrier publishes the copies
    float4 vreg[PHASED ? 4 : 1][kStitchV];
    uint32_t creg[2] = {0, 0};
    uint32_t areg[2] = {0, 0};
    float* out_row = WITH_SUM ? sum_map + ((i0 - row_begin) * g.dw + j0) * n : nullptr;
    int64_t cell_row = (i0 - row_begin) * g.dw + j0;
    const int64_t row_floats = g.dw * n;
    int64_t i = i0;
    while (i < i1) {
        // rows [i, i + run) share one class: same covering patch rows, same side of the last-row boundary
        Cover cy;
        cy.lo = rc.lo_raw;
        cy.hi = rc.hi_raw > g.ny - 1 ? g.ny - 1 : rc.hi_raw;
        cy.last = i >= g.lastrow_cell;
        int64_t run = rc.run_length(g.stride, g.d);
        if (!cy.last && i + run > g.lastrow_cell) run = g.lastrow_cell - i;
        if (i + run > i1) run = i1 - i;
        {
            __syncthreads();  // previous class has been read into registers by everybody
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
                const int t = tid + kc * kStitchThreads;       // column class
                if (t >= ncls) break;
                Cover cx;
                { const int2 cc = ccov[t]; cx.lo = cc.x; cx.hi = cc.y >> 1; cx.last = (cc.y & 1) != 0; }
                const int64_t main_n = g.ny * g.nx;
                auto p_main = [&](int64_t gy, int64_t gx) { return staged ? s_main + (((int)gy - gy0) * nc + ((int)gx - gx0)) * n : logits + (gy * g.nx + gx) * n; };
                auto p_lastcol = [&](int64_t gy) { return staged ? s_main + o_lc + ((int)gy - gy0) * n : logits + (main_n + gy) * n; };
                auto p_lastrow = [&](int64_t gx) { return staged ? s_main + o_lr + ((int)gx - gx0) * n : logits + (main_n + g.ny + gx) * n; };
                auto p_corner = [&](int64_t k) { return staged ? s_main + o_cn + (int)k * n : logits + (g.N - 1 + k) * n; };
                float best = 0.f;
                int best_c = 0;
                if (n <= kStitchNC) {
                    // Patches outer, classes inner: per class the adds keep the reference's order, but the n loads of a patch are
                    // independent and the loads of patch p+1 are issued before the adds of patch p -- the covering patches cost about
                    // one L2 round trip instead of one per (patch, class) (the d = 16 map is latency-bound: profiles/r01_stitch.md).
                    float acc[kStitchNC], cur[kStitchNC];
#pragma unroll
                    for (int c = 0; c < kStitchNC; ++c) { acc[c] = 0.f; cur[c] = 0.f; }
                    bool have = false;
                    auto visit = [&](const float* lg) {      // lg: the patch's logits, in shared memory (staged) or in HBM
                        float nxt[kStitchNC];
#pragma unroll
                        for (int c = 0; c < kStitchNC; ++c) nxt[c] = c < n ? lg[c] : 0.f;
                        if (have) {
#pragma unroll
                            for (int c = 0; c < kStitchNC; ++c) acc[c] = __fadd_rn(acc[c], cur[c]);
                        }
#pragma unroll
                        for (int c = 0; c < kStitchNC; ++c) cur[c] = nxt[c];
                        have = true;
                    };
                    for (int64_t gy = cy.lo; gy <= cy.hi; ++gy)
                        for (int64_t gx = cx.lo; gx <= cx.hi; ++gx) visit(p_main(gy, gx));
                    if (cx.last)
                        for (int64_t gy = cy.lo; gy <= cy.hi; ++gy) visit(p_lastcol(gy));
                    if (cy.last)
                        for (int64_t gx = cx.lo; gx <= cx.hi; ++gx) visit(p_lastrow(gx));
                    if (cx.last && cy.last)
                        for (int64_t k = 0; k <= g.pads; ++k) visit(p_corner(k));
                    if (have) {
#pragma unroll
                        for (int c = 0; c < kStitchNC; ++c) acc[c] = __fadd_rn(acc[c], cur[c]);
                    }
#pragma unroll
                    for (int c = 0; c < kStitchNC; ++c) {
                        if (c < n) {
                            if (WITH_SUM) cvals[t * n + c] = acc[c];
                            if (WITH_ARGMAX) {
                                if (c == 0 || acc[c] > best || (acc[c] != acc[c] && best == best)) { best = acc[c]; best_c = c; }  // np.argmax: first maximum; NaN wins
                            }
                        }
                    }
                } else {
                for (int c = 0; c < n; ++c) {
                    float acc = 0.f;
                    for (int64_t gy = cy.lo; gy <= cy.hi; ++gy)
                        for (int64_t gx = cx.lo; gx <= cx.hi; ++gx) acc = __fadd_rn(acc, p_main(gy, gx)[c]);
                    if (cx.last)
                        for (int64_t gy = cy.lo; gy <= cy.hi; ++gy) acc = __fadd_rn(acc, p_lastcol(gy)[c]);
                    if (cy.last)
                        for (int64_t gx = cx.lo; gx <= cx.hi; ++gx) acc = __fadd_rn(acc, p_lastrow(gx)[c]);
                    if (cx.last && cy.last)
                        for (int64_t k = 0; k <= g.pads; ++k) acc = __fadd_rn(acc, p_corner(k)[c]);
                    if (WITH_SUM) cvals[t * n + c] = acc;
                    if (WITH_ARGMAX) {
                        if (c == 0 || acc > best || (acc != acc && best == best)) { best = acc; best_c = c; }  // np.argmax: first maximum; NaN wins
                    }
                }
                }
                if (WITH_COUNT) {
                    int64_t ry = cy.hi >= cy.lo ? cy.hi - cy.lo + 1 : 0;
                    int64_t rx = cx.hi >= cx.lo ? cx.hi - cx.lo + 1 : 0;
                    ccnt[t] = (uint32_t)(ry * rx + (cx.last ? ry : 0) + (cy.last ? rx : 0) + ((cx.last && cy.last) ? 1 + g.pads : 0));
                }
                if (WITH_ARGMAX) camax[t] = (uint8_t)best_c;
            }
            __syncthreads();
            // every thread copies the class values of its own cells
#pragma unroll
            for (int kc = 0; kc < 2; ++kc) {
                const int t = tid + kc * kStitchThreads;
                if (t >= tj) break;
                const int k = my_cls[kc];
                if (WITH_SUM)
                    for (int c = 0; c < n; ++c) vals[t * n + c] = cvals[k * n + c];
                if (WITH_COUNT) cnts[t] = ccnt[k];
                if (WITH_ARGMAX) amax[t] = camax[k];
            }
            __syncthreads();
            if (WITH_SUM) {
                if constexpr (PHASED) {
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int head = (4 - p) & 3;
#pragma unroll
                        for (int v = 0; v < kStitchV; ++v) {
                            const int b = head + 4 * (tid + v * kStitchThreads);
                            if (b + 3 < tile_floats) vreg[p][v] = make_float4(vals[b], vals[b + 1], vals[b + 2], vals[b + 3]);
                        }
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < kStitchV; ++v) {
                        const int q = tid + v * kStitchThreads;
                        if (q < nvec) vreg[0][v] = reinterpret_cast<const float4*>(vals)[q];
                    }
                }
            }
            if (WITH_COUNT) {
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const int t = tid + v * kStitchThreads;
                    if (t < tj) creg[v] = cnts[t];
                }
            }
            if (WITH_ARGMAX) {
                if (amax_vec) {  // 4 cells per 32-bit store (dw % 4 == 0, so j0 and the row starts are 4-byte aligned)
                    if (tid < ((tj + 3) >> 2)) areg[0] = reinterpret_cast<const uint32_t*>(amax)[tid];
                } else {
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const int t = tid + v * kStitchThreads;
                        if (t < tj) areg[v] = amax[t];
                    }
                }
            }
        }
        // stream the class out: nothing but stores and pointer increments per row
        if (WITH_SUM) {
            if constexpr (PHASED) {
                float* rowp = out_row;
                int p = (int)(((i - row_begin) * row_floats + j0 * n) & 3);   // sum_map is 16-byte aligned (host check)
                const int pstep = (int)(row_floats & 3);
                for (int64_t r = 0; r < run; ++r) {
                    int head = (4 - p) & 3;
                    if (head > tile_floats) head = tile_floats;            // a last tile of one or two floats
                    const int nv = (tile_floats - head) >> 2;
                    const int tail0 = head + 4 * nv;
                    float4* o = reinterpret_cast<float4*>(rowp + head) + tid;
                    const bool h0 = tid < nv, h1 = tid + kStitchThreads < nv;
                    switch (p) {
                        case 0: if (h0) __stcs(o, vreg[0][0]); if (h1) __stcs(o + kStitchThreads, vreg[0][1]); break;
                        case 1: if (h0) __stcs(o, vreg[1][0]); if (h1) __stcs(o + kStitchThreads, vreg[1][1]); break;
                        case 2: if (h0) __stcs(o, vreg[2][0]); if (h1) __stcs(o + kStitchThreads, vreg[2][1]); break;
                        default: if (h0) __stcs(o, vreg[3][0]); if (h1) __stcs(o + kStitchThreads, vreg[3][1]); break;
                    }
                    if (tid < head) rowp[tid] = vals[tid];
                    else if (tid >= 32 && tid - 32 < tile_floats - tail0) rowp[tail0 + tid - 32] = vals[tail0 + tid - 32];
                    rowp += row_floats;
                    p = (p + pstep) & 3;
                }
            } else {
                float4* o0 = reinterpret_cast<float4*>(out_row) + tid;
                const bool h0 = tid < nvec, h1 = tid + kStitchThreads < nvec;
                for (int64_t r = 0; r < run; ++r) {
                    if (h0) __stcs(o0, vreg[0][0]);
                    if (h1) __stcs(o0 + kStitchThreads, vreg[0][1]);
                    o0 = reinterpret_cast<float4*>(reinterpret_cast<float*>(o0) + row_floats);
                }
            }
            out_row += run * row_floats;
        }
        if (WITH_COUNT) {
            uint32_t* c0 = count_map + cell_row + tid;
            const bool h0 = tid < tj, h1 = tid + kStitchThreads < tj;
            for (int64_t r = 0; r < run; ++r) {
                if (h0) __stcs(c0, creg[0]);
                if (h1) __stcs(c0 + kStitchThreads, creg[1]);
                c0 += g.dw;
            }
        }
        if (WITH_ARGMAX) {
            if (amax_vec) {
                uint8_t* a0 = argmax_map + cell_row;
                for (int64_t r = 0; r < run; ++r) {
                    if (tid < (tj >> 2)) __stcs(reinterpret_cast<uint32_t*>(a0) + tid, areg[0]);
                    else if (tid == (tj >> 2) && (tj & 3)) {
                        for (int b = 0; b < (tj & 3); ++b) a0[(tj & ~3) + b] = (uint8_t)(areg[0] >> (8 * b));
                    }
                    a0 += g.dw;
                }
            } else {
                uint8_t* a0 = argmax_map + cell_row + tid;
                const bool h0 = tid < tj, h1 = tid + kStitchThreads < tj;
                for (int64_t r = 0; r < run; ++r) {
                    if (h0) __stcs(a0, (uint8_t)areg[0]);
                    if (h1) __stcs(a0 + kStitchThreads, (uint8_t)areg[1]);
                    a0 += g.dw;
                }
            }
        }
        cell_row += run * g.dw;
        i += run;
        rc.advance((int)run, g.stride, g.d);
    }
}

// ---- scatter (arbitrary coordinates) ------------------------------------------------------------
// One block per patch, one warp per footprint row. A footprint row is a contiguous run of fw*n floats of the map; its 16-byte
// aligned body is added with vector reductions (red.global.add.v4.f32: 4x fewer atomic operations than one per float), the
// unaligned head and tail with scalar ones. Value of float f of the run = logit[(f) mod n].
__global__ void __launch_bounds__(256) stitch_scatter_kernel(const float* __restrict__ logits, const int32_t* __restrict__ coords,
                                                             int64_t P, int ps, int d, int n, float* __restrict__ sum_map,
                                                             uint32_t* __restrict__ count_map, int64_t rows, int64_t dw,
                                                             int64_t row_offset, int vec_ok) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int64_t patch = blockIdx.x; patch < P; patch += gridDim.x) {
        const int y = __ldg(coords + 2 * patch), x = __ldg(coords + 2 * patch + 1);
        // numpy slice semantics: start/stop clipped to [0, size]
        int64_t r0 = y / d, r1 = ((int64_t)y + ps) / d, c0 = x / d, c1 = ((int64_t)x + ps) / d;
        if (y < 0) r0 = 0;  // callers never pass negative origins; keep the footprint inside the map
        if (x < 0) c0 = 0;
        if (c1 > dw) c1 = dw;
        int64_t lo = r0 > row_offset ? r0 : row_offset;
        int64_t hi = r1 < row_offset + rows ? r1 : row_offset + rows;
        if (hi <= lo || c1 <= c0) continue;
        const int fw = (int)(c1 - c0);
        const int len = fw * n;
        const float* lg = logits + patch * n;
        for (int64_t rr = lo + warp; rr < hi; rr += nwarps) {
            const int64_t cell0 = (rr - row_offset) * dw + c0;
            if (sum_map) {
                float* run = sum_map + cell0 * n;
                int head = vec_ok ? (int)((4 - ((cell0 * n) & 3)) & 3) : len;   // floats before the first 16-byte boundary
                if (head > len) head = len;
                const int nvec = (len - head) >> 2;
                const int tail0 = head + 4 * nvec;
                if (lane < head) atomicAdd(run + lane, __ldg(lg + lane % n));
                for (int q = lane; q < nvec; q += 32) {
                    const int f = head + 4 * q;
                    int c = f % n;
                    float4 v;
                    v.x = __ldg(lg + c); c = c + 1 == n ? 0 : c + 1;
                    v.y = __ldg(lg + c); c = c + 1 == n ? 0 : c + 1;
                    v.z = __ldg(lg + c); c = c + 1 == n ? 0 : c + 1;
                    v.w = __ldg(lg + c);
                    atomicAdd(reinterpret_cast<float4*>(run + f), v);
                }
                if (vec_ok) {
                    if (lane < len - tail0) atomicAdd(run + tail0 + lane, __ldg(lg + (tail0 + lane) % n));
                } else {
                    for (int f = lane; f < len; f += 32) atomicAdd(run + f, __ldg(lg + f % n));
                }
            }
            if (count_map)
                for (int cc = lane; cc < fw; cc += 32) atomicAdd(count_map + cell0 + cc, 1u);
        }
    }
}

__global__ void __launch_bounds__(256) stitch_finalize_kernel(const float* __restrict__ sum_map, const uint32_t* __restrict__ count_map,
                                                              int64_t cells, int n, float* __restrict__ norm_map,
                                                              uint8_t* __restrict__ argmax_map) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
        float cntf = 1.f;
        if (count_map) { uint32_t c = count_map[i]; cntf = c ? (float)c : 1.f; }
        float best = 0.f;
        int best_c = 0;
        for (int c = 0; c < n; ++c) {
            float v = sum_map[i * n + c];
            if (norm_map) norm_map[i * n + c] = __fdiv_rn(v, cntf);
            if (c == 0 || v > best || (v != v && best == best)) { best = v; best_c = c; }
        }
        if (argmax_map) argmax_map[i] = (uint8_t)best_c;
    }
}

// ---- prediction post-processing ---------------------------------------------------------
// Reference: examples/predict_full_patched.py:81-113 perform_and_save_visualizations
//   colored_image[pred == anno.id] = anno.color                      -> class-colour LUT
//   img = psim.get_region((0,0), (H,W), target_hw=(h,w))             -> slide thumbnail [h][w][3]
//   (img * alpha + colored_image * (1 - alpha)).astype(np.uint8)     -> overlay, float64 arithmetic, truncation
// The thumbnail here is the exact integer AREA AVERAGE of the d x d slide block under every map cell, rounded half up:
// (sum + d*d/2) / (d*d). (psimage's own resampling filter is unknown -- the package is not in the reference tree -- so the
// thumbnail is this build's definition; the LUT and the blend follow the reference arithmetic exactly.)
// One thread per (cell, channel); a warp covers ~11 neighbouring cells, so the d x d x 3 byte block of the slide is read from
// DRAM once and re-touched through L1. Algorithmic bytes: dh*d * dw*d * 3 read + up to 3 * dh*dw*3 written.
__global__ void __launch_bounds__(256) colorize_overlay_kernel(const uint8_t* __restrict__ argmax_map, const uint8_t* __restrict__ slide,
                                                               int64_t pitch, int64_t dh, int64_t dw, int d,
                                                               const uint8_t* __restrict__ lut, double alpha,
                                                               uint8_t* __restrict__ mask_out, uint8_t* __restrict__ thumb_out,
                                                               uint8_t* __restrict__ overlay_out) {
    const int64_t row_elems = dw * 3;
    const int64_t total = dh * row_elems;
    const uint32_t area = (uint32_t)d * (uint32_t)d;
    const double beta = __dsub_rn(1.0, alpha);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / row_elems;
        const int64_t e = t - i * row_elems;   // 3*j + c
        const int64_t j = e / 3;
        const int c = (int)(e - 3 * j);
        const uint32_t col = lut[3 * (uint32_t)argmax_map[i * dw + j] + c];
        if (mask_out) mask_out[t] = (uint8_t)col;
        if (thumb_out || overlay_out) {
            const uint8_t* src = slide + (i * d) * pitch + 3 * (j * d) + c;
            uint32_t sum = 0;
            for (int r = 0; r < d; ++r) {
                const uint8_t* row = src + (int64_t)r * pitch;
#pragma unroll 4
                for (int px = 0; px < d; ++px) sum += __ldg(row + 3 * px);
            }
            const uint32_t img = (sum + area / 2) / area;
            if (thumb_out) thumb_out[t] = (uint8_t)img;
            if (overlay_out) overlay_out[t] = (uint8_t)(int)__dadd_rn(__dmul_rn((double)img, alpha), __dmul_rn((double)col, beta));
        }
    }
}

// Vector variant for d % 4 == 0 (VW = 1: 32-bit loads) and d % 16 == 0 (VW = 4: 128-bit loads): one thread per CELL reads its
// 3*d-byte row segments as words; a group of 3 words is 4 RGB pixels with the channel phases (0,1,2,0)(1,2,0,1)(2,0,1,2), so the
// per-channel sums are nine DP4A with constant byte masks.
__device__ __forceinline__ void rgb_sums3(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t& r, uint32_t& g, uint32_t& b) {
    r = __dp4a(w0, 0x01000001u, r); g = __dp4a(w0, 0x00000100u, g); b = __dp4a(w0, 0x00010000u, b);
    r = __dp4a(w1, 0x00010000u, r); g = __dp4a(w1, 0x01000001u, g); b = __dp4a(w1, 0x00000100u, b);
    r = __dp4a(w2, 0x00000100u, r); g = __dp4a(w2, 0x00010000u, g); b = __dp4a(w2, 0x01000001u, b);
}

template <int VW>
__global__ void __launch_bounds__(256) colorize_overlay_vec_kernel(const uint8_t* __restrict__ argmax_map, const uint8_t* __restrict__ slide,
                                                                   int64_t pitch, int64_t dh, int64_t dw, int d,
                                                                   const uint8_t* __restrict__ lut, double alpha,
                                                                   uint8_t* __restrict__ mask_out, uint8_t* __restrict__ thumb_out,
                                                                   uint8_t* __restrict__ overlay_out) {
    const int64_t total = dh * dw;
    const uint32_t area = (uint32_t)d * (uint32_t)d;
    const double beta = __dsub_rn(1.0, alpha);
    const int groups = d / 4;  // groups of 3 words (12 bytes = 4 pixels) per row segment
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / dw, j = t - i * dw;
        const uint8_t* src = slide + (i * d) * pitch + 3 * (j * d);
        uint32_t sr = 0, sg = 0, sb = 0;
        for (int r = 0; r < d; ++r) {
            const uint8_t* row = src + (int64_t)r * pitch;
            if (VW == 4) {
                const uint4* q = reinterpret_cast<const uint4*>(row);
                for (int k = 0; k < groups / 4; ++k) {  // 3 x 16 bytes = 4 groups
                    const uint4 a = __ldg(q + 3 * k), b = __ldg(q + 3 * k + 1), c = __ldg(q + 3 * k + 2);
                    rgb_sums3(a.x, a.y, a.z, sr, sg, sb);
                    rgb_sums3(a.w, b.x, b.y, sr, sg, sb);
                    rgb_sums3(b.z, b.w, c.x, sr, sg, sb);
                    rgb_sums3(c.y, c.z, c.w, sr, sg, sb);
                }
            } else {
                const uint32_t* q = reinterpret_cast<const uint32_t*>(row);
                for (int k = 0; k < groups; ++k) rgb_sums3(__ldg(q + 3 * k), __ldg(q + 3 * k + 1), __ldg(q + 3 * k + 2), sr, sg, sb);
            }
        }
        const uint32_t cls = argmax_map[t];
        const uint32_t sum[3] = {sr, sg, sb};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint32_t col = lut[3 * cls + c];
            const uint32_t img = (sum[c] + area / 2) / area;
            if (mask_out) mask_out[3 * t + c] = (uint8_t)col;
            if (thumb_out) thumb_out[3 * t + c] = (uint8_t)img;
            if (overlay_out) overlay_out[3 * t + c] = (uint8_t)(int)__dadd_rn(__dmul_rn((double)img, alpha), __dmul_rn((double)col, beta));
        }
    }
}

// d = 4 or 8: one thread per 16-PIXEL column block (48 bytes per slide row = three 128-bit loads = CP = 16 / d cells), so loads and
// stores are vectors instead of the 12-byte segments / single bytes of the per-cell kernel. The last block of a map row may hold
// fewer than CP cells: those are read with 32-bit loads like the per-cell kernel.
template <int CP>
__global__ void __launch_bounds__(256) colorize_overlay_cols_kernel(const uint8_t* __restrict__ argmax_map, const uint8_t* __restrict__ slide,
                                                                    int64_t pitch, int64_t dh, int64_t dw, const uint8_t* __restrict__ lut,
                                                                    double alpha, uint8_t* __restrict__ mask_out, uint8_t* __restrict__ thumb_out,
                                                                    uint8_t* __restrict__ overlay_out) {
    constexpr int d = 16 / CP;
    constexpr uint32_t area = d * d;
    const int64_t blocks_per_row = (dw + CP - 1) / CP;
    const int64_t total = dh * blocks_per_row;
    const double beta = __dsub_rn(1.0, alpha);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / blocks_per_row, jb = t - i * blocks_per_row;
        const int64_t j0 = jb * CP;
        const int ncell = (int)(dw - j0 < CP ? dw - j0 : CP);
        const uint8_t* src = slide + (i * d) * pitch + 48 * jb;
        uint32_t sums[CP][3];
#pragma unroll
        for (int k = 0; k < CP; ++k) sums[k][0] = sums[k][1] = sums[k][2] = 0;
        if (ncell == CP) {
#pragma unroll
            for (int r = 0; r < d; ++r) {
                const uint4* q = reinterpret_cast<const uint4*>(src + (int64_t)r * pitch);
                const uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
                rgb_sums3(a.x, a.y, a.z, sums[0][0], sums[0][1], sums[0][2]);                                  // pixels 0-3
                rgb_sums3(a.w, b.x, b.y, sums[4 / d][0], sums[4 / d][1], sums[4 / d][2]);                      // pixels 4-7
                rgb_sums3(b.z, b.w, c.x, sums[8 / d][0], sums[8 / d][1], sums[8 / d][2]);                      // pixels 8-11
                rgb_sums3(c.y, c.z, c.w, sums[12 / d][0], sums[12 / d][1], sums[12 / d][2]);                   // pixels 12-15
            }
        } else {
#pragma unroll
            for (int k = 0; k < CP; ++k) {
                if (k >= ncell) continue;
                for (int r = 0; r < d; ++r) {
                    const uint32_t* q = reinterpret_cast<const uint32_t*>(src + (int64_t)r * pitch + 3 * d * k);
#pragma unroll
                    for (int gq = 0; gq < d / 4; ++gq) rgb_sums3(__ldg(q + 3 * gq), __ldg(q + 3 * gq + 1), __ldg(q + 3 * gq + 2), sums[k][0], sums[k][1], sums[k][2]);
                }
            }
        }
        uint8_t om[3 * CP], ot[3 * CP], oo[3 * CP];
#pragma unroll
        for (int k = 0; k < CP; ++k) {
            const uint32_t cls = k < ncell ? argmax_map[i * dw + j0 + k] : 0u;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint32_t col = lut[3 * cls + c];
                const uint32_t img = (sums[k][c] + area / 2) / area;
                om[3 * k + c] = (uint8_t)col;
                ot[3 * k + c] = (uint8_t)img;
                oo[3 * k + c] = (uint8_t)(int)__dadd_rn(__dmul_rn((double)img, alpha), __dmul_rn((double)col, beta));
            }
        }
        const int64_t o = 3 * (i * dw + j0);
        auto put = [&](uint8_t* out, const uint8_t* v) {
            if (!out) return;
            // full blocks: 3 * CP bytes at a (3 * CP)-byte multiple offset from the row start; rows start 2-byte aligned only when
            // 3 * dw is even, so vector stores are used when the absolute address allows it
            const uintptr_t addr = reinterpret_cast<uintptr_t>(out + o);
            if (ncell == CP && CP == 4 && addr % 4 == 0) {
                uint32_t* w = reinterpret_cast<uint32_t*>(out + o);
#pragma unroll
                for (int q = 0; q < 3; ++q) w[q] = (uint32_t)v[4 * q] | ((uint32_t)v[4 * q + 1] << 8) | ((uint32_t)v[4 * q + 2] << 16) | ((uint32_t)v[4 * q + 3] << 24);
            } else if (ncell == CP && CP == 2 && addr % 2 == 0) {
                uint16_t* w = reinterpret_cast<uint16_t*>(out + o);
#pragma unroll
                for (int q = 0; q < 3; ++q) w[q] = (uint16_t)((uint32_t)v[2 * q] | ((uint32_t)v[2 * q + 1] << 8));
            } else {
#pragma unroll
                for (int q = 0; q < 3 * CP; ++q)
                    if (q < 3 * ncell) out[o + q] = v[q];
            }
        };
        put(mask_out, om);
        put(thumb_out, ot);
        put(overlay_out, oo);
    }
}

static int make_stitch_grid(int64_t H, int64_t W, int ps, int stride, int d, int n, int batch_size, StitchGrid* g) {
    DH_REQUIRE(ps > 0 && stride > 0 && d > 0, "stitch: ps, stride and downscale must be positive");
    DH_REQUIRE(n > 0 && n <= 64, "stitch: n classes %d outside 1..64", n);
    DH_REQUIRE(H >= ps && W >= ps, "stitch: slide smaller than a patch");
    DH_REQUIRE(H + d < (1ll << 31) && W + d < (1ll << 31) && ps < (1 << 30) && stride < (1 << 30), "stitch: sizes exceed 32-bit pixel coordinates");
    g->H = H; g->W = W; g->ps = ps; g->stride = stride; g->d = d; g->n = n;
    g->ny = (H - ps) <= 0 ? 0 : (H - ps + stride - 1) / stride;
    g->nx = (W - ps) <= 0 ? 0 : (W - ps + stride - 1) / stride;
    g->N = g->ny * g->nx + g->ny + g->nx + 1;
    int64_t npad = batch_size > 0 ? (g->N + batch_size - 1) / batch_size * batch_size : g->N;
    g->pads = npad - g->N;
    g->dh = H / d; g->dw = W / d;
    g->lastrow_cell = (H - ps) / d;
    g->lastcol_cell = (W - ps) / d;
    return DH_OK;
}

}  // namespace dh

using namespace dh;

extern "C" DH_API int dh_stitch_dense_ex(const float* logits, int64_t H, int64_t W, int ps, int stride, int d, int n,
                                  int batch_size, float* sum_map, uint32_t* count_map, uint8_t* argmax_u8,
                                  int64_t row_begin, int64_t row_end, void* stream) {
    StitchGrid g;
    int rc = make_stitch_grid(H, W, ps, stride, d, n, batch_size, &g);
    if (rc != DH_OK) return rc;
    DH_REQUIRE(logits, "dh_stitch_dense: null logits");
    DH_REQUIRE(sum_map || argmax_u8 || count_map, "dh_stitch_dense: no output requested");
    DH_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= g.dh, "dh_stitch_dense: rows [%lld,%lld) outside map of %lld rows",
               (long long)row_begin, (long long)row_end, (long long)g.dh);
    DH_REQUIRE(!sum_map || reinterpret_cast<uintptr_t>(sum_map) % 16 == 0, "dh_stitch_dense: sum_map must be 16-byte aligned");
    DH_REQUIRE(!argmax_u8 || n <= 256, "dh_stitch_dense: argmax_u8 needs n <= 256");
    if (row_end == row_begin || g.dw == 0) return DH_OK;
    const int64_t rows = row_end - row_begin;
    cudaStream_t st = as_stream(stream);
    const bool s = sum_map != nullptr, a = argmax_u8 != nullptr, c = count_map != nullptr;
    {
        const bool phased = s && (g.dw * n) % 4 != 0;
        // aligned fast path: column tile = multiple of 16 cells with at most kStitchV float4 per thread per row
        int tj = (kStitchThreads * 4 * kStitchV / n) & ~15;
        if (tj > 2 * kStitchThreads) tj = 2 * kStitchThreads;
        if (tj < 16) tj = 16;  // n <= 64 -> 16 * 64 / 4 = 256 vectors: still one per thread
        const int64_t col_tiles = (g.dw + tj - 1) / tj;
        int64_t rpb = rows * col_tiles / ((int64_t)kNumSMs * 8);
        rpb = rpb < 4 ? 4 : (rpb > 128 ? 128 : rpb);
        {
            // Small maps (the d = 16 map of a 40k^2 slide is 125 MB): 1.x - 3.x waves of short blocks lose a large part of the last wave and
            // pay the block set-up many times. Then: ONE wave of longer blocks, rows per block a multiple of the row-class length so that
            // no class is cut in two (measured at d = 16: 14 rows = 2.8 waves 0.45 of the HBM peak, 42 rows = one wave 0.52;
            // large maps keep the short blocks: one wave of 589-row blocks at d = 4 is 0.91 vs 0.99, profiles/r02_stitch.md)
            const int64_t per_wave = (int64_t)kNumSMs * (phased ? 2 : DH_STITCH_MINB);
            const int64_t blocks = (rows + rpb - 1) / rpb * col_tiles;
            if (blocks > per_wave && blocks < 4 * per_wave && col_tiles <= per_wave) {
                const int64_t groups = per_wave / col_tiles;
                int64_t r1 = (rows + groups - 1) / groups;
                const int64_t cls = stride % d == 0 ? stride / d : 1;
                r1 = (r1 + cls - 1) / cls * cls;
                rpb = r1;
            }
        }
        if (const char* e = getenv("DH_STITCH_RPB")) { const int v = atoi(e); if (v > 0) rpb = v; }   // profiling override
        const int64_t row_groups = (rows + rpb - 1) / rpb;
        DH_REQUIRE(row_groups <= 65535, "dh_stitch_dense: too many row groups (%lld); stitch in bands", (long long)row_groups);
        const int amax_vec = (a && g.dw % 4 == 0 && reinterpret_cast<uintptr_t>(argmax_u8) % 4 == 0) ? 1 : 0;
        dim3 grid((unsigned)col_tiles, (unsigned)row_groups);
        size_t smem = (size_t)tj * n * sizeof(float) + (size_t)tj * sizeof(uint32_t) + (size_t)tj;
        const int cls_off = (int)((smem + 15) / 16 * 4);   // in floats, 16-byte aligned: per-class sums, covers, counts, scan totals, arg max
        smem = (size_t)cls_off * 4 + (size_t)tj * (n > 2 ? n : 2) * 4 + (size_t)tj * 8 + (size_t)tj * 4 + 64 + (size_t)tj;
        // logits a block can touch: patch rows x patch columns of the main grid + last column + last row + corner copies
        const int64_t nr_max = ((rpb - 1) * d + ps - 1) / stride + 2, nc_max = ((int64_t)(tj - 1) * d + ps - 1) / stride + 2;
        const int64_t stage_floats = (nr_max * nc_max + nr_max + nc_max + g.pads + 1) * n;
        int stage_off = 0;
        if ((s || a) && g_stitch_stage && smem + 16 + (size_t)stage_floats * 4 <= 48 * 1024) {
            stage_off = (int)((smem + 15) / 16 * 4);     // in floats, 16-byte aligned
            smem = (size_t)stage_off * 4 + (size_t)stage_floats * 4;
        }
#define DH_STA(S, A, C, P) stitch_dense_aligned_kernel<S, A, C, P><<<grid, kStitchThreads, smem, st>>>(logits, g, sum_map, count_map, argmax_u8, row_begin, row_end, tj, (int)rpb, amax_vec, cls_off, stage_off)
        if (s && phased) { if (a) { if (c) DH_STA(true, true, true, true); else DH_STA(true, true, false, true); } else { if (c) DH_STA(true, false, true, true); else DH_STA(true, false, false, true); } }
        else if (s) { if (a) { if (c) DH_STA(true, true, true, false); else DH_STA(true, true, false, false); } else { if (c) DH_STA(true, false, true, false); else DH_STA(true, false, false, false); } }
        else   { if (a) { if (c) DH_STA(false, true, true, false); else DH_STA(false, true, false, false); } else { DH_STA(false, false, true, false); } }
#undef DH_STA
        DH_CHECK_LAUNCH("stitch_dense_aligned_kernel");
        return DH_OK;
    }
}

extern "C" DH_API int dh_stitch_dense_set_variant(int variant) {
    if (variant < 0 || variant > 1) { set_error("dh_stitch_dense_set_variant: variant must be 0 (logits staged in shared memory per block) or 1 (read per row class)"); return DH_ERR_INVALID; }
    g_stitch_stage = variant == 0 ? 1 : 0;
    return DH_OK;
}

extern "C" DH_API int dh_stitch_dense(const float* logits, int64_t H, int64_t W, int ps, int stride, int d, int n, int batch_size,
                               float* sum_map, uint32_t* count_map, int64_t row_begin, int64_t row_end, void* stream) {
    return dh_stitch_dense_ex(logits, H, W, ps, stride, d, n, batch_size, sum_map, count_map, nullptr, row_begin, row_end, stream);
}

extern "C" DH_API int dh_stitch_scatter(const float* logits, const int32_t* coords, int64_t P, int ps, int d, int n, float* sum_map,
                                 uint32_t* count_map, int64_t rows, int64_t dw, int64_t row_offset, void* stream) {
    if (P == 0) return DH_OK;  // nothing to add (empty tensors have null data pointers)
    DH_REQUIRE(logits && coords, "dh_stitch_scatter: null input");
    DH_REQUIRE(sum_map || count_map, "dh_stitch_scatter: no output requested");
    DH_REQUIRE(ps > 0 && d > 0 && n > 0 && rows >= 0 && dw >= 0 && P >= 0 && row_offset >= 0, "dh_stitch_scatter: bad sizes");
    if (P == 0 || rows == 0 || dw == 0) return DH_OK;
    int grid = (int)(P < (int64_t)kNumSMs * 8 ? P : (int64_t)kNumSMs * 8);
    const int vec_ok = sum_map && reinterpret_cast<uintptr_t>(sum_map) % 16 == 0 ? 1 : 0;  // float4 reductions need a 16-byte aligned map
    stitch_scatter_kernel<<<grid, 256, 0, as_stream(stream)>>>(logits, coords, P, ps, d, n, sum_map, count_map, rows, dw, row_offset, vec_ok);
    DH_CHECK_LAUNCH("stitch_scatter_kernel");
    return DH_OK;
}

extern "C" DH_API int dh_stitch_finalize(const float* sum_map, const uint32_t* count_map, int64_t cells, int n, float* norm_map,
                                  uint8_t* argmax_u8, void* stream) {
    DH_REQUIRE(sum_map, "dh_stitch_finalize: null sum map");
    DH_REQUIRE(norm_map || argmax_u8, "dh_stitch_finalize: no output requested");
    DH_REQUIRE(cells >= 0 && n > 0 && n <= 256, "dh_stitch_finalize: bad sizes");
    if (cells == 0) return DH_OK;
    int64_t blocks = (cells + 255) / 256;
    int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
    stitch_finalize_kernel<<<grid, 256, 0, as_stream(stream)>>>(sum_map, count_map, cells, n, norm_map, argmax_u8);
    DH_CHECK_LAUNCH("stitch_finalize_kernel");
    return DH_OK;
}

extern "C" DH_API int dh_colorize_overlay(const uint8_t* argmax_u8, const uint8_t* slide, int64_t H, int64_t W, int64_t pitch, int64_t dh_,
                                          int64_t dw_, int d, const uint8_t* lut_rgb, double alpha, uint8_t* mask_out, uint8_t* thumb_out,
                                          uint8_t* overlay_out, void* stream) {
    DH_REQUIRE(argmax_u8 && lut_rgb, "dh_colorize_overlay: null class map or colour table");
    DH_REQUIRE(mask_out || thumb_out || overlay_out, "dh_colorize_overlay: no output requested");
    DH_REQUIRE(dh_ >= 0 && dw_ >= 0 && d >= 1 && d <= 4096, "dh_colorize_overlay: bad sizes");
    if (thumb_out || overlay_out) {
        DH_REQUIRE(slide, "dh_colorize_overlay: the thumbnail and the overlay need the slide");
        DH_REQUIRE(dh_ * d <= H && dw_ * d <= W && pitch >= 3 * W, "dh_colorize_overlay: map %lldx%lld at downscale %d does not fit the %lldx%lld slide",
                   (long long)dh_, (long long)dw_, d, (long long)H, (long long)W);
    }
    if (dh_ == 0 || dw_ == 0) return DH_OK;
    if ((thumb_out || overlay_out) && d % 4 == 0 && pitch % 16 == 0 && reinterpret_cast<uintptr_t>(slide) % 16 == 0 && d <= 256) {
        const int64_t cells = dh_ * dw_;  // d <= 256: the per-channel sums (255 * d * d) fit 32 bits
        const int64_t blocks = (cells + 255) / 256;
        const int grid = (int)(blocks < (int64_t)kNumSMs * 32 ? blocks : (int64_t)kNumSMs * 32);
        if (d == 4 || d == 8) {  // 16-pixel column blocks: 128-bit loads, word / half-word stores
            const int cp = 16 / d;
            const int64_t items = dh_ * ((dw_ + cp - 1) / cp);
            const int64_t bl = (items + 255) / 256;
            const int gr = (int)(bl < (int64_t)kNumSMs * 32 ? bl : (int64_t)kNumSMs * 32);
            if (d == 4) colorize_overlay_cols_kernel<4><<<gr, 256, 0, as_stream(stream)>>>(argmax_u8, slide, pitch, dh_, dw_, lut_rgb, alpha, mask_out, thumb_out, overlay_out);
            else colorize_overlay_cols_kernel<2><<<gr, 256, 0, as_stream(stream)>>>(argmax_u8, slide, pitch, dh_, dw_, lut_rgb, alpha, mask_out, thumb_out, overlay_out);
            DH_CHECK_LAUNCH("colorize_overlay_cols_kernel");
            return DH_OK;
        }
        if (d % 16 == 0)
            colorize_overlay_vec_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(argmax_u8, slide, pitch, dh_, dw_, d, lut_rgb, alpha, mask_out,
                                                                             thumb_out, overlay_out);
        else
            colorize_overlay_vec_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(argmax_u8, slide, pitch, dh_, dw_, d, lut_rgb, alpha, mask_out,
                                                                             thumb_out, overlay_out);
        DH_CHECK_LAUNCH("colorize_overlay_vec_kernel");
        return DH_OK;
    }
    const int64_t total = dh_ * dw_ * 3;
    const int64_t blocks = (total + 255) / 256;
    const int grid = (int)(blocks < (int64_t)kNumSMs * 32 ? blocks : (int64_t)kNumSMs * 32);
    colorize_overlay_kernel<<<grid, 256, 0, as_stream(stream)>>>(argmax_u8, slide, pitch, dh_, dw_, d, lut_rgb, alpha, mask_out, thumb_out,
                                                                overlay_out);
    DH_CHECK_LAUNCH("colorize_overlay_kernel");
    return DH_OK;
}
