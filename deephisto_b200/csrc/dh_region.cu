// K3/K4/K7: annotation-polygon acceptance, random region sampling, polygon rasterisation.
// Compiled with -fmad=false: every float64 operation is a separately rounded IEEE op so that the
// CPU restatement (oracle/region.py, numpy float64) reproduces the decisions bit for bit.
//
// Reference path replaced:
//   patch_samplers/region_samplers.py:82-143   RegionAnnotation._extract_patch_coords_rnd
//   patch_samplers/region_samplers.py:145-191  RegionAnnotation._extract_patch_coords_dense
//   patch_samplers/region_samplers.py:525-591  AnnoRegionRndSampler._gen_single_proc (class / region draws)
// The acceptance criterion there is shapely's `polygon.intersection(patch_square).area > ps*ps*ri`.
// GEOS is not part of the reference tree; the area of (polygon ∩ axis-aligned square) is computed here
// in closed form as the boundary integral  A = | sum_e sgn_e * Int_e (clamp(x, xa, xb) - xa) dy |  over
// the polygon edges clipped to ya <= y <= yb (Green's theorem with the square's indicator folded into the
// integrand) -- O(1) per edge, no clipped-polygon construction, no per-thread arrays.
//
// Edge table (built on the host once per dataset, deephisto_b200/geometry.py): 8 doubles per
// non-horizontal edge, oriented so that yA < yB:
//   [xA, yA, xB, yB, m = (xB-xA)/(yB-yA), r = (yB-yA)/(xB-xA) (0 if vertical), sgn (+1: original edge
//    went up in y, -1: down), 0]
#include "dh_common.cuh"

namespace dh {

constexpr int kEdgeStride = 8;

__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }

// signed contribution of one edge to the area of polygon ∩ [xa,xb]x[ya,yb]
__device__ __forceinline__ double edge_term(const double* __restrict__ e, double xa, double xb, double ya, double yb) {
    const double xA = e[0], yA = e[1], xB = e[2], yB = e[3], m = e[4], r = e[5], sgn = e[6];
    const double ys = dmax(yA, ya);
    const double ye = dmin(yB, yb);
    if (!(ys < ye)) return 0.0;
    const double xs = (ys == yA) ? xA : xA + (ys - yA) * m;
    const double xe = (ye == yB) ? xB : xA + (ye - yA) * m;
    const double wx = xb - xa;
    double val;
    if (xs == xe) {
        double g = dmin(dmax(xs, xa), xb) - xa;
        val = (ye - ys) * g;
    } else {
        const double xmin = dmin(xs, xe), xmax = dmax(xs, xe);
        const double ar = fabs(r);
        val = 0.0;
        const double cl = dmax(xmin, xa), ch = dmin(xmax, xb);
        if (cl < ch) val = ((ch - cl) * ar) * (((cl - xa) + (ch - xa)) * 0.5);
        const double ul = dmax(xmin, xb);
        if (ul < xmax) val = val + ((xmax - ul) * ar) * wx;
    }
    return sgn * val;
}

__device__ __forceinline__ double clip_area(const double* __restrict__ edges, int e0, int e1, double x, double y, double ps) {
    const double xa = x, xb = x + ps, ya = y, yb = y + ps;
    double acc = 0.0;
    for (int e = e0; e < e1; ++e) acc = acc + edge_term(edges + (int64_t)e * kEdgeStride, xa, xb, ya, yb);
    return fabs(acc);
}

// ---- D: dense acceptance ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) region_accept_dense_kernel(const double* __restrict__ edges, int e0, int e1,
                                                                  int64_t y0, int64_t x0, int64_t ny, int64_t nx, int stride,
                                                                  int ps, double thr, uint8_t* __restrict__ mask,
                                                                  double* __restrict__ area_out) {
    const int64_t total = ny * nx;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t iy = t / nx, ix = t - iy * nx;
        double y = (double)(y0 + iy * stride), x = (double)(x0 + ix * stride);
        double a = clip_area(edges, e0, e1, x, y, (double)ps);
        mask[t] = a > thr ? 1 : 0;
        if (area_out) area_out[t] = a;
    }
}

// ordered compaction of accepted candidates, single block (candidate grids are small: bbox / stride)
__global__ void __launch_bounds__(1024) compact_coords_kernel(const uint8_t* __restrict__ mask, int64_t y0, int64_t x0,
                                                              int64_t ny, int64_t nx, int stride, int32_t* __restrict__ coords,
                                                              int32_t* __restrict__ n_out) {
    __shared__ int warp_sums[32];
    __shared__ int base;
    const int64_t total = ny * nx;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t t0 = 0; t0 < total; t0 += blockDim.x) {
        int64_t t = t0 + threadIdx.x;
        int flag = (t < total) ? (int)mask[t] : 0;
        unsigned bal = __ballot_sync(0xffffffffu, flag);
        int prefix = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_sums[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            int s = warp_sums[w];
            if (w < wid) woff += s;
            tot += s;
        }
        if (flag) {
            int pos = base + woff + prefix;
            int64_t iy = t / nx, ix = t - iy * nx;
            coords[2 * pos] = (int32_t)(y0 + iy * stride);
            coords[2 * pos + 1] = (int32_t)(x0 + ix * stride);
        }
        __syncthreads();
        if (threadIdx.x == 0) base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = base;
}

// ---- C/G: random region sampling ----------------------------------------------------------------
struct RegionTablesDev {
    const double* edges;
    const int32_t* edge_off;
    const double* reg_bbox;
    const double* reg_area;
    const int32_t* reg_image;
    const int32_t* img_hw;
    const int32_t* tbl_cls_off;
    const int32_t* tbl_cls;
    const int32_t* cat_off;
    const int32_t* cat_region;
    const double* cat_cdf;
    const double* img_cdf;
    int32_t n_tables, n_classes, n_regions, n_images;
};

__device__ __forceinline__ double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

// first index i in [lo, hi) with cdf[i] > u, clamped to hi-1
__device__ __forceinline__ int cdf_search(const double* __restrict__ cdf, int lo, int hi, double u) {
    int a = lo, b = hi - 1;
    while (a < b) {
        int mid = (a + b) >> 1;
        if (cdf[mid] > u) b = mid; else a = mid + 1;
    }
    return a;
}

constexpr int kSampleWarps = 4;
constexpr int kSmemEdges = 96;  // edges of the group's region staged per warp (larger polygons read the rest from L1/L2)

// clip area with the first `n_sm` edges in shared memory and the remainder in global memory (same order, same ops)
__device__ __forceinline__ double clip_area_staged(const double* __restrict__ sm_edges, int n_sm, const double* __restrict__ edges,
                                                   int e0, int e1, double x, double y, double ps) {
    const double xa = x, xb = x + ps, ya = y, yb = y + ps;
    double acc = 0.0;
    for (int e = 0; e < n_sm; ++e) acc = acc + edge_term(sm_edges + e * kEdgeStride, xa, xb, ya, yb);
    for (int e = e0 + n_sm; e < e1; ++e) acc = acc + edge_term(edges + (int64_t)e * kEdgeStride, xa, xb, ya, yb);
    return fabs(acc);
}

__global__ void __launch_bounds__(kSampleWarps * 32) region_sample_kernel(RegionTablesDev T, int64_t n_slots, int k, int ps, double thr,
                                                                         int miss_limit, int max_redraw, int fixed_class,
                                                                         int64_t slots_per_table_draw, uint32_t key0, uint32_t key1,
                                                                         uint64_t slot_offset, int32_t* __restrict__ coords_out,
                                                                         int64_t* __restrict__ label_out, int32_t* __restrict__ image_out,
                                                                         uint8_t* __restrict__ status_out) {
    __shared__ double s_edges[kSampleWarps][kSmemEdges * kEdgeStride];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t n_groups = (n_slots + k - 1) / k;
    const int64_t group = (int64_t)blockIdx.x * kSampleWarps + wib;
    if (group >= n_groups) return;
    const int64_t s0 = group * k;
    const int kk = (int)((n_slots - s0) < k ? (n_slots - s0) : k);
    const uint64_t g0 = slot_offset + (uint64_t)s0;  // global index of the group's first slot
    double* my_edges = s_edges[wib];

    // lanes = k slots x apr attempts per round; slot of this lane and its attempt sub-index
    const int apr = 32 / k;
    const int slot = lane / apr, sub = lane - slot * apr;
    const bool lane_live = slot < kk;
    const unsigned slot_mask = (apr == 32 ? 0xffffffffu : ((1u << apr) - 1u)) << (slot * apr & 31);

    // table (image) choice: one draw per chunk of slots_per_table_draw global slots
    int table = 0;
    if (T.n_tables > 1) {
        uint64_t chunk = g0 / (uint64_t)slots_per_table_draw;
        Philox4 pt = philox4x32_10((uint32_t)chunk, (uint32_t)(chunk >> 32), 0u, kStreamTable, key0, key1);
        table = cdf_search(T.img_cdf, 0, T.n_tables, u01(pt.v[0]));
    }
    const int cls_lo = T.tbl_cls_off[table], ncls = T.tbl_cls_off[table + 1] - cls_lo;
    const uint64_t gs = g0 + (uint64_t)(lane_live ? slot : 0);

    uint8_t fail = DH_SLOT_MISS_LIMIT;
    for (int rd = 0; rd < max_redraw; ++rd) {
        Philox4 pg = philox4x32_10((uint32_t)g0, (uint32_t)(g0 >> 32), (uint32_t)rd, kStreamGroup, key0, key1);
        int cls = fixed_class >= 0 ? fixed_class : T.tbl_cls[cls_lo + (int)bounded_u32(pg.v[0], (uint32_t)ncls)];
        const int cat = table * T.n_classes + cls;
        const int rlo = T.cat_off[cat], rhi = T.cat_off[cat + 1];
        if (rhi <= rlo) { fail = DH_SLOT_EMPTY_RANGE; continue; }
        const int region = T.cat_region[cdf_search(T.cat_cdf, rlo, rhi, u01(pg.v[1]))];
        // region_samplers.py:117-118 "Region is too small."
        if (T.reg_area[region] < thr) { fail = DH_SLOT_MISS_LIMIT; continue; }
        const double bx0 = T.reg_bbox[4 * region], by0 = T.reg_bbox[4 * region + 1];
        const double bx1 = T.reg_bbox[4 * region + 2], by1 = T.reg_bbox[4 * region + 3];
        const int img = T.reg_image[region];
        const int h = T.img_hw[2 * img], w = T.img_hw[2 * img + 1];
        // region_samplers.py:123-124  randint(x0, min(max(x0 + 1, x1 - ps), w)), float bounds truncated;
        // upper bound additionally clamped to w - ps + 1 and lower bound to 0 (never read outside the slide)
        const double dps = (double)ps;
        int64_t xlo = (int64_t)bx0, ylo = (int64_t)by0;
        int64_t xhi = (int64_t)dmin(dmax(bx0 + 1.0, bx1 - dps), (double)w);
        int64_t yhi = (int64_t)dmin(dmax(by0 + 1.0, by1 - dps), (double)h);
        if (xhi > (int64_t)w - ps + 1) xhi = (int64_t)w - ps + 1;
        if (yhi > (int64_t)h - ps + 1) yhi = (int64_t)h - ps + 1;
        if (xlo < 0) xlo = 0;
        if (ylo < 0) ylo = 0;
        if (xhi <= xlo || yhi <= ylo) { fail = DH_SLOT_EMPTY_RANGE; continue; }
        const uint32_t xr = (uint32_t)(xhi - xlo), yr = (uint32_t)(yhi - ylo);
        const int e0 = T.edge_off[region], e1 = T.edge_off[region + 1];

        // stage the region's edge table in shared memory (coalesced), once per draw
        const int n_sm = (e1 - e0) < kSmemEdges ? (e1 - e0) : kSmemEdges;
        __syncwarp();
        for (int i = lane; i < n_sm * kEdgeStride; i += 32) my_edges[i] = T.edges[(int64_t)e0 * kEdgeStride + i];
        __syncwarp();

        // every round evaluates attempts [a0, a0 + apr) of all unfinished slots at once; per slot the lowest accepted attempt
        // index wins, which is what the reference's sequential rejection loop returns for the same draws
        bool done = !lane_live;
        int my_y = 0, my_x = 0;
        unsigned pending = __ballot_sync(0xffffffffu, !done);
        for (int a0 = 0; a0 < miss_limit && pending; a0 += apr) {
            const int a = a0 + sub;
            bool ok = false;
            int x = 0, y = 0;
            if (!done && a < miss_limit) {
                Philox4 pa = philox4x32_10((uint32_t)gs, (uint32_t)(gs >> 32), ((uint32_t)rd << 16) | (uint32_t)a, kStreamAttempt, key0, key1);
                x = (int)(xlo + (int64_t)bounded_u32(pa.v[0], xr));
                y = (int)(ylo + (int64_t)bounded_u32(pa.v[1], yr));
                ok = clip_area_staged(my_edges, n_sm, T.edges, e0, e1, (double)x, (double)y, dps) > thr;
            }
            const unsigned hits = __ballot_sync(0xffffffffu, ok) & slot_mask;
            const int src = hits ? __ffs(hits) - 1 : lane;
            const int yy = __shfl_sync(0xffffffffu, y, src), xx = __shfl_sync(0xffffffffu, x, src);
            if (!done && hits) { my_y = yy; my_x = xx; done = true; }
            pending = __ballot_sync(0xffffffffu, !done);
        }
        if (pending) { fail = DH_SLOT_MISS_LIMIT; continue; }
        if (lane_live && sub == 0) {
            const int64_t o = s0 + slot;
            coords_out[2 * o] = my_y;
            coords_out[2 * o + 1] = my_x;
            if (label_out) label_out[o] = cls;
            if (image_out) image_out[o] = img;
            if (status_out) status_out[o] = DH_SLOT_OK;
        }
        return;
    }
    if (lane_live && sub == 0) {
        const int64_t o = s0 + slot;
        coords_out[2 * o] = 0;
        coords_out[2 * o + 1] = 0;
        if (label_out) label_out[o] = -1;
        if (image_out) image_out[o] = -1;
        if (status_out) status_out[o] = fail;
    }
}

// ---- R: rasterisation (pixel-centre even-odd rule) -------------------------------------------------
__global__ void __launch_bounds__(256) rasterize_kernel(const double* __restrict__ edges, const int32_t* __restrict__ edge_off,
                                                        const double* __restrict__ reg_bbox, int n_regions, double scale,
                                                        int32_t* __restrict__ label, int64_t mh, int64_t mw) {
    const int64_t total = mh * mw;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t my = t / mw, mx = t - my * mw;
        const double px = ((double)mx + 0.5) * scale, py = ((double)my + 0.5) * scale;
        int32_t lab = 0;
        for (int r = 0; r < n_regions; ++r) {
            const double* bb = reg_bbox + 4 * r;
            if (px < bb[0] || px > bb[2] || py < bb[1] || py > bb[3]) continue;
            int crossings = 0;
            for (int e = edge_off[r]; e < edge_off[r + 1]; ++e) {
                const double* ed = edges + (int64_t)e * kEdgeStride;
                // half-open in y so that a vertex shared by two edges is counted once
                if (ed[1] <= py && py < ed[3]) {
                    double xi = ed[0] + (py - ed[1]) * ed[4];
                    if (xi > px) ++crossings;
                }
            }
            if (crossings & 1) lab = r + 1;
        }
        label[t] = lab;
    }
}

}  // namespace dh

using namespace dh;

extern "C" DH_API int dh_region_accept_dense(const double* edges, int edge_begin, int edge_end, int64_t y0, int64_t x0, int64_t ny,
                                      int64_t nx, int stride, int ps, double threshold, uint8_t* mask_u8, double* area_out,
                                      void* stream) {
    DH_REQUIRE(edges && mask_u8, "dh_region_accept_dense: null pointer");
    DH_REQUIRE(ny >= 0 && nx >= 0 && stride > 0 && ps > 0, "dh_region_accept_dense: bad candidate grid");
    DH_REQUIRE(edge_begin >= 0 && edge_end >= edge_begin, "dh_region_accept_dense: bad edge range [%d, %d)", edge_begin, edge_end);
    if (ny == 0 || nx == 0) return DH_OK;
    int64_t total = ny * nx;
    int64_t blocks = (total + 255) / 256;
    int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
    region_accept_dense_kernel<<<grid, 256, 0, as_stream(stream)>>>(edges, edge_begin, edge_end, y0, x0, ny, nx, stride, ps, threshold,
                                                                    mask_u8, area_out);
    DH_CHECK_LAUNCH("region_accept_dense_kernel");
    return DH_OK;
}

extern "C" DH_API int dh_compact_coords(const uint8_t* mask_u8, int64_t y0, int64_t x0, int64_t ny, int64_t nx, int stride,
                                 int32_t* coords_out, int32_t* n_out, void* stream) {
    DH_REQUIRE(mask_u8 && coords_out && n_out, "dh_compact_coords: null pointer");
    DH_REQUIRE(ny >= 0 && nx >= 0 && ny * nx < (1ll << 31), "dh_compact_coords: bad grid");
    compact_coords_kernel<<<1, 1024, 0, as_stream(stream)>>>(mask_u8, y0, x0, ny, nx, stride, coords_out, n_out);
    DH_CHECK_LAUNCH("compact_coords_kernel");
    return DH_OK;
}

extern "C" DH_API int dh_region_sample(const dh_region_tables* t, int64_t n_slots, int k, int ps, double threshold, int miss_limit,
                                int max_redraw, int fixed_class, int64_t slots_per_table_draw, uint64_t seed,
                                uint64_t slot_offset, int32_t* coords_out, int64_t* label_out, int32_t* image_out,
                                uint8_t* status_out, void* stream) {
    if (n_slots <= 0) return DH_OK;
    DH_REQUIRE(t && coords_out, "dh_region_sample: null pointer");
    DH_REQUIRE(t->edges && t->edge_off && t->reg_bbox && t->reg_area && t->reg_image && t->img_hw && t->tbl_cls_off && t->tbl_cls &&
                   t->cat_off && t->cat_region && t->cat_cdf,
               "dh_region_sample: incomplete tables");
    DH_REQUIRE(t->n_tables >= 1 && t->n_classes >= 1 && t->n_regions >= 1, "dh_region_sample: empty tables");
    DH_REQUIRE(t->n_tables == 1 || t->img_cdf, "dh_region_sample: img_cdf required with more than one table");
    DH_REQUIRE(k >= 1 && k <= 32, "dh_region_sample: patches_from_one_region %d outside 1..32", k);
    DH_REQUIRE(ps > 0 && miss_limit >= 1 && miss_limit <= 65536, "dh_region_sample: bad ps / miss_limit");
    DH_REQUIRE(max_redraw >= 1 && max_redraw <= 65535, "dh_region_sample: max_redraw outside 1..65535");
    DH_REQUIRE(fixed_class < t->n_classes, "dh_region_sample: class index %d out of range", fixed_class);
    DH_REQUIRE(slots_per_table_draw >= 1, "dh_region_sample: slots_per_table_draw must be >= 1");
    if (n_slots <= 0) return DH_OK;
    RegionTablesDev T;
    T.edges = t->edges; T.edge_off = t->edge_off; T.reg_bbox = t->reg_bbox; T.reg_area = t->reg_area; T.reg_image = t->reg_image;
    T.img_hw = t->img_hw; T.tbl_cls_off = t->tbl_cls_off; T.tbl_cls = t->tbl_cls; T.cat_off = t->cat_off;
    T.cat_region = t->cat_region; T.cat_cdf = t->cat_cdf; T.img_cdf = t->img_cdf;
    T.n_tables = t->n_tables; T.n_classes = t->n_classes; T.n_regions = t->n_regions; T.n_images = t->n_images;
    int64_t groups = (n_slots + k - 1) / k;
    int64_t blocks = (groups + kSampleWarps - 1) / kSampleWarps;
    DH_REQUIRE(blocks < (1ll << 31), "dh_region_sample: too many slots");
    region_sample_kernel<<<(unsigned)blocks, kSampleWarps * 32, 0, as_stream(stream)>>>(
        T, n_slots, k, ps, threshold, miss_limit, max_redraw, fixed_class < 0 ? -1 : fixed_class, slots_per_table_draw, (uint32_t)seed,
        (uint32_t)(seed >> 32), slot_offset, coords_out, label_out, image_out, status_out);
    DH_CHECK_LAUNCH("region_sample_kernel");
    return DH_OK;
}

extern "C" DH_API int dh_rasterize_polygons(const double* edges, const int32_t* edge_off, const double* reg_bbox, int n_regions, double scale,
                                     int32_t* label_out, int64_t mh, int64_t mw, void* stream) {
    DH_REQUIRE(edges && edge_off && reg_bbox && label_out, "dh_rasterize_polygons: null pointer");
    DH_REQUIRE(n_regions >= 0 && scale > 0 && mh >= 0 && mw >= 0, "dh_rasterize_polygons: bad sizes");
    if (mh == 0 || mw == 0) return DH_OK;
    int64_t total = mh * mw;
    int64_t blocks = (total + 255) / 256;
    int grid = (int)(blocks < (int64_t)kNumSMs * 8 ? blocks : (int64_t)kNumSMs * 8);
    rasterize_kernel<<<grid, 256, 0, as_stream(stream)>>>(edges, edge_off, reg_bbox, n_regions, scale, label_out, mh, mw);
    DH_CHECK_LAUNCH("rasterize_kernel");
    return DH_OK;
}
