/*
 * deephisto_b200.h -- C-ABI of libdeephisto_b200.so
 *
 * B200 (sm_100a) implementation of the DeepHisto patch-sampling / patched-prediction hot path.
 * Every entry point replaces one piece of the reference's Python/numpy/GEOS host path; the
 * reference location is cited as path:line relative to the xubiker/deephisto tree.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every call only enqueues work: no allocation, no synchronisation, no global state
 *     (a per-thread error string and a per-process tensor-map cache are the only statics);
 *   - return value: 0 (DH_OK) or a negative dh_status; dh_last_error() gives the text;
 *   - coordinates are int32 pairs (y, x) of the patch's top-left corner, like the reference's
 *     `(y, x)` tuples (full_samplers.py:380-397, region_samplers.py:135,190);
 *   - the slide is uint8 [H][W][3] RGB with a row pitch in bytes (pitch >= 3*W).
 */
#ifndef DEEPHISTO_B200_H
#define DEEPHISTO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DH_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define DH_API __attribute__((visibility("default")))
#else
#define DH_API
#endif

typedef enum dh_status {
    DH_OK = 0,
    DH_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, misaligned buffer) */
    DH_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
    DH_ERR_UNSUPPORTED = -3, /* combination not implemented */
    DH_ERR_NO_DEVICE = -4    /* no sm_100 device / kernel image cannot run here */
} dh_status;

typedef enum dh_dtype { DH_F32 = 0, DH_BF16 = 1, DH_U8 = 2 } dh_dtype;
typedef enum dh_layout {
    DH_NHWC = 0,
    DH_NCHW = 1,
    /* dh_gather_normalize, bf16 only: the patch as a 2x2 SPACE-TO-DEPTH image with 16 channels and room for a zero border,
     * out[b][ps/2 + 3][ps/2 + 3][16]: channel p*8 + q*3 + c of pixel (y' + 2, x' + 2) = channel c of patch pixel (2y' + p, 2x' + q);
     * channels 6, 7, 14, 15 are written as zero. The border pixels (2 top / left, 1 bottom / right) are NOT written: pass a buffer
     * whose border is zero. It is the input of a 4x4 stride-1 convolution that equals the ResNet stem's 7x7 stride-2 convolution
     * (examples/predict_full_patched.py:66-78 batch_predictor -> model.conv1), with 16 instead of 3 input channels. */
    DH_S2D16 = 2,
    /* dh_gather_normalize, bf16 only: the 4x4 space-to-depth image out[b][ps/4][ps/4][48] (ps % 4 == 0): channel p*12 + q*3 + c of
     * block (Y, X) = channel c of patch pixel (4Y + p, 4X + q). The same bytes as DH_NHWC in another order, no padding. Input of a
     * 3x3 stride-1 padding-1 convolution with 4 x 64 output channels (2x2 output pixels per block) that equals the 7x7 stride-2 stem:
     * 48 input and 256 output channels keep cuDNN's tensor-core kernels busy (3x faster than the 16-channel variant). */
    DH_S2D48 = 3
} dh_layout;

/* flip bits for dh_gather_normalize (train.py:71-81 RandomHorizontalFlip / RandomVerticalFlip) */
#define DH_FLIP_H 1u
#define DH_FLIP_V 2u

/* per-slot status written by the region samplers */
#define DH_SLOT_OK 0u
#define DH_SLOT_MISS_LIMIT 1u  /* region_samplers.py:139-142 "Miss limit reached" */
#define DH_SLOT_EMPTY_RANGE 2u /* np.random.randint(low >= high) would raise, region_samplers.py:123-124 */

DH_API int dh_version(void);
DH_API const char* dh_last_error(void);
/* 0 if the current device can run the sm_100a kernels, DH_ERR_NO_DEVICE otherwise. */
DH_API int dh_device_check(void);

/* ------------------------------------------------------------------------------------------
 * Synthetic slide: byte k of the logical [H][W*3] array is
 *   (mix32(seed_lo ^ (k>>2)) ^ seed_hi-fold) >> (8*(k&3)) & 255,  see oracle/synth.py.
 * Replaces: psimage get_region_from_layer of the whole layer (full_samplers.py:53-55,328-330)
 * for benchmark inputs (no 30 GB host->device copy for the 100k x 100k case).
 * ------------------------------------------------------------------------------------------ */
DH_API int dh_synth_slide(uint8_t* slide, int64_t H, int64_t W, int64_t pitch, uint64_t seed, void* stream);
/* rows [y0, y0+rows) of the same H x W slide into a buffer whose row 0 is slide row y0 (row bands) */
DH_API int dh_synth_slide_rows(uint8_t* slide, int64_t H, int64_t W, int64_t pitch, int64_t y0, int64_t rows, uint64_t seed,
                               void* stream);

/* Slide ingestion for the annotated samplers: the reference reads each patch from storage (region_samplers.py:513-520), so only
 * pixels inside annotated regions ever travel. dh_upload_rects copies the listed rectangles of a HOST slide (same row pitch as the
 * device slide; pinned memory makes the copies asynchronous) into the device slide: rects_host is a HOST array [n_rects][4] =
 * {y0, y1, byte_x0, byte_x1}; one cudaMemcpy2DAsync per rectangle on `stream`. Bytes outside the rectangles are left untouched. */
DH_API int dh_upload_rects(uint8_t* slide_dev, int64_t H, int64_t pitch, const uint8_t* slide_host, int64_t n_rects,
                           const int64_t* rects_host, void* stream);

/* Zero-copy ingestion: the device-visible address of a PAGE-LOCKED host buffer (cudaHostAlloc / cudaHostRegister under unified
 * addressing). The gather kernels accept it as `slide`: their bulk copies then read the patch rows straight from host memory over
 * PCIe -- the reference reads a patch from storage at the moment it is drawn (region_samplers.py:513-520), and a short job that
 * touches a fraction of the slide should not wait for the whole layer to be uploaded (full_samplers.py:53-55). Host-only call:
 * writes the address to *device_ptr_out_host; DH_ERR_UNSUPPORTED when the pointer is not mapped page-locked host memory. */
DH_API int dh_host_device_pointer(const void* host_ptr, uint64_t* device_ptr_out_host);

/* ------------------------------------------------------------------------------------------
 * A1  FullImageDenseSampler._create_batched_coords (full_samplers.py:374-404)
 * Enumeration: main grid (y outer, x inner), last column, last row, corner, then the last batch
 * is padded with copies of the corner.
 * dh_dense_count: host-only arithmetic; returns N (unpadded) and writes N rounded up to a
 * multiple of batch_size to *n_padded_host (may be NULL). Returns <0 on invalid arguments.
 * dh_dense_coords: writes coords[i] for i in [first, first+count) of the PADDED enumeration.
 * ------------------------------------------------------------------------------------------ */
DH_API int64_t dh_dense_count(int64_t H, int64_t W, int ps, int stride, int batch_size, int64_t* n_padded_host);
DH_API int dh_dense_coords(int64_t H, int64_t W, int ps, int stride, int batch_size, int64_t first, int64_t count,
                    int32_t* coords_out /* [count][2] */, void* stream);

/* ------------------------------------------------------------------------------------------
 * A2/A3/H  gather + normalise
 *   FullImageDenseSampler._generate_batch_memory + generator_torch (full_samplers.py:353-369,437-452)
 *   FullImageRndSampler._extract_patches_np + generator_torch     (full_samplers.py:187-202,282-290)
 *   batch_predictor's stack / 255 / permute                        (examples/predict_full_patched.py:66-71)
 *   _gen_single_proc_torch's tensor(data)/255                      (region_samplers.py:616)
 * out[b] = patch at coords[b]; value = scale255 ? float(u8)/255 (IEEE fp32 division) : float(u8);
 * then, if mean3_host/std3_host are given, (value - mean[c]) / std[c] in IEEE fp32.
 * out_dtype DH_F32 | DH_BF16 (round-to-nearest-even of the fp32 value) | DH_U8 (raw copy; scale,
 * mean, std ignored). out_layout DH_NHWC [B][ps][ps][3] or DH_NCHW [B][3][ps][ps].
 * out_index (optional): patch b is written to slot out_index[b] instead of b.
 * flip (optional): per-patch DH_FLIP_H / DH_FLIP_V bits applied to the written patch.
 * Pixels outside the slide are read as 0.
 * ------------------------------------------------------------------------------------------ */
DH_API int dh_gather_normalize(const uint8_t* slide, int64_t H, int64_t W, int64_t pitch, const int32_t* coords,
                        const int32_t* out_index, int64_t B, int ps, void* out, int out_dtype, int out_layout,
                        int scale255, const float* mean3_host, const float* std3_host, const uint8_t* flip,
                        void* stream);

/* The same gather over SEVERAL resident slides in one launch (datasets of many images, region_samplers.py:484-523 opens the
 * image of every region): slides_dev is a device table int64 [n_slides][4] = {data pointer, H, W, pitch} (16-byte aligned),
 * slides_host the same table in host memory (validated here), image_of_patch the slide index of every patch (device int32 [B]).
 * Needs every slide 16-byte aligned with pitch % 16 == 0 and a patch size the TMA-staged kernel takes (ps % 4 == 0 for f32,
 * ps % 8 == 0 for bf16); otherwise returns DH_ERR_UNSUPPORTED and the caller gathers slide by slide. */
DH_API int dh_gather_normalize_multi(const int64_t* slides_host, const int64_t* slides_dev, int n_slides,
                                     const int32_t* image_of_patch, const int32_t* coords, const int32_t* out_index, int64_t B, int ps,
                                     void* out, int out_dtype, int out_layout, int scale255, const float* mean3_host,
                                     const float* std3_host, const uint8_t* flip, void* stream);

/* Variant selector for profiling: 0 = auto, 1 = direct (LDG/STG) kernel, 2 = TMA-staged kernel;
 * 3 / 4 / 5 = TMA-staged kernel with the loads / the stores / both switched off (WRONG RESULTS: ceiling measurements only);
 * 6 = TMA-staged kernel with default-policy instead of streaming stores; 7 / 8 / 9 = contiguous tile range per CTA /
 * L2 evict-first hint on the bulk loads / both (correct results; profiling). */
DH_API int dh_gather_set_variant(int variant);

/* ------------------------------------------------------------------------------------------
 * A4  ImagePredictorPatched.process (examples/predict_full_patched.py:40-63)
 *   prediction[y//d:(y+ps)//d, x//d:(x+ps)//d, :] += logits[i]   in sampler order, then argmax.
 * dh_stitch_dense: gather formulation for the dense enumeration of A1 -- every output cell sums the
 *   patches that cover it in reference order (fp32 adds, no FMA) => bit-exact sum map. logits is
 *   [n_padded][n]; the (n_padded - N) padding duplicates of the corner are added again, as the
 *   reference does, unless batch_size <= 0 (no padding).
 *   Only map rows [row_begin, row_end) are produced; sum_map / count_map point at row `row_begin`.
 * dh_stitch_scatter: arbitrary coordinates (random sampler); fp32 atomics, order not preserved.
 *   The map passed holds rows [row_offset, row_offset+rows) of the full [dh][dw] map.
 * dh_stitch_finalize: count normalisation (sum / max(count,1)) and argmax (first maximum, like
 *   np.argmax) over the n classes; any of norm_map / argmax_u8 / count_map may be NULL.
 * ------------------------------------------------------------------------------------------ */
DH_API int dh_stitch_dense(const float* logits, int64_t H, int64_t W, int ps, int stride, int d, int n,
                    int batch_size, float* sum_map, uint32_t* count_map, int64_t row_begin, int64_t row_end,
                    void* stream);
/* same, with the argmax fused into the store epilogue; sum_map and/or count_map may be NULL
 * (argmax-only mode writes dh*dw bytes instead of dh*dw*n*4). */
DH_API int dh_stitch_dense_ex(const float* logits, int64_t H, int64_t W, int ps, int stride, int d, int n,
                       int batch_size, float* sum_map, uint32_t* count_map, uint8_t* argmax_u8,
                       int64_t row_begin, int64_t row_end, void* stream);
DH_API int dh_stitch_scatter(const float* logits, const int32_t* coords, int64_t P, int ps, int d, int n,
                      float* sum_map, uint32_t* count_map, int64_t rows, int64_t dw, int64_t row_offset,
                      void* stream);
DH_API int dh_stitch_finalize(const float* sum_map, const uint32_t* count_map, int64_t cells, int n,
                       float* norm_map, uint8_t* argmax_u8, void* stream);
/* dh_stitch_binned: the same arbitrary coordinate list, DETERMINISTIC and bit-identical to the reference loop
 * (predict_full_patched.py:47-54): patches are binned into warp-sized map tiles and every cell adds its covering
 * patches in ascending list index with plain fp32 adds from 0 -- no atomics on the map, each output written once.
 * The outputs are overwritten (not accumulated into): pass the whole list of a slide (all batches concatenated).
 * Any of sum_map / count_map / argmax_u8 may be NULL (argmax of more than 8 classes needs sum_map). The maps hold
 * rows [row_offset, row_offset+rows) of the full map. scratch: device bytes from dh_stitch_binned_scratch_bytes.
 * dh_stitch_binned_set_tile_rows: profiling override of the tile height (0 = heuristic). */
DH_API int64_t dh_stitch_binned_scratch_bytes(int64_t P, int ps, int d, int n, int64_t rows, int64_t dw);
DH_API int dh_stitch_binned(const float* logits, const int32_t* coords, int64_t P, int ps, int d, int n,
                     float* sum_map, uint32_t* count_map, uint8_t* argmax_u8, int64_t rows, int64_t dw,
                     int64_t row_offset, void* scratch, int64_t scratch_bytes, void* stream);
DH_API int dh_stitch_binned_set_tile_rows(int rows);
/* dh_stitch_dense A/B switch (same bits): 0 = a block stages the logits its rows and cells can touch in shared memory once (default
 * whenever they fit 48 KB), 1 = every row class reads its covering patches from HBM / L2 (round-1 behaviour). */
DH_API int dh_stitch_dense_set_variant(int variant);
/* Tile-kernel formulation (same bits every way, tests/test_gpu_parity.py). 0 = auto: for sum maps of n <= 8 classes the cell-lane
 * kernel (a lane owns one cell and its n class sums; the run's row image goes through shared memory) on 16-byte aligned rows with
 * footprints under 2048 floats and on unaligned rows with footprints under 24 cells, the segment kernel (one lane per (row run,
 * column segment) region) on unaligned rows with wider footprints, the row-run kernels (every lane re-sums its own floats at each
 * footprint boundary) otherwise; 1 = row-run kernels only; 2 = segment kernel wherever it applies; 3 = cell-lane kernel wherever it
 * applies; 4 = profiling only: the cell-lane kernel without its stores. Measurements: profiles/r02_stitch.md. */
DH_API int dh_stitch_binned_set_variant(int variant);

/* ------------------------------------------------------------------------------------------
 * A3  batch_predictor's model forward (examples/predict_full_patched.py:66-78; ResNet18 stem, models/patch_cls_simple/model.py:5-11):
 * the 3x3 / stride 2 / padding 1 max pooling behind conv1 + ReLU, over an NHWC tensor [B][H][W][C] -> [B][(H-1)/2+1][(W-1)/2+1][C].
 * -inf padding, NaN propagates (torch.nn.functional.max_pool2d semantics). dtype: DH_BF16; C % 8 == 0; 16-byte aligned buffers.
 * The convolutions stay with cuDNN (torch); this is the HBM-bound step between them.
 * ------------------------------------------------------------------------------------------ */
DH_API int dh_maxpool3x3s2_nhwc(const void* in, int64_t B, int H, int W, int C, void* out, int dtype, void* stream);
/* The same pooling when the convolution output is stored DEPTH-TO-SPACE, as the 4x4 space-to-depth stem produces it:
 * in [B][H][W][4*C], channel (P*2 + Q)*C + o of block (Y, X) = pixel (2Y + P, 2X + Q), channel o of the [2H][2W][C] image;
 * out [B][H][W][C] = max_pool2d(kernel 3, stride 2, padding 1) of that image. */
DH_API int dh_maxpool3x3s2_d2s(const void* in, int64_t B, int H, int W, int C, void* out, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Prediction post-processing: perform_and_save_visualizations (examples/predict_full_patched.py:81-113).
 *   mask    [dh][dw][3] = lut_rgb[class]                                   (:89-95)
 *   thumb   [dh][dw][3] = integer area average of the d x d slide block under the cell, rounded half up
 *                         (stands in for psim.get_region(..., target_hw=(h, w)), :103-104; psimage's filter is unknown)
 *   overlay [dh][dw][3] = uint8(thumb * alpha + mask * (1 - alpha)), float64, truncated   (:108-110, alpha = 0.6)
 * lut_rgb: device uint8 [256][3]. Any of the three outputs may be NULL; slide may be NULL when only the mask is requested.
 * ------------------------------------------------------------------------------------------ */
DH_API int dh_colorize_overlay(const uint8_t* argmax_u8, const uint8_t* slide, int64_t H, int64_t W, int64_t pitch, int64_t dh,
                               int64_t dw, int d, const uint8_t* lut_rgb, double alpha, uint8_t* mask_out, uint8_t* thumb_out,
                               uint8_t* overlay_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * B1-B3  FullImageRndSampler (full_samplers.py:81-94,105-114,125-162,263-274)
 * Coverage-driven random sampling on the 1/speedup coarse accumulator.
 * One call = one batch: eligible cells (accum < dense_level) are taken in index order, topped
 * up with random non-eligible cells when fewer than B, B distinct cells are drawn (partial
 * Fisher-Yates, Philox4x32-10 keyed by seed, counters documented in oracle/cover.py), jittered,
 * clamped, written to coords_out, and the accumulator footprint of every patch is incremented.
 * scratch: uint32 [dh_cover_scratch_words(dh, dw)] -- it holds STATE between calls (eligibility bitmask, per-block eligible
 * counts, non-zero count), maintained incrementally by the accumulator update, so a batch does not rescan the coarse grid.
 * The state is built from `accum` by the call with batch_index == 0, or explicitly by dh_cover_init (resuming a run from a
 * restored accumulator). dense_level must not change between calls. nonzero_out: device uint32 count of non-zero accumulator
 * cells after the update (filled_ratio = nonzero / (dh*dw)). B <= 2048.
 * ------------------------------------------------------------------------------------------ */
DH_API int64_t dh_cover_scratch_words(int64_t dh, int64_t dw);
DH_API int dh_cover_init(const uint32_t* accum, int64_t dh, int64_t dw, int dense_level, uint32_t* scratch, void* stream);
/* One batch = ONE launch. stop_when_full != 0: a call made when every coarse cell is already covered changes nothing (coords_out
 * untouched) and only reports the count -- the reference's loop ends at filled_ratio >= 1 (full_samplers.py:263-274), so a host
 * can enqueue several batches before it reads a count back. */
DH_API int dh_cover_sample(uint32_t* accum, int64_t dh, int64_t dw, int64_t H, int64_t W, int ps, int speedup,
                    int dense_level, int B, uint64_t seed, uint64_t batch_index, int32_t* coords_out,
                    uint32_t* nonzero_out, uint32_t* scratch, int stop_when_full, void* stream);
/* n_batches consecutive batches (batch_index = first_batch_index + g) enqueued by one call, each with stop_when_full semantics:
 * coords_out [n_batches][B][2] (pre-zeroed by the caller: batches after full coverage are left untouched), nonzero_out [n_batches]. */
DH_API int dh_cover_sample_group(uint32_t* accum, int64_t dh, int64_t dw, int64_t H, int64_t W, int ps, int speedup,
                          int dense_level, int B, uint64_t seed, uint64_t first_batch_index, int n_batches,
                          int32_t* coords_out, uint32_t* nonzero_out, uint32_t* scratch, void* stream);
/* Profiling / tests: 0 = auto (one persistent launch per group, state as a count tree in shared memory), 1 = one launch per batch
 * with a full scan of the block counts (the only variant for coarse grids beyond 67 M cells). Same coordinates either way. */
DH_API int dh_cover_set_variant(int variant);

/* ------------------------------------------------------------------------------------------
 * C/D  RegionAnnotation._extract_patch_coords_dense / _rnd (region_samplers.py:82-191)
 * Acceptance = float64 area(polygon ∩ [x,x+ps]x[y,y+ps]) > threshold (strict), threshold =
 * ps*ps*region_intersection computed by the caller exactly as the reference does (:134,189).
 * The area is the boundary integral of the clamped polygon edges (every op separately rounded;
 * oracle/region.py restates the same operation order; shapely/GEOS is absent from the reference
 * tree). Polygons are passed as an EDGE TABLE built once on the host (deephisto_b200/geometry.py):
 *   edges [E][8] float64 = xA, yA, xB, yB (yA < yB), m = (xB-xA)/(yB-yA), r = (yB-yA)/(xB-xA)
 *                          (0 for vertical edges), sgn (+1 if the polygon edge ran A->B, else -1), 0
 *   horizontal edges are dropped; region i owns edges [edge_off[i], edge_off[i+1]).
 *
 * dh_region_accept_dense: candidates (y0 + iy*stride, x0 + ix*stride), iy < ny, ix < nx, row-major;
 *   writes mask_u8[ny*nx] (1 = accepted) and, if area_out != NULL, the clip area per candidate.
 * dh_compact_coords: ordered compaction of the accepted candidates into coords_out, count to *n_out
 *   (device int32); single-block ordered scan, candidates < 2^31.
 * ------------------------------------------------------------------------------------------ */
DH_API int dh_region_accept_dense(const double* edges, int edge_begin, int edge_end, int64_t y0, int64_t x0,
                           int64_t ny, int64_t nx, int stride, int ps, double threshold, uint8_t* mask_u8,
                           double* area_out, void* stream);
DH_API int dh_compact_coords(const uint8_t* mask_u8, int64_t y0, int64_t x0, int64_t ny, int64_t nx, int stride,
                      int32_t* coords_out, int32_t* n_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * F/G  AnnoRegionRndSampler._gen_single_proc + _patches_one_region (region_samplers.py:484-591)
 * Sampling tables (built on the host by region_samplers.AnnoRegionRndSampler, float64 exactly as
 * _calc_area_weights / _calc_weights :339-482); all pointers are device pointers:
 *   T "tables" (1 in global mode, one per image with one_image_for_batch);
 *   tbl_cls_off int32 [T+1]      -> classes usable in table t: tbl_cls[tbl_cls_off[t] .. )
 *   cat_off     int32 [T*C+1]    -> regions of (table t, class c): cat_region / cat_cdf slices
 *   cat_cdf     float64          -> inclusive cumulative weights (last = 1)
 *   img_cdf     float64 [T]      -> table (image) weights, used when T > 1
 *   reg_bbox    float64 [R][4]   -> polygon.bounds (x0, y0, x1, y1); reg_area float64 [R]
 *   reg_image   int32  [R], img_hw int32 [M][2] (layer h, w of the region's image)
 * One warp per group of k = patches_from_one_region slots (group g = local slots [g*k, g*k+k));
 * 32 attempts are evaluated in parallel per slot and the lowest accepted attempt index wins, so
 * the result equals the reference's sequential rejection loop driven by the same Philox stream
 * (counters: DESIGN.md "Philox contract"). A group whose region fails (too small, empty range,
 * miss limit) redraws class and region, up to max_redraw times (the reference's except/continue).
 * slot_offset = global index of local slot 0 (batch_index * batch_size): disjoint ranges give
 * independent streams, which is how ranks shard the sampler.
 * Outputs per slot: coords (y,x) int32, label int64 (class index), image int32, status u8.
 * ------------------------------------------------------------------------------------------ */
typedef struct dh_region_tables {
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
    int32_t n_tables;
    int32_t n_classes;
    int32_t n_regions;
    int32_t n_images;
} dh_region_tables;

DH_API int dh_region_sample(const dh_region_tables* tables_host, int64_t n_slots, int k, int ps, double threshold,
                     int miss_limit, int max_redraw, int fixed_class, int64_t slots_per_table_draw,
                     uint64_t seed, uint64_t slot_offset, int32_t* coords_out, int64_t* label_out,
                     int32_t* image_out, uint8_t* status_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * R  polygon rasterisation (north-star item (a); anno/utils.py:308-320 is the visual analogue).
 * label_out[my][mx] = 1 + index of the LAST polygon whose interior contains the centre of mask
 * pixel (my, mx) (even-odd rule, float64), 0 if none. Pixel centre = ((mx+0.5)*scale, (my+0.5)*scale).
 * ------------------------------------------------------------------------------------------ */
DH_API int dh_rasterize_polygons(const double* edges, const int32_t* edge_off, const double* reg_bbox, int n_regions,
                          double scale, int32_t* label_out, int64_t mh, int64_t mw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPHISTO_B200_H */
