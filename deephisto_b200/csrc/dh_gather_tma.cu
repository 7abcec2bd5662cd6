// K1 (TMA-engine variant): fused patch gather + normalise, warp-specialised.
//
// One producer warp stages patch rows in shared memory with asynchronous bulk copies
// (cp.async.bulk.shared.global -> SASS UBLKCP, executed by the TMA unit) behind a ring of full/empty
// mbarriers; seven consumer warps do LDS -> byte unpack -> exact /255 -> one coalesced 16-byte store per
// unit. There is no block-wide barrier in the steady state.
//
// History (profiles/r01_gather.md): the direct LDG kernel (dh_gather.cu) was latency-bound; the first TMA
// version was ISSUE-bound -- 81 warp instructions per 16-byte store, most of it per-tile bookkeeping done by
// every thread (64-bit division, coordinate loads, guarded register arrays) plus a __syncthreads per tile
// (33 % of stall samples). Here the per-tile work is done once by the producer and handed over through a
// 32-byte header in shared memory, the consumer loop is ~26 instructions per store, and stages are recycled
// through mbarriers.
//
// Alignment: a patch row starts at byte 3*x of its slide row, which is arbitrary, while bulk copies need
// 16-byte aligned global addresses and sizes (a cp.async.bulk.tensor tile whose innermost start is not a
// multiple of 16 bytes faults on B200: profiles/tma_probe.py). So each row copy starts at the 16-byte boundary
// below 3*x (a = 3*x mod 16 extra leading bytes) and the consumers realign with a funnel shift of LDS words.
// The copy never leaves the slide row: its end is roundup16(3*x + 3*ps) <= pitch when the patch is inside
// the slide. Patches that are not entirely inside the slide (zero fill) and horizontally flipped patches take a
// guarded byte path in the same kernel (NCHW output flips horizontally on the fast path: mirrored unit, reversed pixels).
//
// Concurrency is deliberately low: a 2-deep ring and at most 3 resident CTAs per SM. Deeper rings and higher occupancy were
// measured slower (profiles/r01_gather.md): the kernel is HBM-bound and more reads in flight disturb the store stream.
//
// Tile = R consecutive output rows of one patch; persistent CTAs walk tiles round-robin (no division in the loop).
// Unit = 16 bytes of output: NHWC 4 (f32) / 8 (bf16) consecutive elements of a patch row; NCHW 4 / 8 consecutive
// pixels -> one 16-byte store into each of the three channel planes.
#include <stdlib.h>

#include "dh_common.cuh"

namespace dh {

constexpr int kConsumerWarps = 7;
constexpr int kConsumers = 32 * kConsumerWarps;   // 224: divides the 1344 units of an 8-row 224-wide fp32 NHWC tile
constexpr int kTmaThreads = kConsumers + 32;      // warp 0 = producer
constexpr int kTmaMaxStages = 8;                  // ring depth is a launch parameter (p.stages <= kTmaMaxStages)
constexpr int kTmaMaxUnits = 6;                   // units per consumer thread per tile

struct TmaGatherParams {
    const uint8_t* slide;
    int64_t H, W, pitch;
    const int64_t* slides;   // multi-slide mode: device table [n_slides][4] = {data pointer, H, W, pitch}, else NULL
    const int32_t* image;    // multi-slide mode: slide index of every patch
    int n_slides;
    const int32_t* coords;
    const int32_t* out_index;
    const uint8_t* flip;
    void* out;
    int64_t B;
    int ps;
    int R;            // rows per tile
    int tiles_per_patch;
    int row_pitch;    // bytes per staged row in shared memory (multiple of 16, >= 3*ps + 32)
    int units_per_row;
    int stages;       // depth of the stage ring
    int debug;        // profiling only: 1 = no bulk loads, 2 = no stores, 4 = default-policy stores, 8 = contiguous tiles per CTA, 16 = evict-first loads
    float mean[3];
    float stdv[3];
};

// per-stage header written by the producer (plain st.shared, published by its mbarrier arrive)
struct __align__(16) TileMeta {
    long long out_off;  // element offset of the tile's first output element (plane 0 for NCHW)
    int y, x;           // patch origin
    int flags;          // DH_FLIP_H | DH_FLIP_V | kInside
    int tr;             // tile index inside the patch
    int pad[2];
};
constexpr int kInside = 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (TMA unit, SASS UBLKCP)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_load_hint(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
struct SlideRef {
    const uint8_t* data;
    int64_t H, W, pitch;
};
// the slide a patch reads from: the launch's single slide, or entry `img` of the device table (multi-slide datasets)
__device__ __forceinline__ SlideRef slide_of(const TmaGatherParams& p, int img) {
    SlideRef r;
    if (p.slides == nullptr) { r.data = p.slide; r.H = p.H; r.W = p.W; r.pitch = p.pitch; return r; }
    img = img < 0 ? 0 : (img >= p.n_slides ? p.n_slides - 1 : img);
    const longlong2* t = reinterpret_cast<const longlong2*>(p.slides + 4 * (int64_t)img);
    const longlong2 a = __ldg(t), b = __ldg(t + 1);
    r.data = reinterpret_cast<const uint8_t*>(a.x); r.H = a.y; r.W = b.x; r.pitch = b.y;
    return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// float(v) for a byte without the conversion pipe: 0x4B000000 | v is the float 8388608 + v
// (k is a compile-time constant after unrolling: one PRMT + one FADD)
// `magic` = 0x4B000000 held in a register so that the selector is the PRMT immediate
__device__ __forceinline__ float byte_f(uint32_t word, int k, uint32_t magic) {
    return __uint_as_float(__byte_perm(word, magic, 0x7440u + (uint32_t)k)) - 8388608.0f;
}
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

template <bool SCALE, bool AFFINE>
__device__ __forceinline__ float norm_f(float f, int c, const TmaGatherParams& p) {
    if (SCALE) f = div255_exact(f);
    if (AFFINE) {
        const float m = c == 0 ? p.mean[0] : (c == 1 ? p.mean[1] : p.mean[2]);
        const float s = c == 0 ? p.stdv[0] : (c == 1 ? p.stdv[1] : p.stdv[2]);
        f = __fdiv_rn(__fsub_rn(f, m), s);
    }
    return f;
}

// byte k of `word` -> normalised value, for the fast path.
// bf16 output with plain /255: ONE FMA on the magic float m = 2^23 + v (exact): fma(m, c, -2^23 * c) = RN(v * c), c = RN(1/255) -- the
// 2^23 * c term is a power-of-two multiple of c, so it cancels exactly inside the FMA. RN(v * c) differs from the IEEE quotient
// float(v) / 255 on 126 of the 256 bytes (last fp32 bit), but its bf16 rounding is the same for ALL 256 (exhaustive check in
// tests/test_oracle_cpu.py::test_bf16_div255_shortcut), so bf16 batches stay bit-identical to bf16(float32(v) / 255): 2.5 instead of
// 4.5 instructions per element on the output mode that is closest to the issue limit (2 bytes written per input byte).
template <typename OutT, bool SCALE, bool AFFINE>
__device__ __forceinline__ float conv_byte(uint32_t word, int k, uint32_t magic, int c, const TmaGatherParams& p) {
    if constexpr (SCALE && !AFFINE && sizeof(OutT) == 2) {
        const float m = __uint_as_float(__byte_perm(word, magic, 0x7440u + (uint32_t)k));
        return __fmaf_rn(m, 0x1.010102p-8f, -0x1.010102p+15f);
    } else {
        return norm_f<SCALE, AFFINE>(byte_f(word, k, magic), c, p);
    }
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// 16-byte streaming store of 4 floats / 8 bf16
__device__ __forceinline__ void store_unit(float* dst, const float (&f)[4], bool plain = false) {
    if (plain) *reinterpret_cast<float4*>(dst) = make_float4(f[0], f[1], f[2], f[3]);
    else __stcs(reinterpret_cast<float4*>(dst), make_float4(f[0], f[1], f[2], f[3]));
}
__device__ __forceinline__ void store_unit(__nv_bfloat16* dst, const float (&f)[8], bool plain = false) {
    uint4 v;
    v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]); v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
    if (plain) *reinterpret_cast<uint4*>(dst) = v;
    else __stcs(reinterpret_cast<uint4*>(dst), v);
}

template <typename OutT> struct UnitOf { static constexpr int kElems = 16 / (int)sizeof(OutT); };

// LAY: DH_NHWC, DH_NCHW, or DH_S2D16 (bf16 only): the 2x2 space-to-depth image of the patch with 16 channels per pixel and a zero
// border of 2 (top / left) and 1 (bottom / right) pixels, [ps/2 + 3][ps/2 + 3][16] -- the input of the predictor's 4x4 stem convolution
// (examples/predict_full_patched.py FusedResNetForward). Channel p * 8 + q * 3 + c of s2d pixel (y', x') is channel c of patch pixel
// (2 y' + p, 2 x' + q); channels 6, 7, 14, 15 are zero. One unit = the 8 channels of one p = 6 consecutive bytes of ONE staged row.
// The border is NOT written: the caller passes a buffer whose border is zero (it never changes between launches).
// DBG: profiling instantiation (p.debug switches; only fp32 NHWC /255 FULL is built with it). Production kernels carry none of it.
template <typename OutT, int LAY, bool SCALE, bool AFFINE, bool FULL, bool DBG = false>
__global__ void __launch_bounds__(kTmaThreads, 4) gather_tma_kernel(const TmaGatherParams p) {
    constexpr bool NCHW = LAY == DH_NCHW;
    constexpr bool S2D = LAY == DH_S2D16;
    // DH_S2D48 (bf16 only): the 4x4 space-to-depth image [ps/4][ps/4][48], channel p*12 + q*3 + c of block (Y, X) = channel c of patch
    // pixel (4Y + p, 4X + q). Same bytes as NHWC in another order; a unit (8 channels) is 8 bytes of one input row, or 4 + 4 bytes of
    // two consecutive rows (units 1 and 4 of a block's six).
    constexpr bool S4 = LAY == DH_S2D48;
    static_assert(!(S2D || S4) || sizeof(OutT) == 2, "the space-to-depth layouts are built for 2-byte outputs");
    constexpr int E = UnitOf<OutT>::kElems;          // elements (NHWC) or pixels (NCHW) per unit: 4 or 8
    constexpr int IN_BYTES = S2D ? 8 : (NCHW ? 3 * E : E);   // input bytes fetched per unit (S2D: 6 used)
    constexpr int NW = IN_BYTES / 4;                 // aligned words per unit after the funnel shift
    // FULL: every consumer thread owns exactly KU units of every tile (the ps = 224 shapes) -> no per-unit guards
    constexpr int KU = FULL ? (S2D ? 4 : (NCHW ? (E == 4 ? 4 : 2) : 6)) : kTmaMaxUnits;
    extern __shared__ __align__(128) uint8_t stages[];
    __shared__ __align__(16) TileMeta meta[kTmaMaxStages];
    __shared__ __align__(8) uint64_t full[kTmaMaxStages];
    __shared__ __align__(8) uint64_t empty[kTmaMaxStages];

    const int R = p.R, ps = p.ps, RP = p.row_pitch;
    const int row_bytes = 3 * ps;
    const int stage_bytes = R * RP;
    const int tpp = p.tiles_per_patch;
    const int S = p.stages;
    const int dbg = DBG ? p.debug : 0;
    const int64_t n_tiles = p.B * (int64_t)tpp;
    // tile assignment: round-robin over the grid (default) or one contiguous range per CTA (debug bit 8, profiling)
    const bool blocked = dbg & 8;
    const int64_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t first_tile = blocked ? (int64_t)blockIdx.x * per_cta : (int64_t)blockIdx.x;
    const int64_t tile_step = blocked ? 1 : (int64_t)gridDim.x;
    const int64_t my_tiles = blocked ? (first_tile >= n_tiles ? 0 : (n_tiles - first_tile < per_cta ? n_tiles - first_tile : per_cta))
                                     : (first_tile < n_tiles ? (n_tiles - first_tile + gridDim.x - 1) / gridDim.x : 0);
    const int64_t plane = (int64_t)ps * ps;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x < 32) {
        // ---------------- producer warp ----------------
        const int lane = threadIdx.x;
        int64_t patch = first_tile / tpp;                      // once; the loop advances (patch, tr) incrementally
        int tr = (int)(first_tile - patch * tpp);
        const int64_t dq = tile_step / tpp;
        const int dr = (int)(tile_step - dq * tpp);
        uint64_t policy = 0;
        if (dbg & 16) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        const uint32_t stage0 = smem_u32(stages);
        int s = 0;
        uint32_t ph = 1;  // a fresh mbarrier passes a wait on parity 1: the first S tiles do not wait
        for (int64_t i = 0; i < my_tiles; ++i) {
            mbar_wait(&empty[s], ph);
            const int y = __ldg(p.coords + 2 * patch);
            const int x = __ldg(p.coords + 2 * patch + 1);
            const uint32_t fl = p.flip ? (uint32_t)__ldg(p.flip + patch) : 0u;
            const int img = p.image ? __ldg(p.image + patch) : 0;
            const SlideRef sl = slide_of(p, img);
            const bool inside = (y >= 0) && (x >= 0) && ((int64_t)y + ps <= sl.H) && ((int64_t)x + ps <= sl.W);
            const int a = (3 * x) & 15;
            const uint32_t bytes = (uint32_t)((a + row_bytes + 15) & ~15);
            const bool stage_it = inside && !(dbg & 1);
            if (lane == 0) {
                const int64_t slot = p.out_index ? (int64_t)__ldg(p.out_index + patch) : patch;
                TileMeta m;
                if (S2D) {
                    const int64_t PH = ps / 2 + 3;
                    m.out_off = slot * PH * PH * 16 + ((2 + (int64_t)tr * (R / 2)) * PH + 2) * 16;
                } else {
                    m.out_off = slot * 3 * plane + (int64_t)tr * R * (NCHW ? ps : row_bytes);
                }
                m.y = y; m.x = x; m.flags = (int)fl | (inside ? kInside : 0); m.tr = tr; m.pad[0] = img; m.pad[1] = 0;
                meta[s] = m;
                if (stage_it) mbar_expect_tx(&full[s], bytes * (uint32_t)R);
                else mbar_arrive(&full[s]);  // guarded path reads global memory directly: nothing to stage
            }
            __syncwarp();
            if (stage_it && lane < R) {
                const int orow = tr * R + lane;                                // output row of the patch
                const int srow = (fl & DH_FLIP_V) ? ps - 1 - orow : orow;      // source row
                const uint8_t* src = sl.data + (int64_t)(y + srow) * sl.pitch + ((3 * (int64_t)x) & ~(int64_t)15);
                if (dbg & 16) bulk_load_hint(stage0 + (uint32_t)(s * stage_bytes + lane * RP), src, bytes, &full[s], policy);
                else bulk_load(stage0 + (uint32_t)(s * stage_bytes + lane * RP), src, bytes, &full[s]);
            }
            tr += dr; patch += dq;
            if (tr >= tpp) { tr -= tpp; ++patch; }
            if (++s == S) { s = 0; ph ^= 1u; }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int tid = threadIdx.x - 32;
    const int upr = p.units_per_row;             // S2D: units per space-to-depth row = 2 halves x ps/2 pixels
    const int units_per_tile = (S4 ? R / 4 : (S2D ? R / 2 : R)) * upr;   // S4: upr = units per block row = 6 * ps/4
    // per-thread unit geometry is the same for every tile: unit u_k = tid + k * kConsumers
    uint32_t s_off[kTmaMaxUnits];
    uint32_t s_mirror[kTmaMaxUnits];   // s_off of the mirrored unit of the same row = s_mirror - s_off (NCHW horizontal flip)
    int cph[kTmaMaxUnits];             // S2D: element offset of the unit inside the tile's output
    int nu = 0;
#pragma unroll
    for (int k = 0; k < kTmaMaxUnits; ++k) {
        const int u = tid + k * kConsumers;
        const int r = u / upr, c = u - r * upr;
        if (S2D) {
            s_off[k] = (uint32_t)((2 * r + (c & 1)) * RP + 6 * (c >> 1));      // staged row 2r + p, byte 6 x' (not word aligned for odd x')
            s_mirror[k] = 0;
            cph[k] = (r * (ps / 2 + 3) + (c >> 1)) * 16 + 8 * (c & 1);
        } else if (S4) {
            const int X = c / 6, j = c - 6 * X;                               // block column, unit of the block's six
            const int row_a = 4 * r + (2 * j) / 3, byte_a = (8 * j) % 12;     // first staged row of the unit and its byte inside the block's 12
            s_off[k] = (uint32_t)(row_a * RP + 12 * X + byte_a);              // word aligned: 12 X + {0, 4, 8}
            // where the second half of a two-row unit starts (one-row units re-read their own words: the load is unconditional, and
            // row_a + 1 of a block's last row would lie behind the stage)
            s_mirror[k] = byte_a == 8 ? (uint32_t)((row_a + 1) * RP + 12 * X) : s_off[k];
            cph[k] = ((8 * j) % 3) | (byte_a == 8 ? 4 : 0);                   // channel of the unit's first byte | two-row flag
        } else {
            s_off[k] = (uint32_t)(r * RP + IN_BYTES * c);
            s_mirror[k] = (uint32_t)(2 * r * RP + IN_BYTES * (upr - 1));
            cph[k] = (E * c) % 3;                    // channel of the unit's first element (NHWC, AFFINE only)
        }
        if (u < units_per_tile) nu = k + 1;
    }
    const uint32_t stage0 = smem_u32(stages);
    const uint32_t magic = opaque_u32(0x4B000000u);
    OutT* const out_base = reinterpret_cast<OutT*>(p.out) + (S2D ? 0 : (int64_t)E * tid);

    int s = 0;
    uint32_t ph = 0;
    for (int64_t i = 0; i < my_tiles; ++i) {
        mbar_wait(&full[s], ph);
        const TileMeta m = meta[s];
        const uint32_t sbase = stage0 + (uint32_t)(s * stage_bytes);
        OutT* const o = out_base + m.out_off;
        if ((m.flags & kInside) && (NCHW || !(m.flags & DH_FLIP_H))) {
          if constexpr (S4) {
            const int a = (3 * m.x) & 15;
            const uint32_t sh = (uint32_t)(a & 3) * 8u;
            const uint32_t abase = sbase + (uint32_t)(a & ~3);
#pragma unroll
            for (int k = 0; k < KU; ++k) {
                if (FULL || k < nu) {
                    const uint32_t q0 = lds32(abase + s_off[k]), q1 = lds32(abase + s_off[k] + 4), q2 = lds32(abase + s_off[k] + 8);
                    const uint32_t r0 = lds32(abase + s_mirror[k]), r1 = lds32(abase + s_mirror[k] + 4);
                    const uint32_t w0 = __funnelshift_r(q0, q1, sh);
                    const uint32_t w1 = (cph[k] & 4) ? __funnelshift_r(r0, r1, sh) : __funnelshift_r(q1, q2, sh);
                    float f[8];
                    int c = cph[k] & 3;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        f[j] = conv_byte<OutT, SCALE, AFFINE>(j < 4 ? w0 : w1, j & 3, magic, c, p);
                        if (AFFINE) c = c == 2 ? 0 : c + 1;
                    }
                    store_unit(o + k * (kConsumers * E), f);
                }
            }
          } else if constexpr (S2D) {
            const int a = (3 * m.x) & 15;
#pragma unroll
            for (int k = 0; k < KU; ++k) {
                if (FULL || k < nu) {
                    const uint32_t t = (uint32_t)(a & 3) + (s_off[k] & 3u);    // byte phase of this unit relative to two word-aligned bases
                    const uint32_t addr = sbase + (uint32_t)(a & ~3) + (s_off[k] & ~3u) + (t & ~3u);
                    const uint32_t sh = (t & 3u) * 8u;
                    const uint32_t q0 = lds32(addr), q1 = lds32(addr + 4), q2 = lds32(addr + 8);
                    const uint32_t w0 = __funnelshift_r(q0, q1, sh), w1 = __funnelshift_r(q1, q2, sh);
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 6; ++j) f[j] = conv_byte<OutT, SCALE, AFFINE>(j < 4 ? w0 : w1, j & 3, magic, j % 3, p);
                    f[6] = 0.f; f[7] = 0.f;
                    store_unit(o + cph[k], f);
                }
            }
          } else {
            // fast path: aligned LDS words + funnel shift. A horizontal flip in NCHW mode reads the mirrored unit of the row and
            // reverses the pixel order inside the unit (compile-time byte permutation); in NHWC mode it takes the byte path.
            const int a = (3 * m.x) & 15;
            const uint32_t sh = (uint32_t)(a & 3) * 8u;
            const uint32_t abase = sbase + (uint32_t)(a & ~3);
            const bool fh = NCHW && (m.flags & DH_FLIP_H);
#define DH_STORE(ptr, vals) do { if (!(dbg & 2) || (vals)[0] == 12345.f) store_unit(ptr, vals, dbg & 4); } while (0)
#pragma unroll
            for (int k = 0; k < KU; ++k) {
                if (FULL || k < nu) {
                    uint32_t q[NW + 1], w[NW];
                    const uint32_t so = fh ? s_mirror[k] - s_off[k] : s_off[k];
#pragma unroll
                    for (int j = 0; j <= NW; ++j) q[j] = lds32(abase + so + 4 * j);
#pragma unroll
                    for (int j = 0; j < NW; ++j) w[j] = __funnelshift_r(q[j], q[j + 1], sh);
                    OutT* const ok = o + k * (kConsumers * E);
                    if (!NCHW) {
                        float f[E];
                        int c = cph[k];
#pragma unroll
                        for (int j = 0; j < E; ++j) {
                            f[j] = conv_byte<OutT, SCALE, AFFINE>(w[j >> 2], j & 3, magic, c, p);
                            if (AFFINE) c = c == 2 ? 0 : c + 1;
                        }
                        DH_STORE(ok, f);
                    } else if (!fh) {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            float f[E];
#pragma unroll
                            for (int j = 0; j < E; ++j) {
                                const int b = 3 * j + ch;  // byte of pixel j, channel ch
                                f[j] = conv_byte<OutT, SCALE, AFFINE>(w[b >> 2], b & 3, magic, ch, p);
                            }
                            DH_STORE(ok + ch * plane, f);
                        }
                    } else {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            float f[E];
#pragma unroll
                            for (int j = 0; j < E; ++j) {
                                const int b = 3 * (E - 1 - j) + ch;  // output pixel j = source pixel E-1-j of the mirrored unit
                                f[j] = conv_byte<OutT, SCALE, AFFINE>(w[b >> 2], b & 3, magic, ch, p);
                            }
                            DH_STORE(ok + ch * plane, f);
                        }
                    }
                }
            }
#undef DH_STORE
          }
        } else {
            // slow path: horizontally flipped patches (bytes from the stage) and patches that overhang the slide (guarded global loads)
            const bool inside = m.flags & kInside, fv = m.flags & DH_FLIP_V, fh = m.flags & DH_FLIP_H;
            const int a = (3 * m.x) & 15;
            const SlideRef sl = slide_of(p, m.pad[0]);
            for (int u = tid; u < units_per_tile; u += kConsumers) {
                const int ur0 = u / upr, uc = u - ur0 * upr;
                const int ur = S2D ? 2 * ur0 + (uc & 1) : ur0;              // staged row of this unit (S4: per element, below)
                const int orow = m.tr * R + ur;
                const int srow = fv ? ps - 1 - orow : orow;
                auto pix = [&](int scol, int ch) -> float {
                    uint32_t v = 0;
                    if (inside) {
                        v = stages[(size_t)s * stage_bytes + ur * RP + a + 3 * scol + ch];
                    } else {
                        const int64_t yy = (int64_t)m.y + srow, xx = (int64_t)m.x + scol;
                        if (yy >= 0 && yy < sl.H && xx >= 0 && xx < sl.W) v = __ldg(sl.data + yy * sl.pitch + 3 * xx + ch);
                    }
                    return norm_f<SCALE, AFFINE>((float)v, ch, p);
                };
                OutT* const ou = reinterpret_cast<OutT*>(p.out) + m.out_off + (int64_t)E * u;
                if constexpr (S4) {
                    const int X = uc / 6, j6 = uc - 6 * X;
                    float f[E];
#pragma unroll
                    for (int j = 0; j < E; ++j) {
                        const int ch48 = 8 * j6 + j, pr = ch48 / 12, rem = ch48 - 12 * pr, q = rem / 3, ch = rem - 3 * q;
                        const int srow_u = 4 * ur0 + pr;                                   // staged (= output) row inside the tile
                        const int orow_u = m.tr * R + srow_u;
                        const int src_row = fv ? ps - 1 - orow_u : orow_u;
                        const int col = 4 * X + q, scol = fh ? ps - 1 - col : col;
                        uint32_t v = 0;
                        if (inside) {
                            v = stages[(size_t)s * stage_bytes + srow_u * RP + a + 3 * scol + ch];
                        } else {
                            const int64_t yy = (int64_t)m.y + src_row, xx = (int64_t)m.x + scol;
                            if (yy >= 0 && yy < sl.H && xx >= 0 && xx < sl.W) v = __ldg(sl.data + yy * sl.pitch + 3 * xx + ch);
                        }
                        f[j] = norm_f<SCALE, AFFINE>((float)v, ch, p);
                    }
                    store_unit(ou, f);
                } else if constexpr (S2D) {
                    float f[E];
#pragma unroll
                    for (int j = 0; j < E; ++j) {
                        const int col = 2 * (uc >> 1) + j / 3;
                        f[j] = j < 6 ? pix(fh ? ps - 1 - col : col, j % 3) : 0.f;
                    }
                    store_unit(reinterpret_cast<OutT*>(p.out) + m.out_off + ((int64_t)ur0 * (ps / 2 + 3) + (uc >> 1)) * 16 + 8 * (uc & 1), f);
                } else if (!NCHW) {
                    float f[E];
#pragma unroll
                    for (int j = 0; j < E; ++j) {
                        const int e = E * uc + j, col = e / 3, ch = e - 3 * col;
                        f[j] = pix(fh ? ps - 1 - col : col, ch);
                    }
                    store_unit(ou, f);
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        float f[E];
#pragma unroll
                        for (int j = 0; j < E; ++j) f[j] = pix(fh ? ps - 1 - (E * uc + j) : E * uc + j, ch);
                        store_unit(ou + ch * plane, f);
                    }
                }
            }
        }
        __syncwarp();  // all lanes of this warp are done reading stage s
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
        if (++s == S) { s = 0; ph ^= 1u; }
    }
}

// persistent grid = SMs x resident CTAs of this instantiation, so every CTA is co-resident
template <typename K>
static int launch_one(K kernel, const TmaGatherParams& p, int64_t n_tiles, size_t smem, cudaStream_t st, int occ_override, int occ_cap = 3) {
    static int occ_cache[64] = {0};  // per instantiation, indexed by shared-memory size in KB
    int& occ = occ_cache[(smem >> 10) & 63];
    if (occ == 0) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kTmaThreads, smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
        if (occ < 1) occ = 1;
        // measured (profiles/r01_gather.md): 3 resident CTAs per SM beat 4 (and 5, 6) -- less concurrency, less HBM read/write interference;
        // bf16 NHWC is best with 2 (0.846 vs 0.808 of the measured peak; the other bf16 layouts lose 10-15 % at 2: profiles/r02_gather.md)
        if (occ > occ_cap) occ = occ_cap;
    }
    const int64_t want = (int64_t)kNumSMs * (occ_override > 0 ? occ_override : occ);   // profiling: DH_GATHER_OCC resident CTAs per SM
    const int grid = (int)(n_tiles < want ? n_tiles : want);
    kernel<<<grid, kTmaThreads, smem, st>>>(p);
    return DH_OK;
}

// Returns DH_ERR_UNSUPPORTED (without touching the error string) when the shape does not fit this kernel.
int gather_tma_launch(const uint8_t* slide, int64_t H, int64_t W, int64_t pitch, const int64_t* slides_dev, int n_slides,
                      const int32_t* image, const int32_t* coords, const int32_t* out_index,
                      int64_t B, int ps, void* out, int out_dtype, int out_layout, int scale255, const float* mean3, const float* std3,
                      const uint8_t* flip, int debug, cudaStream_t st) {
    const bool nchw = out_layout == DH_NCHW;
    const bool s2d = out_layout == DH_S2D16, s4 = out_layout == DH_S2D48;
    if (out_dtype == DH_U8) return DH_ERR_UNSUPPORTED;
    if (s2d && (out_dtype != DH_BF16 || ps % 2 != 0)) return DH_ERR_UNSUPPORTED;
    if (s4 && (out_dtype != DH_BF16 || ps % 4 != 0)) return DH_ERR_UNSUPPORTED;
    const int E = out_dtype == DH_F32 ? 4 : 8;
    if ((!s2d && !s4 && ps % E != 0) || pitch % 16 != 0 || reinterpret_cast<uintptr_t>(slide) % 16 != 0) return DH_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(out) % 16 != 0) return DH_ERR_UNSUPPORTED;
    const int row_bytes = 3 * ps;
    if (!nchw && !s2d && !s4 && row_bytes % E != 0) return DH_ERR_UNSUPPORTED;
    const int row_pitch = ((15 + row_bytes + 15) & ~15) + 16;  // largest copy + one spare 16-byte line for the funnel shift
    // units per (output) row: s2d per space-to-depth row (two half pixels per pixel), s4 per block row (six units per 4x4 block)
    const int units_per_row = s4 ? (ps / 4) * 6 : (s2d ? ps : (nchw ? ps / E : row_bytes / E));
    const int rows_per_out = s4 ? 4 : (s2d ? 2 : 1);
    int R = 0;
    const char* env_rows = getenv("DH_GATHER_ROWS");  // profiling override: rows per tile (stage of up to 24 KB instead of 12 KB)
    const int want_rows = env_rows ? atoi(env_rows) : 0;
    for (int r = 32; r >= 1; --r) {
        if (r % rows_per_out) continue;                                     // a space-to-depth row needs all of its input rows in the tile
        if (want_rows > 0 && r != want_rows) continue;
        const int tile_units = (r / rows_per_out) * units_per_row;
        if (ps % r == 0 && tile_units <= kConsumers * kTmaMaxUnits && r * row_pitch <= (want_rows > 0 ? 24 : 12) * 1024) { R = r; break; }
    }
    if (!R) return DH_ERR_UNSUPPORTED;
    if (B * (int64_t)(ps / R) >= (1ll << 40)) return DH_ERR_UNSUPPORTED;

    TmaGatherParams p{};
    p.slide = slide; p.H = H; p.W = W; p.pitch = pitch;
    p.slides = slides_dev; p.image = slides_dev ? image : nullptr; p.n_slides = n_slides;
    p.coords = coords; p.out_index = out_index; p.flip = flip; p.out = out; p.B = B; p.ps = ps; p.R = R;
    p.tiles_per_patch = ps / R; p.row_pitch = row_pitch; p.units_per_row = units_per_row; p.debug = debug;
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean3 ? mean3[c] : 0.f; p.stdv[c] = std3 ? std3[c] : 1.f; }
    const bool affine = mean3 != nullptr;
    const char* env = getenv("DH_GATHER_STAGES");  // profiling override
    const int env_stages = env ? atoi(env) : 0;
    // measured (profiles/r01_gather.md, r02_gather.md): a shallow ring and few resident CTAs interfere least with the store stream; the optimum
    // is sharp and layout dependent -- fp32 NHWC: 3 stages x 2 CTAs per SM (0.862 / 0.873 of the measured peak at 5 120 / 8 192 patches per launch
    // vs 0.852 / 0.861 with 2 x 3), bf16 NHWC: 2 x 2, every other mode: 2 x 3
    const bool f32_nhwc = out_dtype == DH_F32 && !nchw && !s2d && !s4;
    int stages = env_stages >= 2 && env_stages <= kTmaMaxStages ? env_stages : (f32_nhwc ? 3 : 2);
    while (stages > 2 && (size_t)stages * R * row_pitch > 56 * 1024) --stages;
    p.stages = stages;
    const size_t smem = (size_t)stages * R * row_pitch;
    const int64_t n_tiles = B * (int64_t)p.tiles_per_patch;
    int rc_launch = DH_OK;
    // every consumer thread owns exactly KU units (see the kernel): true for ps = 224 in all output modes
    const bool full_units = (R / rows_per_out) * units_per_row == kConsumers * (s2d ? 4 : (nchw ? (E == 4 ? 4 : 2) : 6));
    const char* env_occ = getenv("DH_GATHER_OCC");  // profiling override of the resident CTAs per SM (grid size), 0 = default
    const int occ_o = env_occ ? atoi(env_occ) : 0;
#define DH_TMA(T, N, S, A) rc_launch = full_units ? launch_one(gather_tma_kernel<T, N, S, A, true>, p, n_tiles, smem, st, occ_o, (N) == DH_NHWC ? 2 : 3) \
                                            : launch_one(gather_tma_kernel<T, N, S, A, false>, p, n_tiles, smem, st, occ_o, (N) == DH_NHWC ? 2 : 3)
#define DH_TMA_SA(T, N)                                                                  \
    do {                                                                                 \
        if (scale255) { if (affine) DH_TMA(T, N, true, true); else DH_TMA(T, N, true, false); } \
        else          { if (affine) DH_TMA(T, N, false, true); else DH_TMA(T, N, false, false); } \
    } while (0)
    if (debug) {  // profiling switches exist for one instantiation only
        if (!(out_dtype == DH_F32 && !nchw && scale255 && !affine && full_units)) return DH_ERR_UNSUPPORTED;
        rc_launch = launch_one(gather_tma_kernel<float, DH_NHWC, true, false, true, true>, p, n_tiles, smem, st, occ_o);
    } else if (s2d)          { DH_TMA_SA(__nv_bfloat16, DH_S2D16); }
    else if (s4)             { DH_TMA_SA(__nv_bfloat16, DH_S2D48); }
    else if (out_dtype == DH_F32) { if (nchw) DH_TMA_SA(float, DH_NCHW); else DH_TMA_SA(float, DH_NHWC); }
    else                     { if (nchw) DH_TMA_SA(__nv_bfloat16, DH_NCHW); else DH_TMA_SA(__nv_bfloat16, DH_NHWC); }
#undef DH_TMA_SA
#undef DH_TMA
    if (rc_launch != DH_OK) return rc_launch;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return cuda_fail(e, "gather_tma_kernel");
    return DH_OK;
}

}  // namespace dh
