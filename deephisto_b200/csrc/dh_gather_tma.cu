// K1 (TMA-engine variant): fused patch gather + normalise with the patch rows staged through shared memory by
// asynchronous bulk copies (cp.async.bulk.shared.global -> SASS UBLKCP, executed by the TMA unit) behind an
// mbarrier ring, S tiles ahead of the consumers.
//
// Why: the direct LDG kernel (dh_gather.cu) is latency-bound -- ncu shows ~70 % of its stall samples on the
// funnel-shift that consumes the two global loads (profiles/r01_gather.md). Here the loads are issued by one thread,
// complete asynchronously, and the SM only does LDS -> convert -> one coalesced 16-byte store per 4 outputs.
//
// Alignment: a patch row starts at byte 3*x of its slide row, which is arbitrary, while bulk copies (and TMA tensor
// tiles: measured on B200, a cp.async.bulk.tensor whose innermost start byte is not a multiple of 16 faults with
// "illegal instruction") need 16-byte aligned global addresses. So each row copy starts at the 16-byte boundary below
// 3*x (a = 3*x mod 16 extra leading bytes) and the consumers realign with a funnel shift of two LDS words.
// The copy never leaves the slide row: its end is roundup16(3*x + 3*ps) <= pitch when the patch is inside the slide.
// Patches that are not entirely inside the slide (zero fill) take a guarded global-load path in the same kernel.
//
// Tile = R consecutive output rows of one patch; persistent CTAs walk tiles round-robin.
#include "dh_common.cuh"

namespace dh {

constexpr int kTmaThreads = 256;
constexpr int kTmaStages = 4;
constexpr int kTmaMaxUnits = 8;  // units per thread per tile

struct TmaGatherParams {
    const uint8_t* slide;
    int64_t H, W, pitch;
    const int32_t* coords;
    const int32_t* out_index;
    const uint8_t* flip;
    void* out;
    int64_t B;
    int ps;
    int R;            // rows per tile
    int tiles_per_patch;
    int row_pitch;    // bytes per staged row in shared memory (multiple of 16, >= 3*ps + 16)
    int affine;
    float mean[3];
    float stdv[3];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// float(v) for a byte without the XU pipe: 0x4B000000 | v is the float 8388608 + v
__device__ __forceinline__ float byte_to_float(uint32_t word, uint32_t sel) {
    // sel = 0x744k: byte k of `word` in the low byte, 0x00 0x00 0x4B above it
    return __uint_as_float(__byte_perm(word, 0x4B000000u, sel)) - 8388608.0f;
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

template <typename OutT>
__device__ __forceinline__ void store4(OutT* dst, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* dst, float a, float b, float c, float d) {
    __stcs(reinterpret_cast<float4*>(dst), make_float4(a, b, c, d));
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* dst, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&lo);
    v.y = *reinterpret_cast<uint32_t*>(&hi);
    __stcs(reinterpret_cast<uint2*>(dst), v);
}

struct TileInfo {
    int64_t patch, slot;
    int tr, y, x;
    uint32_t fl;
    bool inside;
};

__device__ __forceinline__ TileInfo tile_info(const TmaGatherParams& p, int64_t t) {
    TileInfo ti;
    ti.patch = t / p.tiles_per_patch;
    ti.tr = (int)(t - ti.patch * p.tiles_per_patch);
    ti.y = __ldg(p.coords + 2 * ti.patch);
    ti.x = __ldg(p.coords + 2 * ti.patch + 1);
    ti.slot = p.out_index ? (int64_t)__ldg(p.out_index + ti.patch) : ti.patch;
    ti.fl = p.flip ? (uint32_t)__ldg(p.flip + ti.patch) : 0u;
    ti.inside = (ti.y >= 0) && (ti.x >= 0) && ((int64_t)ti.y + p.ps <= p.H) && ((int64_t)ti.x + p.ps <= p.W);
    return ti;
}

template <typename OutT, bool NCHW, bool SCALE, bool AFFINE>
__global__ void __launch_bounds__(kTmaThreads) gather_tma_kernel(const TmaGatherParams p) {
    extern __shared__ __align__(128) uint8_t stages[];
    __shared__ __align__(8) uint64_t full[kTmaStages];
    const int R = p.R, ps = p.ps, RP = p.row_pitch;
    const int row_bytes = 3 * ps;
    const int stage_bytes = R * RP;

    const int64_t n_tiles = p.B * (int64_t)p.tiles_per_patch;
    const int64_t first = blockIdx.x;
    const int64_t stride = gridDim.x;
    const int64_t my_tiles = first < n_tiles ? (n_tiles - first + stride - 1) / stride : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTmaStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // warp 0 issues the R row copies of my i-th tile into stage i % S (lane r copies row r; lane 0 arms the barrier first)
    auto issue = [&](int64_t i) {
        const TileInfo ti = tile_info(p, first + i * stride);
        const int s = (int)(i % kTmaStages);
        if (!ti.inside) {  // guarded path reads global memory directly: nothing to stage, but the phase must still complete
            if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
            return;
        }
        const int a = (3 * ti.x) & 15;
        const uint32_t bytes = (uint32_t)((a + row_bytes + 15) & ~15);
        const int lane = threadIdx.x;
        if (lane == 0) mbar_expect_tx(&full[s], bytes * (uint32_t)R);
        __syncwarp();
        if (lane < R) {
            const int orow = ti.tr * R + lane;                                   // output row of the patch
            const int srow = (ti.fl & DH_FLIP_V) ? ps - 1 - orow : orow;         // source row
            const uint8_t* src = p.slide + (int64_t)(ti.y + srow) * p.pitch + ((3 * (int64_t)ti.x) & ~(int64_t)15);
            bulk_load(stages + (size_t)s * stage_bytes + (size_t)lane * RP, src, bytes, &full[s]);
        }
    };

    if (threadIdx.x < 32) {
        for (int64_t i = 0; i < kTmaStages - 1 && i < my_tiles; ++i) issue(i);
    }

    // per-thread unit geometry is the same for every tile
    const int units_per_row = NCHW ? ps / 4 : row_bytes / 4;
    const int units_per_tile = R * units_per_row;
    int u_row[kTmaMaxUnits], u_col[kTmaMaxUnits], s_off[kTmaMaxUnits], o_off[kTmaMaxUnits];
#pragma unroll
    for (int k = 0; k < kTmaMaxUnits; ++k) {
        int u = threadIdx.x + k * kTmaThreads;
        int r = u / units_per_row;
        u_row[k] = u < units_per_tile ? r : -1;
        u_col[k] = u - r * units_per_row;                       // unit index inside the row
        s_off[k] = r * RP + (NCHW ? 12 : 4) * u_col[k];         // byte offset of the unit in the stage (before the +a shift)
        o_off[k] = NCHW ? r * ps + 4 * u_col[k] : r * row_bytes + 4 * u_col[k];
    }

    const int64_t plane = (int64_t)ps * ps;
    for (int64_t i = 0; i < my_tiles; ++i) {
        if (threadIdx.x < 32 && i + kTmaStages - 1 < my_tiles) issue(i + kTmaStages - 1);
        const TileInfo ti = tile_info(p, first + i * stride);
        const bool fv = ti.fl & DH_FLIP_V, fh = ti.fl & DH_FLIP_H;
        const int s = (int)(i % kTmaStages);
        const uint8_t* src = stages + (size_t)s * stage_bytes;
        const int a = (3 * ti.x) & 15;
        const uint32_t sh = (uint32_t)(a & 3) * 8u;
        const int a4 = a & ~3;
        mbar_wait(&full[s], (uint32_t)((i / kTmaStages) & 1));

        OutT* outp = reinterpret_cast<OutT*>(p.out) + ti.slot * 3 * plane + (int64_t)ti.tr * R * (NCHW ? ps : row_bytes);
        if (ti.inside && !fh) {
            // fast path: aligned LDS words + funnel shift
#pragma unroll
            for (int k = 0; k < kTmaMaxUnits; ++k) {
                if (u_row[k] >= 0) {
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(src + s_off[k] + a4);
                    if (!NCHW) {
                        const uint32_t w = __funnelshift_r(wp[0], wp[1], sh);
                        const int c0 = (4 * u_col[k]) % 3, c1 = c0 == 2 ? 0 : c0 + 1, c2 = c1 == 2 ? 0 : c1 + 1;
                        float f0 = norm_f<SCALE, AFFINE>(byte_to_float(w, 0x7440), c0, p);
                        float f1 = norm_f<SCALE, AFFINE>(byte_to_float(w, 0x7441), c1, p);
                        float f2 = norm_f<SCALE, AFFINE>(byte_to_float(w, 0x7442), c2, p);
                        float f3 = norm_f<SCALE, AFFINE>(byte_to_float(w, 0x7443), c0, p);
                        store4<OutT>(outp + o_off[k], f0, f1, f2, f3);
                    } else {
                        const uint32_t q0 = wp[0], q1 = wp[1], q2 = wp[2], q3 = wp[3];
                        const uint32_t w0 = __funnelshift_r(q0, q1, sh), w1 = __funnelshift_r(q1, q2, sh), w2 = __funnelshift_r(q2, q3, sh);
                        // bytes: w0 = R0 G0 B0 R1 | w1 = G1 B1 R2 G2 | w2 = B2 R3 G3 B3
                        OutT* o = outp + o_off[k];
                        store4<OutT>(o, norm_f<SCALE, AFFINE>(byte_to_float(w0, 0x7440), 0, p), norm_f<SCALE, AFFINE>(byte_to_float(w0, 0x7443), 0, p),
                                     norm_f<SCALE, AFFINE>(byte_to_float(w1, 0x7442), 0, p), norm_f<SCALE, AFFINE>(byte_to_float(w2, 0x7441), 0, p));
                        store4<OutT>(o + plane, norm_f<SCALE, AFFINE>(byte_to_float(w0, 0x7441), 1, p), norm_f<SCALE, AFFINE>(byte_to_float(w1, 0x7440), 1, p),
                                     norm_f<SCALE, AFFINE>(byte_to_float(w1, 0x7443), 1, p), norm_f<SCALE, AFFINE>(byte_to_float(w2, 0x7442), 1, p));
                        store4<OutT>(o + 2 * plane, norm_f<SCALE, AFFINE>(byte_to_float(w0, 0x7442), 2, p), norm_f<SCALE, AFFINE>(byte_to_float(w1, 0x7441), 2, p),
                                     norm_f<SCALE, AFFINE>(byte_to_float(w2, 0x7440), 2, p), norm_f<SCALE, AFFINE>(byte_to_float(w2, 0x7443), 2, p));
                    }
                }
            }
        } else {
            // slow path: horizontally flipped patches (bytes from the stage) and patches that overhang the slide (guarded global loads)
            // (recomputes the unit geometry from the unit index so that the register arrays above are never indexed dynamically)
            for (int u = threadIdx.x; u < units_per_tile; u += kTmaThreads) {
                const int ur = u / units_per_row, uc = u - ur * units_per_row;
                const int orow = ti.tr * R + ur;
                const int srow = fv ? ps - 1 - orow : orow;
                auto pix = [&](int scol, int ch) -> float {
                    uint32_t v = 0;
                    if (ti.inside) {
                        v = src[ur * RP + a + 3 * scol + ch];
                    } else {
                        const int64_t yy = (int64_t)ti.y + srow, xx = (int64_t)ti.x + scol;
                        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) v = __ldg(p.slide + yy * p.pitch + 3 * xx + ch);
                    }
                    return norm_f<SCALE, AFFINE>((float)v, ch, p);
                };
                if (!NCHW) {
                    float f[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int e = 4 * uc + j, col = e / 3, ch = e - 3 * col;
                        f[j] = pix(fh ? ps - 1 - col : col, ch);
                    }
                    store4<OutT>(outp + ur * row_bytes + 4 * uc, f[0], f[1], f[2], f[3]);
                } else {
                    const int col = 4 * uc;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float f[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) f[j] = pix(fh ? ps - 1 - (col + j) : col + j, c);
                        store4<OutT>(outp + ur * ps + col + c * plane, f[0], f[1], f[2], f[3]);
                    }
                }
            }
        }
        __syncthreads();  // every thread is done with stage s before warp 0 refills it (next iteration)
    }
}

// persistent grid = SMs x resident CTAs of this instantiation (registers decide: 3 or 4 per SM), so every CTA is co-resident
template <typename K>
static int launch_one(K kernel, const TmaGatherParams& p, int64_t n_tiles, size_t smem, cudaStream_t st) {
    static int occ_cache[64] = {0};  // per instantiation, indexed by shared-memory size in KB
    int& occ = occ_cache[(smem >> 10) & 63];
    if (occ == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kTmaThreads, smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
        if (occ < 1) occ = 1;
    }
    const int64_t want = (int64_t)kNumSMs * occ;
    const int grid = (int)(n_tiles < want ? n_tiles : want);
    kernel<<<grid, kTmaThreads, smem, st>>>(p);
    return DH_OK;
}

// Returns DH_ERR_UNSUPPORTED (without touching the error string) when the shape does not fit this kernel.
int gather_tma_launch(const uint8_t* slide, int64_t H, int64_t W, int64_t pitch, const int32_t* coords, const int32_t* out_index,
                      int64_t B, int ps, void* out, int out_dtype, int out_layout, int scale255, const float* mean3, const float* std3,
                      const uint8_t* flip, cudaStream_t st) {
    const bool nchw = out_layout == DH_NCHW;
    if (out_dtype == DH_U8) return DH_ERR_UNSUPPORTED;
    if (ps % 4 != 0 || pitch % 16 != 0 || reinterpret_cast<uintptr_t>(slide) % 16 != 0) return DH_ERR_UNSUPPORTED;
    const size_t esz = out_dtype == DH_F32 ? 4 : 2;
    if (reinterpret_cast<uintptr_t>(out) % (4 * esz) != 0) return DH_ERR_UNSUPPORTED;
    const int row_bytes = 3 * ps;
    const int row_pitch = ((15 + row_bytes + 15) & ~15) + 16;  // largest copy + one spare word for the funnel shift
    const int units_per_row = nchw ? ps / 4 : row_bytes / 4;
    int R = 0;
    for (int r = 32; r >= 1; --r)
        if (ps % r == 0 && r * units_per_row <= kTmaThreads * kTmaMaxUnits && r * row_pitch <= 11 * 1024) { R = r; break; }
    if (!R) return DH_ERR_UNSUPPORTED;

    TmaGatherParams p{};
    p.slide = slide; p.H = H; p.W = W; p.pitch = pitch;
    p.coords = coords; p.out_index = out_index; p.flip = flip; p.out = out; p.B = B; p.ps = ps; p.R = R;
    p.tiles_per_patch = ps / R; p.row_pitch = row_pitch; p.affine = mean3 ? 1 : 0;
    for (int c = 0; c < 3; ++c) { p.mean[c] = mean3 ? mean3[c] : 0.f; p.stdv[c] = std3 ? std3[c] : 1.f; }
    const size_t smem = (size_t)kTmaStages * R * row_pitch;
    const int64_t n_tiles = B * (int64_t)p.tiles_per_patch;
    int rc_launch = DH_OK;
#define DH_TMA(T, N, S, A) rc_launch = launch_one(gather_tma_kernel<T, N, S, A>, p, n_tiles, smem, st)
#define DH_TMA_SA(T, N)                                                                  \
    do {                                                                                 \
        if (scale255) { if (p.affine) DH_TMA(T, N, true, true); else DH_TMA(T, N, true, false); } \
        else          { if (p.affine) DH_TMA(T, N, false, true); else DH_TMA(T, N, false, false); } \
    } while (0)
    if (out_dtype == DH_F32) { if (nchw) DH_TMA_SA(float, true); else DH_TMA_SA(float, false); }
    else                     { if (nchw) DH_TMA_SA(__nv_bfloat16, true); else DH_TMA_SA(__nv_bfloat16, false); }
#undef DH_TMA_SA
#undef DH_TMA
    if (rc_launch != DH_OK) return rc_launch;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return cuda_fail(e, "gather_tma_kernel");
    return DH_OK;
}

}  // namespace dh
