// K1: fused patch gather + normalise.
//
// Reference path replaced (xubiker/deephisto):
//   patch_samplers/full_samplers.py:353-369  _generate_batch_memory (views data[y:y+ps, x:x+ps, :])
//   patch_samplers/full_samplers.py:437-452  generator_torch: np.stack -> astype(float32) -> /255 -> torch.tensor
//   patch_samplers/full_samplers.py:282-290  rnd generator_torch (no /255)
//   examples/predict_full_patched.py:66-71   batch_predictor: stack/255, permute(0,3,1,2).contiguous()
//   patch_samplers/region_samplers.py:616    torch.tensor(patch.data, float32) / 255
//
// The slide is resident in HBM as uint8 [H][W][3] (row pitch in bytes). One pass: every input byte
// is read once (neighbouring lanes share words through L1), every output element is written once
// with full-line coalesced vector stores. The kernel is write-dominated (1 B in -> 4 B / 2 B out),
// so the store path decides the achieved fraction of HBM bandwidth.
#include "dh_common.cuh"

namespace dh {

struct FastDiv {
    uint32_t mul, shr, den;
    __host__ void init(uint32_t d) {
        den = d;
        if (d == 1) { mul = 0; shr = 0; return; }
        uint32_t lg = 0;
        while ((1ull << lg) < d) ++lg;
        uint32_t p = 31 + lg;
        mul = (uint32_t)(((1ull << p) + d - 1) / d);
        shr = p - 32;
    }
    // valid for n < 2^31
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return den == 1 ? n : (__umulhi(n, mul) >> shr); }
};

struct GatherParams {
    const uint8_t* slide;
    int64_t H, W, pitch;
    const int32_t* coords;
    const int32_t* out_index;
    const uint8_t* flip;
    void* out;
    int64_t B;
    int ps;
    int scale255;
    int affine;
    float mean[3];
    float stdv[3];
    FastDiv row_units;   // units per patch row
    uint32_t units;      // units per patch
    uint32_t chunks;     // chunks per patch
};

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kChunkUnits = kThreads * kUnroll;

template <bool SCALE, bool AFFINE>
__device__ __forceinline__ float norm_value(uint32_t v, int c, const GatherParams& p) {
    float f = (float)v;
    if (SCALE) f = div255_exact(f);
    if (AFFINE) f = __fdiv_rn(__fsub_rn(f, p.mean[c]), p.stdv[c]);
    return f;
}

template <typename OutT> struct Pack4;
template <> struct Pack4<float> {
    static __device__ __forceinline__ void store(float* dst, float a, float b, float c, float d) {
        __stcs(reinterpret_cast<float4*>(dst), make_float4(a, b, c, d));
    }
};
template <> struct Pack4<__nv_bfloat16> {
    static __device__ __forceinline__ void store(__nv_bfloat16* dst, float a, float b, float c, float d) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(a, b);
        __nv_bfloat162 hi = __floats2bfloat162_rn(c, d);
        uint2 v;
        v.x = *reinterpret_cast<uint32_t*>(&lo);
        v.y = *reinterpret_cast<uint32_t*>(&hi);
        __stcs(reinterpret_cast<uint2*>(dst), v);
    }
};

// ---------------------------------------------------------------------------------------------
// Vector kernel, NHWC. unit = 4 consecutive output elements = 4 consecutive input bytes of a row.
// Requires ps % 4 == 0, slide and pitch 4-byte aligned. Out-of-slide / flipped patches take the
// byte path inside the same kernel (warp-uniform branch except at patch boundaries).
// ---------------------------------------------------------------------------------------------
template <typename OutT, bool SCALE, bool AFFINE>
__global__ void __launch_bounds__(kThreads) gather_nhwc_vec(const GatherParams p) {
    const uint32_t rowlen = 3u * (uint32_t)p.ps;
    const int64_t total = p.B * (int64_t)p.chunks;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int64_t patch = w / p.chunks;
        const uint32_t chunk = (uint32_t)(w - patch * p.chunks);
        const int y = __ldg(p.coords + 2 * patch);
        const int x = __ldg(p.coords + 2 * patch + 1);
        const int64_t slot = p.out_index ? (int64_t)__ldg(p.out_index + patch) : patch;
        const uint32_t fl = p.flip ? (uint32_t)__ldg(p.flip + patch) : 0u;
        OutT* outp = reinterpret_cast<OutT*>(p.out) + slot * (int64_t)p.units * 4;
        const bool inside = (y >= 0) && (x >= 0) && ((int64_t)y + p.ps <= p.H) && ((int64_t)x + p.ps <= p.W);
        const bool fast = inside && !(fl & DH_FLIP_H);

        uint32_t u[kUnroll];
        uint32_t bytes[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            u[j] = chunk * kChunkUnits + j * kThreads + threadIdx.x;
            bytes[j] = 0;
        }
        if (fast) {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                if (u[j] < p.units) {
                    uint32_t r = p.row_units.div(u[j]);
                    uint32_t off = (u[j] - r * p.row_units.den) * 4u;
                    uint32_t sr = (fl & DH_FLIP_V) ? (uint32_t)p.ps - 1u - r : r;
                    const uint8_t* src = p.slide + (int64_t)(y + (int)sr) * p.pitch + 3 * (int64_t)x + off;
                    uintptr_t a = reinterpret_cast<uintptr_t>(src);
                    const uint32_t* a0 = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                    uint32_t sh = (uint32_t)(a & 3) * 8u;
                    uint32_t w0 = __ldg(a0);
                    uint32_t w1 = sh ? __ldg(a0 + 1) : 0u;
                    bytes[j] = __funnelshift_r(w0, w1, sh);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                if (u[j] < p.units) {
                    uint32_t r = p.row_units.div(u[j]);
                    uint32_t off = (u[j] - r * p.row_units.den) * 4u;
                    uint32_t sr = (fl & DH_FLIP_V) ? (uint32_t)p.ps - 1u - r : r;
                    int64_t yy = (int64_t)y + sr;
                    uint32_t acc = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t e = off + i;
                        uint32_t col = e / 3u, ch = e - col * 3u;
                        uint32_t scol = (fl & DH_FLIP_H) ? (uint32_t)p.ps - 1u - col : col;
                        int64_t xx = (int64_t)x + scol;
                        uint32_t v = 0;
                        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) v = __ldg(p.slide + yy * p.pitch + 3 * xx + ch);
                        acc |= v << (8 * i);
                    }
                    bytes[j] = acc;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            if (u[j] < p.units) {
                uint32_t r = p.row_units.div(u[j]);
                uint32_t off = (u[j] - r * p.row_units.den) * 4u;
                uint32_t c0 = off % 3u;  // channel of the first element
                uint32_t c1 = c0 == 2 ? 0 : c0 + 1, c2 = c1 == 2 ? 0 : c1 + 1;
                float f0 = norm_value<SCALE, AFFINE>(bytes[j] & 255u, c0, p);
                float f1 = norm_value<SCALE, AFFINE>((bytes[j] >> 8) & 255u, c1, p);
                float f2 = norm_value<SCALE, AFFINE>((bytes[j] >> 16) & 255u, c2, p);
                float f3 = norm_value<SCALE, AFFINE>(bytes[j] >> 24, c0, p);
                Pack4<OutT>::store(outp + (int64_t)u[j] * 4, f0, f1, f2, f3);
            }
        }
        (void)rowlen;
    }
}

// ---------------------------------------------------------------------------------------------
// Vector kernel, NCHW. unit = 4 consecutive pixels of a patch row = 12 input bytes -> one 4-wide
// store into each of the three channel planes.
// ---------------------------------------------------------------------------------------------
template <typename OutT, bool SCALE, bool AFFINE>
__global__ void __launch_bounds__(kThreads) gather_nchw_vec(const GatherParams p) {
    const int64_t total = p.B * (int64_t)p.chunks;
    const int64_t plane = (int64_t)p.ps * p.ps;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int64_t patch = w / p.chunks;
        const uint32_t chunk = (uint32_t)(w - patch * p.chunks);
        const int y = __ldg(p.coords + 2 * patch);
        const int x = __ldg(p.coords + 2 * patch + 1);
        const int64_t slot = p.out_index ? (int64_t)__ldg(p.out_index + patch) : patch;
        const uint32_t fl = p.flip ? (uint32_t)__ldg(p.flip + patch) : 0u;
        OutT* outp = reinterpret_cast<OutT*>(p.out) + slot * 3 * plane;
        const bool inside = (y >= 0) && (x >= 0) && ((int64_t)y + p.ps <= p.H) && ((int64_t)x + p.ps <= p.W);

        uint32_t u[kUnroll];
        uint32_t v0[kUnroll], v1[kUnroll], v2[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            u[j] = chunk * kChunkUnits + j * kThreads + threadIdx.x;
            v0[j] = v1[j] = v2[j] = 0;
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            if (u[j] < p.units) {
                uint32_t r = p.row_units.div(u[j]);
                uint32_t col = (u[j] - r * p.row_units.den) * 4u;
                uint32_t sr = (fl & DH_FLIP_V) ? (uint32_t)p.ps - 1u - r : r;
                uint32_t scol = (fl & DH_FLIP_H) ? (uint32_t)p.ps - 4u - col : col;  // first source pixel
                if (inside) {
                    const uint8_t* src = p.slide + (int64_t)(y + (int)sr) * p.pitch + 3 * ((int64_t)x + scol);
                    uintptr_t a = reinterpret_cast<uintptr_t>(src);
                    const uint32_t* a0 = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                    uint32_t sh = (uint32_t)(a & 3) * 8u;
                    uint32_t w0 = __ldg(a0), w1 = __ldg(a0 + 1), w2 = __ldg(a0 + 2);
                    uint32_t w3 = sh ? __ldg(a0 + 3) : 0u;
                    v0[j] = __funnelshift_r(w0, w1, sh);
                    v1[j] = __funnelshift_r(w1, w2, sh);
                    v2[j] = __funnelshift_r(w2, w3, sh);
                } else {
                    int64_t yy = (int64_t)y + sr;
                    uint32_t b[3] = {0, 0, 0};
#pragma unroll
                    for (int i = 0; i < 12; ++i) {
                        int64_t xx = (int64_t)x + scol + i / 3;
                        uint32_t v = 0;
                        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) v = __ldg(p.slide + yy * p.pitch + 3 * xx + (i % 3));
                        b[i / 4] |= v << (8 * (i % 4));
                    }
                    v0[j] = b[0]; v1[j] = b[1]; v2[j] = b[2];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kUnroll; ++j) {
            if (u[j] < p.units) {
                uint32_t r = p.row_units.div(u[j]);
                uint32_t col = (u[j] - r * p.row_units.den) * 4u;
                // bytes: v0 = R0 G0 B0 R1 | v1 = G1 B1 R2 G2 | v2 = B2 R3 G3 B3
                uint32_t px[4][3];
                px[0][0] = v0[j] & 255u;         px[0][1] = (v0[j] >> 8) & 255u;  px[0][2] = (v0[j] >> 16) & 255u;
                px[1][0] = v0[j] >> 24;          px[1][1] = v1[j] & 255u;         px[1][2] = (v1[j] >> 8) & 255u;
                px[2][0] = (v1[j] >> 16) & 255u; px[2][1] = v1[j] >> 24;          px[2][2] = v2[j] & 255u;
                px[3][0] = (v2[j] >> 8) & 255u;  px[3][1] = (v2[j] >> 16) & 255u; px[3][2] = v2[j] >> 24;
                const bool fh = fl & DH_FLIP_H;
                OutT* o = outp + (int64_t)r * p.ps + col;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float f0 = norm_value<SCALE, AFFINE>(px[fh ? 3 : 0][c], c, p);
                    float f1 = norm_value<SCALE, AFFINE>(px[fh ? 2 : 1][c], c, p);
                    float f2 = norm_value<SCALE, AFFINE>(px[fh ? 1 : 2][c], c, p);
                    float f3 = norm_value<SCALE, AFFINE>(px[fh ? 0 : 3][c], c, p);
                    Pack4<OutT>::store(o + c * plane, f0, f1, f2, f3);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Generic scalar kernel: any ps / alignment / dtype (incl. raw u8) / layout. One output element
// per thread iteration. Correctness fallback for shapes the vector kernels do not cover.
// ---------------------------------------------------------------------------------------------
template <typename OutT>
__device__ __forceinline__ OutT cast_out(float f);
template <> __device__ __forceinline__ float cast_out<float>(float f) { return f; }
template <> __device__ __forceinline__ __nv_bfloat16 cast_out<__nv_bfloat16>(float f) { return __float2bfloat16_rn(f); }

template <typename OutT, bool NCHW>
__global__ void __launch_bounds__(kThreads) gather_generic(const GatherParams p) {
    const int64_t per_patch = (int64_t)p.ps * p.ps * 3;
    const int64_t total = p.B * per_patch;
    const int64_t plane = (int64_t)p.ps * p.ps;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t patch = g / per_patch;
        const int64_t e = g - patch * per_patch;
        int r, col, ch;
        if (NCHW) {
            ch = (int)(e / plane);
            int64_t rem = e - ch * plane;
            r = (int)(rem / p.ps);
            col = (int)(rem - (int64_t)r * p.ps);
        } else {
            r = (int)(e / (3 * p.ps));
            int64_t rem = e - (int64_t)r * 3 * p.ps;
            col = (int)(rem / 3);
            ch = (int)(rem - col * 3);
        }
        const int y = __ldg(p.coords + 2 * patch);
        const int x = __ldg(p.coords + 2 * patch + 1);
        const int64_t slot = p.out_index ? (int64_t)__ldg(p.out_index + patch) : patch;
        const uint32_t fl = p.flip ? (uint32_t)__ldg(p.flip + patch) : 0u;
        int sr = (fl & DH_FLIP_V) ? p.ps - 1 - r : r;
        int sc = (fl & DH_FLIP_H) ? p.ps - 1 - col : col;
        int64_t yy = (int64_t)y + sr, xx = (int64_t)x + sc;
        uint32_t v = 0;
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) v = __ldg(p.slide + yy * p.pitch + 3 * xx + ch);
        if constexpr (sizeof(OutT) == 1) {
            reinterpret_cast<uint8_t*>(p.out)[slot * per_patch + e] = (uint8_t)v;
        } else {
            float f = (float)v;
            if (p.scale255) f = div255_exact(f);
            if (p.affine) f = __fdiv_rn(__fsub_rn(f, p.mean[ch]), p.stdv[ch]);
            reinterpret_cast<OutT*>(p.out)[slot * per_patch + e] = cast_out<OutT>(f);
        }
    }
}

static int g_variant = 0;

int gather_tma_launch(const uint8_t* slide, int64_t H, int64_t W, int64_t pitch, const int64_t* slides_dev, int n_slides,
                      const int32_t* image, const int32_t* coords, const int32_t* out_index,
                      int64_t B, int ps, void* out, int out_dtype, int out_layout, int scale255, const float* mean3, const float* std3,
                      const uint8_t* flip, int debug, cudaStream_t st);

template <typename OutT>
static void launch_vec(const GatherParams& p, bool nchw, int grid, cudaStream_t st) {
#define DH_LAUNCH(K, S, A) K<OutT, S, A><<<grid, kThreads, 0, st>>>(p)
    if (nchw) {
        if (p.scale255) { if (p.affine) DH_LAUNCH(gather_nchw_vec, true, true); else DH_LAUNCH(gather_nchw_vec, true, false); }
        else            { if (p.affine) DH_LAUNCH(gather_nchw_vec, false, true); else DH_LAUNCH(gather_nchw_vec, false, false); }
    } else {
        if (p.scale255) { if (p.affine) DH_LAUNCH(gather_nhwc_vec, true, true); else DH_LAUNCH(gather_nhwc_vec, true, false); }
        else            { if (p.affine) DH_LAUNCH(gather_nhwc_vec, false, true); else DH_LAUNCH(gather_nhwc_vec, false, false); }
    }
#undef DH_LAUNCH
}

}  // namespace dh

using namespace dh;

extern "C" DH_API int dh_gather_set_variant(int variant) {
    if (variant < 0 || variant > 9) { set_error("dh_gather_set_variant: variant must be 0..9"); return DH_ERR_INVALID; }
    g_variant = variant;
    return DH_OK;
}

extern "C" DH_API int dh_gather_normalize(const uint8_t* slide, int64_t H, int64_t W, int64_t pitch, const int32_t* coords,
                                   const int32_t* out_index, int64_t B, int ps, void* out, int out_dtype,
                                   int out_layout, int scale255, const float* mean3_host, const float* std3_host,
                                   const uint8_t* flip, void* stream) {
    if (B == 0) return DH_OK;  // an empty batch is a no-op (empty tensors have null data pointers)
    DH_REQUIRE(slide && coords && out, "dh_gather_normalize: null pointer");
    DH_REQUIRE(H > 0 && W > 0 && pitch >= 3 * W, "dh_gather_normalize: bad slide shape H=%lld W=%lld pitch=%lld",
               (long long)H, (long long)W, (long long)pitch);
    DH_REQUIRE(ps > 0 && ps <= 8192, "dh_gather_normalize: patch size %d out of range", ps);
    DH_REQUIRE(B >= 0, "dh_gather_normalize: negative batch");
    DH_REQUIRE(out_dtype == DH_F32 || out_dtype == DH_BF16 || out_dtype == DH_U8, "dh_gather_normalize: bad dtype %d", out_dtype);
    const bool s2d_any = out_layout == DH_S2D16 || out_layout == DH_S2D48;
    DH_REQUIRE(out_layout == DH_NHWC || out_layout == DH_NCHW || s2d_any, "dh_gather_normalize: bad layout %d", out_layout);
    DH_REQUIRE(!s2d_any || (out_dtype == DH_BF16 && ps % (out_layout == DH_S2D48 ? 4 : 2) == 0 && pitch % 16 == 0 && reinterpret_cast<uintptr_t>(slide) % 16 == 0 &&
                            reinterpret_cast<uintptr_t>(out) % 16 == 0 && (out_layout == DH_S2D48 || out_index == nullptr)),
               "dh_gather_normalize: the space-to-depth layouts need bf16 output, a patch size divisible by 2 (S2D16) / 4 (S2D48), a 16-byte aligned slide (pitch %% 16 == 0) and output");
    DH_REQUIRE((mean3_host == nullptr) == (std3_host == nullptr), "dh_gather_normalize: mean and std must both be given or both be NULL");
    if (B == 0) return DH_OK;
    cudaStream_t st = as_stream(stream);

    GatherParams p{};
    p.slide = slide; p.H = H; p.W = W; p.pitch = pitch;
    p.coords = coords; p.out_index = out_index; p.flip = flip; p.out = out;
    p.B = B; p.ps = ps; p.scale255 = scale255 ? 1 : 0;
    p.affine = mean3_host ? 1 : 0;
    for (int c = 0; c < 3; ++c) {
        p.mean[c] = mean3_host ? mean3_host[c] : 0.f;
        p.stdv[c] = std3_host ? std3_host[c] : 1.f;
        if (mean3_host) DH_REQUIRE(p.stdv[c] != 0.f, "dh_gather_normalize: std[%d] == 0", c);
    }
    const bool nchw = out_layout == DH_NCHW;
    if (g_variant != 1 || s2d_any) {  // TMA-staged kernel whenever the shape allows it (the only one with the space-to-depth mode)
        int rc = gather_tma_launch(slide, H, W, pitch, nullptr, 1, nullptr, coords, out_index, B, ps, out, out_dtype, out_layout, p.scale255,
                                   mean3_host ? p.mean : nullptr, mean3_host ? p.stdv : nullptr, flip, g_variant >= 3 ? (g_variant == 6 ? 4 : (g_variant == 7 ? 8 : (g_variant == 8 ? 16 : (g_variant == 9 ? 24 : g_variant - 2)))) : 0, st);
        if (rc != DH_ERR_UNSUPPORTED) return rc;
        if (s2d_any) { set_error("dh_gather_normalize: patch size %d not supported by the space-to-depth mode of the TMA-staged kernel", ps); return rc; }
        if (g_variant >= 2) { set_error("dh_gather_normalize: shape not supported by the TMA-staged kernel (needs ps %% 4 == 0 for f32, ps %% 8 == 0 for bf16, pitch %% 16 == 0)"); return rc; }
    }
    const size_t esz = out_dtype == DH_F32 ? 4 : (out_dtype == DH_BF16 ? 2 : 1);
    const bool aligned = (reinterpret_cast<uintptr_t>(slide) % 4 == 0) && (pitch % 4 == 0) && (ps % 4 == 0) &&
                         (reinterpret_cast<uintptr_t>(out) % (4 * esz) == 0);
    if (out_dtype != DH_U8 && aligned) {
        if (nchw) { p.units = (uint32_t)ps * ps / 4; p.row_units.init(ps / 4); }
        else      { p.units = (uint32_t)ps * ps * 3 / 4; p.row_units.init(3 * ps / 4); }
        p.chunks = (p.units + kChunkUnits - 1) / kChunkUnits;
        int64_t work = B * (int64_t)p.chunks;
        int grid = (int)(work < (int64_t)kNumSMs * 8 ? work : (int64_t)kNumSMs * 8);
        if (out_dtype == DH_F32) launch_vec<float>(p, nchw, grid, st);
        else launch_vec<__nv_bfloat16>(p, nchw, grid, st);
        DH_CHECK_LAUNCH("gather_vec");
        return DH_OK;
    }
    int64_t total = B * (int64_t)ps * ps * 3;
    int64_t blocks = (total + kThreads - 1) / kThreads;
    int grid = (int)(blocks < (int64_t)kNumSMs * 16 ? blocks : (int64_t)kNumSMs * 16);
    if (out_dtype == DH_F32) { if (nchw) gather_generic<float, true><<<grid, kThreads, 0, st>>>(p); else gather_generic<float, false><<<grid, kThreads, 0, st>>>(p); }
    else if (out_dtype == DH_BF16) { if (nchw) gather_generic<__nv_bfloat16, true><<<grid, kThreads, 0, st>>>(p); else gather_generic<__nv_bfloat16, false><<<grid, kThreads, 0, st>>>(p); }
    else { if (nchw) gather_generic<uint8_t, true><<<grid, kThreads, 0, st>>>(p); else gather_generic<uint8_t, false><<<grid, kThreads, 0, st>>>(p); }
    DH_CHECK_LAUNCH("gather_generic");
    return DH_OK;
}

extern "C" DH_API int dh_gather_normalize_multi(const int64_t* slides_host, const int64_t* slides_dev, int n_slides,
                                                const int32_t* image_of_patch, const int32_t* coords, const int32_t* out_index, int64_t B,
                                                int ps, void* out, int out_dtype, int out_layout, int scale255,
                                                const float* mean3_host, const float* std3_host, const uint8_t* flip, void* stream) {
    if (B == 0) return DH_OK;
    DH_REQUIRE(slides_host && slides_dev && image_of_patch && coords && out, "dh_gather_normalize_multi: null pointer");
    DH_REQUIRE(n_slides >= 1, "dh_gather_normalize_multi: empty slide table");
    DH_REQUIRE(ps > 0 && ps <= 8192 && B >= 0, "dh_gather_normalize_multi: bad patch size / batch");
    DH_REQUIRE(out_dtype == DH_F32 || out_dtype == DH_BF16, "dh_gather_normalize_multi: output must be f32 or bf16");
    DH_REQUIRE(out_layout == DH_NHWC || out_layout == DH_NCHW, "dh_gather_normalize_multi: bad layout %d", out_layout);
    DH_REQUIRE((mean3_host == nullptr) == (std3_host == nullptr), "dh_gather_normalize_multi: mean and std must both be given or both be NULL");
    DH_REQUIRE(reinterpret_cast<uintptr_t>(slides_dev) % 16 == 0, "dh_gather_normalize_multi: the device slide table must be 16-byte aligned");
    for (int i = 0; i < n_slides; ++i) {
        const int64_t* d = slides_host + 4 * i;
        DH_REQUIRE(d[0] != 0 && d[1] > 0 && d[2] > 0 && d[3] >= 3 * d[2], "dh_gather_normalize_multi: bad descriptor of slide %d", i);
        if (d[0] % 16 != 0 || d[3] % 16 != 0) {
            set_error("dh_gather_normalize_multi: slide %d is not 16-byte aligned (data, pitch): gather it with dh_gather_normalize", i);
            return DH_ERR_UNSUPPORTED;
        }
    }
    if (mean3_host) for (int c = 0; c < 3; ++c) DH_REQUIRE(std3_host[c] != 0.f, "dh_gather_normalize_multi: std[%d] == 0", c);
    if (B == 0) return DH_OK;
    int rc = gather_tma_launch(reinterpret_cast<const uint8_t*>(slides_host[0]), slides_host[1], slides_host[2], slides_host[3], slides_dev, n_slides,
                               image_of_patch, coords, out_index, B, ps, out, out_dtype, out_layout, scale255 ? 1 : 0, mean3_host, std3_host, flip,
                               0, as_stream(stream));
    if (rc == DH_ERR_UNSUPPORTED)
        set_error("dh_gather_normalize_multi: patch size %d not supported by the TMA-staged kernel (needs ps %% 4 == 0 for f32, ps %% 8 == 0 for bf16)", ps);
    return rc;
}
