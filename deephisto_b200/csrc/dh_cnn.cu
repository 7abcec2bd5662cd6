// Predictor-side helper kernel: 3x3 stride-2 max pooling over an NHWC bf16 tensor.
//
// Reference path: examples/predict_full_patched.py:66-78 batch_predictor -> model(features) with the patch_cls_simple ResNet18
// (models/patch_cls_simple/model.py:5-11): its stem is conv1 -> bn1 -> relu -> maxpool(kernel 3, stride 2, padding 1). The
// convolutions stay with cuDNN (torch); the pooling between them is pure HBM traffic -- [B,112,112,64] bf16 read once,
// [B,56,56,64] written once -- and torch's channels_last kernel needs 3.0 ms for it at batch 1024 (15 % of the whole forward,
// profiles/r02_predict.md) where the bytes take 0.3 ms. One block per output row, one thread per output pixel and 8 channels: nine 16-byte loads (the 2.25x
// window overlap is served by L1 / L2: neighbouring outputs share rows that were just read), packed bf16 maxima, one 16-byte store.
// Padding is -inf and NaN propagates, like torch.nn.functional.max_pool2d.
#include "dh_common.cuh"

namespace dh {

__device__ __forceinline__ uint32_t max2_bf16(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}

// One block per output row (b, oy); threads walk (ox, channel group) -- no 64-bit index arithmetic per element (a flat 64-bit index
// cost four divisions per thread: ~0.3 ms of the 0.47 ms the first version took at batch 1024).
__global__ void __launch_bounds__(512) maxpool3x3s2_nhwc_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int C8, int OH,
                                                                     int OW) {
    const uint32_t ninf = 0xFF80FF80u;  // two bf16 -inf
    const int64_t row = blockIdx.x;                       // b * OH + oy
    const int oy = (int)(row % OH);
    const int64_t b = row / OH;
    const uint4* img = in + b * H * (int64_t)W * C8;
    uint4* orow = out + row * (int64_t)OW * C8;
    const int n = OW * C8;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int ox = t / C8, c8 = t - ox * C8;
        uint4 m = make_uint4(ninf, ninf, ninf, ninf);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = 2 * oy - 1 + dy;
            if (iy < 0 || iy >= H) continue;
            const uint4* line = img + (int64_t)iy * W * C8 + c8;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = 2 * ox - 1 + dx;
                if (ix < 0 || ix >= W) continue;
                const uint4 v = __ldg(line + ix * C8);
                m.x = max2_bf16(m.x, v.x); m.y = max2_bf16(m.y, v.y); m.z = max2_bf16(m.z, v.z); m.w = max2_bf16(m.w, v.w);
            }
        }
        orow[t] = m;
    }
}

// The same pooling over a DEPTH-TO-SPACE stem output: in[b][Y][X][(P*2 + Q)*C + o] holds pixel (2Y + P, 2X + Q), channel o, of the
// [2H][2W][C] convolution output (the 4x4 space-to-depth stem produces 2x2 output pixels per block, examples/predict_full_patched.py
// FusedResNetForward); out[b][Y][X][o] = max over rows 2Y-1..2Y+1, columns 2X-1..2X+1 = blocks (Y-1, P=1), (Y, P=0), (Y, P=1) x likewise.
__global__ void __launch_bounds__(512) maxpool3x3s2_d2s_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int C8) {
    const uint32_t ninf = 0xFF80FF80u;
    const int64_t row = blockIdx.x;                       // b * H + oy
    const int oy = (int)(row % H);
    const int64_t b = row / H;
    const uint4* img = in + b * H * (int64_t)W * 4 * C8;
    uint4* orow = out + row * (int64_t)W * C8;
    const int n = W * C8;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int ox = t / C8, c8 = t - ox * C8;
        uint4 m = make_uint4(ninf, ninf, ninf, ninf);
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int Y = dy < 0 ? oy - 1 : oy, P = dy == 0 ? 0 : 1;
            if (Y < 0) continue;
            const uint4* line = img + (int64_t)Y * W * 4 * C8 + (P * 2) * C8 + c8;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int X = dx < 0 ? ox - 1 : ox, Q = dx == 0 ? 0 : 1;
                if (X < 0) continue;
                const uint4 v = __ldg(line + (X * 4 + Q) * C8);
                m.x = max2_bf16(m.x, v.x); m.y = max2_bf16(m.y, v.y); m.z = max2_bf16(m.z, v.z); m.w = max2_bf16(m.w, v.w);
            }
        }
        orow[t] = m;
    }
}

}  // namespace dh

using namespace dh;

extern "C" DH_API int dh_maxpool3x3s2_d2s(const void* in, int64_t B, int H, int W, int C, void* out, int dtype, void* stream) {
    if (B == 0) return DH_OK;
    DH_REQUIRE(in && out, "dh_maxpool3x3s2_d2s: null pointer");
    DH_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "dh_maxpool3x3s2_d2s: bad shape");
    DH_REQUIRE(dtype == DH_BF16, "dh_maxpool3x3s2_d2s: only bf16 is implemented");
    DH_REQUIRE(C % 8 == 0, "dh_maxpool3x3s2_d2s: the channel count must be a multiple of 8 (16-byte vectors)");
    DH_REQUIRE(reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0, "dh_maxpool3x3s2_d2s: buffers must be 16-byte aligned");
    const int C8 = C / 8;
    DH_REQUIRE(B * H < (1ll << 31) && (int64_t)W * 4 * C8 < (1ll << 24), "dh_maxpool3x3s2_d2s: tensor too large for the row-per-block launch");
    const int per_row = W * C8;
    const int threads = per_row >= 512 ? 512 : ((per_row + 31) / 32) * 32;
    maxpool3x3s2_d2s_bf16_kernel<<<(unsigned)(B * H), threads, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), H, W, C8);
    DH_CHECK_LAUNCH("maxpool3x3s2_d2s_bf16_kernel");
    return DH_OK;
}

extern "C" DH_API int dh_maxpool3x3s2_nhwc(const void* in, int64_t B, int H, int W, int C, void* out, int dtype, void* stream) {
    if (B == 0) return DH_OK;
    DH_REQUIRE(in && out, "dh_maxpool3x3s2_nhwc: null pointer");
    DH_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "dh_maxpool3x3s2_nhwc: bad shape");
    DH_REQUIRE(dtype == DH_BF16, "dh_maxpool3x3s2_nhwc: only bf16 is implemented");
    DH_REQUIRE(C % 8 == 0, "dh_maxpool3x3s2_nhwc: the channel count must be a multiple of 8 (16-byte vectors)");
    DH_REQUIRE(reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0, "dh_maxpool3x3s2_nhwc: buffers must be 16-byte aligned");
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    const int C8 = C / 8;
    DH_REQUIRE(B * OH < (1ll << 31) && (int64_t)W * C8 < (1ll << 24), "dh_maxpool3x3s2_nhwc: tensor too large for the row-per-block launch");
    const int per_row = OW * C8;
    const int threads = per_row >= 512 ? 512 : ((per_row + 31) / 32) * 32;
    maxpool3x3s2_nhwc_bf16_kernel<<<(unsigned)(B * OH), threads, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), H, W,
                                                                                         C8, OH, OW);
    DH_CHECK_LAUNCH("maxpool3x3s2_nhwc_bf16_kernel");
    return DH_OK;
}
