// K2: dense grid enumeration + synthetic slide generator + library plumbing (errors, version).
//
// Reference path replaced:
//   patch_samplers/full_samplers.py:374-404  FullImageDenseSampler._create_batched_coords
#include <stdarg.h>
#include <string.h>

#include "dh_common.cuh"

namespace dh {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return DH_ERR_CUDA;
}

// ---- dense enumeration -----------------------------------------------------------------------
// ny = len(range(0, h-ps, stride)), nx likewise; N = ny*nx + ny + nx + 1.
struct DenseGrid {
    int64_t H, W;
    int ps, stride;
    int64_t ny, nx, N, Npad;
};

static inline int64_t range_len(int64_t stop, int64_t step) { return stop <= 0 ? 0 : (stop + step - 1) / step; }

static int make_grid(int64_t H, int64_t W, int ps, int stride, int batch, DenseGrid* g) {
    DH_REQUIRE(ps > 0 && stride > 0, "dense grid: patch size and stride must be positive");
    DH_REQUIRE(H >= ps && W >= ps, "dense grid: slide %lldx%lld smaller than patch %d", (long long)H, (long long)W, ps);
    DH_REQUIRE(H < (1ll << 31) && W < (1ll << 31), "dense grid: slide side must fit int32");
    g->H = H; g->W = W; g->ps = ps; g->stride = stride;
    g->ny = range_len(H - ps, stride);
    g->nx = range_len(W - ps, stride);
    g->N = g->ny * g->nx + g->ny + g->nx + 1;
    g->Npad = batch > 0 ? (g->N + batch - 1) / batch * batch : g->N;
    return DH_OK;
}

__global__ void dense_coords_kernel(DenseGrid g, int64_t first, int64_t count, int32_t* __restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = first + t;
        int64_t main_n = g.ny * g.nx;
        int32_t y, x;
        if (i < main_n) {
            int64_t gy = i / g.nx;
            y = (int32_t)(gy * g.stride);
            x = (int32_t)((i - gy * g.nx) * g.stride);
        } else if (i < main_n + g.ny) {  // last column
            y = (int32_t)((i - main_n) * g.stride);
            x = (int32_t)(g.W - g.ps);
        } else if (i < main_n + g.ny + g.nx) {  // last row
            y = (int32_t)(g.H - g.ps);
            x = (int32_t)((i - main_n - g.ny) * g.stride);
        } else {  // corner and its padding copies
            y = (int32_t)(g.H - g.ps);
            x = (int32_t)(g.W - g.ps);
        }
        reinterpret_cast<int2*>(out)[t] = make_int2(y, x);
    }
}

// ---- synthetic slide ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
__host__ __device__ __forceinline__ uint32_t synth_word(uint64_t widx, uint32_t seed_lo, uint32_t seed_hi) {
    uint32_t a = (uint32_t)widx ^ seed_lo;
    uint32_t b = (uint32_t)(widx >> 32) ^ seed_hi;
    return fmix32(fmix32(a) + b * 0x9E3779B9u);
}

// writes rows [y0, y0 + H) of the logical slide into a buffer whose row 0 is slide row y0
__global__ void synth_slide_kernel(uint8_t* __restrict__ slide, int64_t y0, int64_t H, int64_t W, int64_t pitch, uint32_t seed_lo,
                                   uint32_t seed_hi, bool aligned) {
    const int64_t row_bytes = 3 * W;
    const int64_t words_per_row = (row_bytes + 3) / 4;
    const int64_t total = H * words_per_row;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t y = t / words_per_row;
        int64_t j = t - y * words_per_row;
        uint64_t k0 = (uint64_t)(y0 + y) * (uint64_t)row_bytes + 4ull * j;
        uint64_t wi = k0 >> 2;
        uint32_t sh = (uint32_t)(k0 & 3) * 8u;
        uint32_t h0 = synth_word(wi, seed_lo, seed_hi);
        uint32_t h1 = sh ? synth_word(wi + 1, seed_lo, seed_hi) : 0u;
        uint32_t v = sh ? ((h0 >> sh) | (h1 << (32 - sh))) : h0;
        uint8_t* dst = slide + y * pitch + 4 * j;
        if (aligned && 4 * j + 4 <= row_bytes) {
            *reinterpret_cast<uint32_t*>(dst) = v;
        } else {
            for (int i = 0; i < 4 && 4 * j + i < row_bytes; ++i) dst[i] = (uint8_t)(v >> (8 * i));
        }
    }
}

}  // namespace dh

using namespace dh;

extern "C" DH_API int dh_version(void) { return DH_VERSION; }
extern "C" DH_API const char* dh_last_error(void) { return g_err; }

extern "C" DH_API int dh_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("dh_device_check: no CUDA device (%s)", cudaGetErrorString(e)); return DH_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    if (prop.major != 10) {
        set_error("dh_device_check: device %d is sm_%d%d; this library contains sm_100a code only", dev, prop.major, prop.minor);
        return DH_ERR_NO_DEVICE;
    }
    return DH_OK;
}

extern "C" DH_API int64_t dh_dense_count(int64_t H, int64_t W, int ps, int stride, int batch_size, int64_t* n_padded_host) {
    DenseGrid g;
    int rc = make_grid(H, W, ps, stride, batch_size, &g);
    if (rc != DH_OK) return rc;
    if (n_padded_host) *n_padded_host = g.Npad;
    return g.N;
}

extern "C" DH_API int dh_dense_coords(int64_t H, int64_t W, int ps, int stride, int batch_size, int64_t first, int64_t count,
                               int32_t* coords_out, void* stream) {
    DenseGrid g;
    int rc = make_grid(H, W, ps, stride, batch_size, &g);
    if (rc != DH_OK) return rc;
    if (count == 0 && first >= 0 && first <= g.Npad) return DH_OK;
    DH_REQUIRE(coords_out, "dh_dense_coords: null output");
    DH_REQUIRE(first >= 0 && count >= 0 && first + count <= g.Npad, "dh_dense_coords: range [%lld, %lld) outside the %lld padded patches",
               (long long)first, (long long)(first + count), (long long)g.Npad);
    DH_REQUIRE(reinterpret_cast<uintptr_t>(coords_out) % 8 == 0, "dh_dense_coords: output must be 8-byte aligned");
    if (count == 0) return DH_OK;
    int64_t blocks = (count + 255) / 256;
    int grid = (int)(blocks < kNumSMs * 8 ? blocks : kNumSMs * 8);
    dense_coords_kernel<<<grid, 256, 0, as_stream(stream)>>>(g, first, count, coords_out);
    DH_CHECK_LAUNCH("dense_coords_kernel");
    return DH_OK;
}

extern "C" DH_API int dh_synth_slide_rows(uint8_t* slide, int64_t H, int64_t W, int64_t pitch, int64_t y0, int64_t rows, uint64_t seed,
                                          void* stream) {
    DH_REQUIRE(slide, "dh_synth_slide: null pointer");
    DH_REQUIRE(H > 0 && W > 0 && pitch >= 3 * W, "dh_synth_slide: bad shape");
    DH_REQUIRE(y0 >= 0 && rows >= 0 && y0 + rows <= H, "dh_synth_slide: rows [%lld, %lld) outside the %lld-row slide", (long long)y0,
               (long long)(y0 + rows), (long long)H);
    if (rows == 0) return DH_OK;
    bool aligned = (reinterpret_cast<uintptr_t>(slide) % 4 == 0) && (pitch % 4 == 0);
    int grid = kNumSMs * 16;
    synth_slide_kernel<<<grid, 256, 0, as_stream(stream)>>>(slide, y0, rows, W, pitch, (uint32_t)seed, (uint32_t)(seed >> 32), aligned);
    DH_CHECK_LAUNCH("synth_slide_kernel");
    return DH_OK;
}

extern "C" DH_API int dh_synth_slide(uint8_t* slide, int64_t H, int64_t W, int64_t pitch, uint64_t seed, void* stream) {
    return dh_synth_slide_rows(slide, H, W, pitch, 0, H, seed, stream);
}

extern "C" DH_API int dh_upload_rects(uint8_t* slide_dev, int64_t H, int64_t pitch, const uint8_t* slide_host, int64_t n_rects,
                                      const int64_t* rects_host, void* stream) {
    if (n_rects == 0) return DH_OK;
    DH_REQUIRE(slide_dev && slide_host && rects_host, "dh_upload_rects: null pointer");
    DH_REQUIRE(H > 0 && pitch > 0 && n_rects > 0, "dh_upload_rects: bad sizes");
    for (int64_t i = 0; i < n_rects; ++i) {
        const int64_t y0 = rects_host[4 * i], y1 = rects_host[4 * i + 1], b0 = rects_host[4 * i + 2], b1 = rects_host[4 * i + 3];
        DH_REQUIRE(0 <= y0 && y0 <= y1 && y1 <= H && 0 <= b0 && b0 <= b1 && b1 <= pitch, "dh_upload_rects: rectangle %lld outside the slide",
                   (long long)i);
        if (y1 == y0 || b1 == b0) continue;
        cudaError_t e = cudaMemcpy2DAsync(slide_dev + y0 * pitch + b0, (size_t)pitch, slide_host + y0 * pitch + b0, (size_t)pitch, (size_t)(b1 - b0),
                                          (size_t)(y1 - y0), cudaMemcpyHostToDevice, as_stream(stream));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy2DAsync");
    }
    return DH_OK;
}

extern "C" DH_API int dh_host_device_pointer(const void* host_ptr, uint64_t* device_ptr_out_host) {
    DH_REQUIRE(host_ptr && device_ptr_out_host, "dh_host_device_pointer: null pointer");
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, host_ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(e, "cudaPointerGetAttributes"); }
    if (attr.type != cudaMemoryTypeHost || attr.devicePointer == nullptr) {
        set_error("dh_host_device_pointer: %p is not page-locked host memory mapped into the device address space", host_ptr);
        return DH_ERR_UNSUPPORTED;
    }
    *device_ptr_out_host = (uint64_t)reinterpret_cast<uintptr_t>(attr.devicePointer);
    return DH_OK;
}
