// Shared helpers for the deephisto_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/deephisto_b200.h"

namespace dh {

// ---- error plumbing --------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DH_REQUIRE(cond, ...)              \
    do {                                   \
        if (!(cond)) {                     \
            ::dh::set_error(__VA_ARGS__);  \
            return DH_ERR_INVALID;         \
        }                                  \
    } while (0)

#define DH_CHECK_LAUNCH(what)                                   \
    do {                                                        \
        cudaError_t e__ = cudaPeekAtLastError();                \
        if (e__ != cudaSuccess) return ::dh::cuda_fail(e__, what); \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- Philox4x32-10 (Salmon et al., SC'11). Restated on the CPU in oracle/philox.py. -----------
struct Philox4 {
    uint32_t v[4];
};

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
        philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// stream tags (counter word 3); the full contract is in DESIGN.md "Philox contract"
constexpr uint32_t kStreamTable = 1;    // image / table choice per chunk
constexpr uint32_t kStreamGroup = 2;    // class + region choice per group
constexpr uint32_t kStreamAttempt = 3;  // (x, y) candidate per slot attempt
constexpr uint32_t kStreamCoverTop = 4; // coverage sampler: top-up cells
constexpr uint32_t kStreamCoverPick = 5;// coverage sampler: Fisher-Yates picks
constexpr uint32_t kStreamCoverJit = 6; // coverage sampler: jitter

// uniform integer in [0, n) from one 32-bit word (multiply-shift; bias <= n / 2^32)
__host__ __device__ __forceinline__ uint32_t bounded_u32(uint32_t r, uint32_t n) {
#ifdef __CUDA_ARCH__
    return __umulhi(r, n);
#else
    return (uint32_t)(((uint64_t)r * (uint64_t)n) >> 32);
#endif
}

// exact float(v)/255.0f for v in 0..255 in two operations: 1/255 split into RN(1/255) + a float32 tail, the tail product
// rounded once and the head product added unrounded by the FMA. The result is the correctly rounded quotient for all 256
// inputs (v/255 is never within 2^-33 relative of a rounding boundary, the split's error is ~2^-46; checked exhaustively
// with exact rationals in tests/test_oracle_cpu.py and on the GPU in tests/test_gpu_parity.py).
__device__ __forceinline__ float div255_exact(float x) {
    const float rcp_hi = 0x1.010102p-8f;    // RN(1/255)
    const float rcp_lo = -0x1.fdfdfep-33f;  // RN(1/255 - rcp_hi)
    return __fmaf_rn(x, rcp_hi, __fmul_rn(x, rcp_lo));
}

}  // namespace dh
