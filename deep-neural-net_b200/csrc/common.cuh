// Shared host/device helpers for libbnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/bnn_b200.h"

namespace bnn {

// ---- error reporting across the C ABI (no exceptions cross the boundary) -------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define BNN_CHECK_ARG(cond, ...)                                  \
    do {                                                          \
        if (!(cond)) {                                            \
            ::bnn::set_error(__VA_ARGS__);                        \
            return BNN_ERR_ARG;                                   \
        }                                                         \
    } while (0)

#define BNN_CUDA(call)                                            \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return ::bnn::cuda_fail(_e, #call);\
    } while (0)

#define BNN_LAUNCH_CHECK(name)                                    \
    do {                                                          \
        cudaError_t _e = cudaGetLastError();                      \
        if (_e != cudaSuccess) return ::bnn::cuda_fail(_e, name); \
    } while (0)

__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device

// ---- device helpers -------------------------------------------------------------------------
// Reference sign convention: +1 iff x >= 0 (so -0.0 -> +1, NaN -> "not >= 0" -> bit 0).
// mlpcode/layers.py:278-279, activation.py:110-111.
__device__ __forceinline__ uint32_t sign_bit_ge0(float x) { return x >= 0.0f ? 1u : 0u; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Power-of-two scale that maps |x| <= bound into the fp16 range with head-room:
// s = 2^(14 - ceil(log2(bound))), so |x*s| <= 2^14 < 65504.  bound <= 0 -> s = 1.
__device__ __forceinline__ float split_scale_from_bound(float bound) {
    if (!(bound > 0.0f) || !isfinite(bound)) return 1.0f;
    int e;
    float m = frexpf(bound, &e);  // bound = m * 2^e, m in [0.5,1)
    int ce = (m == 0.5f) ? e - 1 : e;  // ceil(log2(bound))
    int sh = 14 - ce;
    sh = sh > 120 ? 120 : (sh < -120 ? -120 : sh);
    return ldexpf(1.0f, sh);
}

// hi/lo fp16 split of a pre-scaled fp32 value: x*s = hi + lo + O(2^-22 |x*s|).
__device__ __forceinline__ void split_f16(float xs, __half& hi, __half& lo) {
    hi = __float2half_rn(xs);
    lo = __float2half_rn(xs - __half2float(hi));
}

// 4 mask bits -> 4 bytes of +-1 (bit ? 0x01 : 0xFF)
__device__ __forceinline__ uint32_t nibble_to_pm1_bytes(uint32_t nib) {
    const uint32_t spread = ((nib & 0xFu) * 0x00204081u) & 0x01010101u;
    return 0xFFFFFFFFu ^ (spread * 0xFEu);
}
// 4 mask bits -> 4 bytes of +-127 (bit ? 0x7F : 0x81)
__device__ __forceinline__ uint32_t nibble_to_pm127_bytes(uint32_t nib) {
    const uint32_t spread = ((nib & 0xFu) * 0x00204081u) & 0x01010101u;
    return 0x81818181u ^ (spread * 0xFEu);
}

// Fixed-point form of delta' for the int8 dW product (bn.cu, gemm_tc.cu hi_from_pass):
//   q = rn(delta' * s_n) = q1 + 256 * (q2 + 127 * q3),  q1, q3 in [-128, 127], q2 in [-63, 63]
// (the middle radix is 127 because the A piece that carries it, 127 * X with X = +-1, must fit int8).
// Representable: |q| <= 256 * (63 + 127 * 127) - 128 = 4 145 024; s_n = kDwQuantMax / bound_n keeps
// |q| <= kDwQuantMax (+ rounding) for |delta'| <= bound_n: 22.97 bits of the column's bound.
constexpr float kDwQuantMax = 4100000.0f;
constexpr int kDwMidRadix = 127;
// 8 bits -> bit 0 of 8 nibbles (bit j -> bit 4j): three shift-or-mask steps
__host__ __device__ __forceinline__ uint32_t spread_bits_to_nibbles(uint32_t b8) {
    uint32_t x = b8 & 0xFFu;
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
}
// 8 sign bits (1 = +1) -> 8 E2M1 nibbles: +1.0 = 0x2, -1.0 = 0xA
__host__ __device__ __forceinline__ uint32_t bits_to_e2m1x8(uint32_t b8) {
    return 0xAAAAAAAAu ^ (spread_bits_to_nibbles(b8) << 3);
}
// 2 mask bits -> 2 halves of +-1.0 (bit ? 0x3C00 : 0xBC00)
__device__ __forceinline__ uint32_t pair_to_pm1_halves(uint32_t two) {
    return 0x3C003C00u | ((~two & 1u) << 15) | ((~two & 2u) << 30);
}

}  // namespace bnn
