// Per-element arithmetic of the fault-injection kernels that is pure (no memory, no threads), kept in
// __host__ __device__ functions so that the SAME source the kernels run is also compiled for the host
// by tests/test_host_math.py (g++ -x c++) and compared with the oracle without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#include "rng.cuh"

namespace bnn {
namespace fault {

__host__ __device__ inline int popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// XOR mask of the IEEE-754 bit flips for chosen element `e` holding the word `v`
// (ErrorCallback._flipFunc, callbacks.py:35-59).  Reference bit index j (np.unpackbits of the
// little-endian bytes) is word bit 8*(j/8) + 7 - j%8; candidates are ranked by ascending j.
__host__ __device__ inline uint32_t f32_flip_mask(uint32_t v, uint64_t e, double p, int mode,
                                                  uint64_t seed, uint32_t trial) {
    // exponent MSB set: the other exponent bits (29..23) are protected; otherwise the MSB (30) itself
    uint32_t cand = (v & 0x40000000u) ? ~0x3F800000u : ~0x40000000u;
    if (mode == 0) cand &= ~v;
    else if (mode == 1) cand &= v;
    const int cnt = popc32(cand);
    int kb = (int)rint((double)cnt * p);  // Python's round(): half to even
    if (kb > cnt) kb = cnt;
    if (kb <= 0) return 0u;
    const rng::FeistelKey bkey = rng::feistel_key(seed, rng::kStreamF32Bits, trial, (uint32_t)e);
    const int bbits = rng::feistel_bits((uint64_t)cnt);
    uint32_t m = 0;
    for (int j = 0; j < kb; ++j) {
        int rank = (int)rng::feistel_perm((uint64_t)j, (uint64_t)cnt, bbits, bkey);
        for (int idx = 0; idx < 32; ++idx) {
            const int pos = 8 * (idx >> 3) + 7 - (idx & 7);
            if ((cand >> pos) & 1u) {
                if (rank == 0) { m |= 1u << pos; break; }
                --rank;
            }
        }
    }
    return m;
}

// Element i of the draw of `n_sel` distinct elements out of n.
__host__ __device__ inline uint64_t f32_flip_element(uint64_t i, uint64_t n, int ebits, uint64_t seed,
                                                     uint32_t trial) {
    const rng::FeistelKey ekey = rng::feistel_key(seed, rng::kStreamF32Elem, trial, 0u);
    return rng::feistel_perm(i, n, ebits, ekey);
}

// Integer threshold of a uint8-fed layer 0 (see bnn_u8_input_thresholds in the header):
//   s*z >= tz,  z = r*(g + (bias - mn)*S) + new_min*S,  r = (new_max - new_min)/(mx - mn)
//   <=>  s*g >= ceil( (tz - s*new_min*S)/r - s*(bias - mn)*S )
__host__ __device__ inline float u8_input_threshold(float tz, double s, double S, double mn, double mx,
                                                    double new_min, double new_max, double bias) {
    if (!(mx > mn)) return INFINITY;          // the reference divides by zero here
    if (isnan(tz)) return INFINITY;           // NaN >= 0 is false in the reference too
    if (isinf(tz)) return tz;
    const double r = (new_max - new_min) / (mx - mn);
    const double tau = ((double)tz - s * new_min * S) / r - s * (bias - mn) * S;
    const double c = ceil(tau);
    return c > 1.6e7 ? INFINITY : (c < -1.6e7 ? -INFINITY : (float)c);
}

}  // namespace fault
}  // namespace bnn
