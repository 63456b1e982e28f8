// Counter-based RNG and keyed bijection shared by the fault-injection kernels (flip.cu).
// The same functions are restated in plain C in oracle/bnn_oracle.c so masks can be checked
// bit for bit on the CPU.
#pragma once
#include <stdint.h>

namespace bnn {
namespace rng {

struct U4 {
    uint32_t x, y, z, w;
};

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1).
__host__ __device__ inline U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return U4{c0, c1, c2, c3};
}

__host__ __device__ inline uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu;
    h ^= h >> 13; h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}

struct FeistelKey {
    uint32_t k[6];
};

// Smallest bit count (>= 2) whose range covers [0, n): the walk below accepts more than half of the
// domain.  (Round 1 used the smallest EVEN count; for the 6272 candidate bits of an MNIST row that was
// 2^14, 38 % acceptance, and a warp pays the longest walk of its 32 lanes.)
__host__ __device__ inline int feistel_bits(uint64_t n) {
    int b = 2;
    while (b < 64 && ((uint64_t)1 << b) < n) ++b;
    return b;
}

__host__ __device__ inline uint32_t low_mask(int bits) {
    return bits >= 32 ? 0xFFFFFFFFu : (((uint32_t)1 << bits) - 1u);
}

// One application of the keyed bijection F on [0, 2^bits): 6 Feistel rounds on a (high, low) split of
// ceil(bits/2) and floor(bits/2) bits.  Each round maps (L, R) -> (R, L ^ f(R)) with f cut to the width of
// L, so the two widths swap every round and are back after the even number of rounds; for even `bits`
// this is the ordinary balanced network.
__host__ __device__ inline uint64_t feistel_once(uint64_t x, int bits, const FeistelKey& key) {
    int wr = bits >> 1, wl = bits - wr;  // widths of R (low part) and L (high part)
    uint32_t L = (uint32_t)(x >> wr), R = (uint32_t)x & low_mask(wr);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const uint32_t f = fmix32(R ^ key.k[r]) & low_mask(wl);
        const uint32_t nl = R;
        R = L ^ f;
        L = nl;
        const int t = wl; wl = wr; wr = t;
    }
    return ((uint64_t)L << wr) | R;
}

// Keyed bijection on [0, n): F with cycle walking (apply F until the value falls below n).
__host__ __device__ inline uint64_t feistel_perm(uint64_t x, uint64_t n, int bits,
                                                 const FeistelKey& key) {
    do {
        x = feistel_once(x, bits, key);
    } while (x >= n);
    return x;
}

// Stream ids keep the Philox counter spaces of the different consumers disjoint.
constexpr uint32_t kStreamExactK = 0xE7AC0001u;
constexpr uint32_t kStreamImage = 0x1A6E0001u;
constexpr uint32_t kStreamF32Elem = 0xF32E0001u;
constexpr uint32_t kStreamF32Bits = 0xF32B0001u;

__host__ __device__ inline FeistelKey feistel_key(uint64_t seed, uint32_t stream, uint32_t trial,
                                                  uint32_t item) {
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const U4 a = philox4x32_10(item, 0u, stream, trial, k0, k1);
    const U4 b = philox4x32_10(item, 1u, stream, trial, k0, k1);
    FeistelKey key;
    key.k[0] = a.x; key.k[1] = a.y; key.k[2] = a.z; key.k[3] = a.w;
    key.k[4] = b.x; key.k[5] = b.y;
    return key;
}

}  // namespace rng
}  // namespace bnn
