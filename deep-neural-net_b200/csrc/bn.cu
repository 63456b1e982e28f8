// Batch-norm forward/backward, sign activation, softmax + cross-entropy, SGD and accuracy kernels.
// All HBM-bound: each reads its operands once and writes each result once.
//   A4/A5  BatchNormLayer.forward    mlpcode/layers.py:74-99
//   A6     unitstep / softmax        mlpcode/activation.py:106-113, 44-53
//   A7     hard_tanh_derivative      mlpcode/activation.py:122-131 (identity on +-1 inputs)
//   A8     BatchNormLayer.backwards  mlpcode/layers.py:101-127
//   A10    cross_entropy_loss/_derivative  mlpcode/loss.py:10-36, 39-68
#include "common.cuh"

#include <type_traits>

namespace bnn {
namespace {

template <int ZT>
struct ZLoad;
template <>
struct ZLoad<BNN_Z_F32> {
    using T = float;
    __device__ static float get(const void* z, long long i) { return __ldg((const float*)z + i); }
};
template <>
struct ZLoad<BNN_Z_S16> {
    using T = int16_t;
    __device__ static float get(const void* z, long long i) { return (float)__ldg((const int16_t*)z + i); }
};
template <>
struct ZLoad<BNN_Z_S32> {
    using T = int32_t;
    __device__ static float get(const void* z, long long i) { return (float)__ldg((const int32_t*)z + i); }
};

constexpr int CS_COLS = 128;  // columns per block (16 threads x 8)
constexpr int CS_ROWS = 256;  // rows per block

// ---------------------------------------------------------------------------------------------
// 8-column vector access.  Every HBM-bound kernel below gives a thread 8 adjacent columns of a row:
// 16-byte loads of int16 Z, 2 x 16-byte loads of fp32, 8/16-byte stores of the int8 / fp16 outputs.
// `vec` says the base pointer and pitch allow it; otherwise (and on ragged edges) elements are
// touched one by one.
// ---------------------------------------------------------------------------------------------
// int -> float without the conversion (XU) pipe: 0x4B400000 is 12582912.0f = 1.5 * 2^23, whose ulp is 1,
// so adding a small integer to its bit pattern adds it to its value exactly (|v| < 2^22).
__device__ __forceinline__ float int_to_float_exact(int v) {
    return __int_as_float(0x4B400000 + v) - 12582912.0f;
}

// Raw integer view of 8 adjacent columns (int16 / int32 Z only).
template <int ZT>
__device__ __forceinline__ void load8_zi(const void* __restrict__ Z, long long off, int valid, bool vec,
                                         int (&z)[8]) {
    using T = typename ZLoad<ZT>::T;
    const T* p = reinterpret_cast<const T*>(Z) + off;
    if (vec && valid == 8) {
        if (ZT == BNN_Z_S16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                z[2 * j] = (int)(short)(w[j] & 0xffffu);
                z[2 * j + 1] = (int)w[j] >> 16;
            }
        } else {
            const int4 a = __ldg(reinterpret_cast<const int4*>(p));
            const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
            z[0] = a.x; z[1] = a.y; z[2] = a.z; z[3] = a.w;
            z[4] = b.x; z[5] = b.y; z[6] = b.z; z[7] = b.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = j < valid ? (int)__ldg(p + j) : 0;
    }
}

template <int ZT>
__device__ __forceinline__ void load8_z(const void* __restrict__ Z, long long off, int valid, bool vec,
                                        float (&z)[8]) {
    if (ZT == BNN_Z_F32) {
        const float* p = reinterpret_cast<const float*>(Z) + off;
        if (vec && valid == 8) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
            z[0] = a.x; z[1] = a.y; z[2] = a.z; z[3] = a.w;
            z[4] = b.x; z[5] = b.y; z[6] = b.z; z[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) z[j] = j < valid ? __ldg(p + j) : 0.0f;
        }
    } else {
        int zi[8];
        load8_zi < ZT == BNN_Z_F32 ? BNN_Z_S32 : ZT > (Z, off, valid, vec, zi);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            // |z| <= K: exact through the magic constant below 2^22, a plain convert above
            z[j] = (ZT == BNN_Z_S16) ? int_to_float_exact(zi[j])
                                     : ((zi[j] > -4194304 && zi[j] < 4194304) ? int_to_float_exact(zi[j])
                                                                              : (float)zi[j]);
        }
    }
}

// Two-step form of load8_z for kernels that keep several rows in flight: the raw 16 / 32 bytes stay
// packed in registers until they are needed (vector path only; valid == 8).
struct Raw8 {
    uint4 a, b;
};
template <int ZT>
__device__ __forceinline__ Raw8 load8_raw(const void* __restrict__ Z, long long off) {
    using T = typename ZLoad<ZT>::T;
    const T* p = reinterpret_cast<const T*>(Z) + off;
    Raw8 r;
    r.a = __ldg(reinterpret_cast<const uint4*>(p));
    r.b = (ZT == BNN_Z_S16) ? make_uint4(0, 0, 0, 0) : __ldg(reinterpret_cast<const uint4*>(p) + 1);
    return r;
}
template <int ZT>
__device__ __forceinline__ void cvt8_raw(const Raw8& r, float (&z)[8]) {
    if (ZT == BNN_Z_S16) {
        const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            z[2 * j] = int_to_float_exact((int)(short)(w[j] & 0xffffu));
            z[2 * j + 1] = int_to_float_exact((int)w[j] >> 16);
        }
    } else {
        const uint32_t w[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (ZT == BNN_Z_F32) {
                z[j] = __uint_as_float(w[j]);
            } else {
                const int v = (int)w[j];
                z[j] = (v > -4194304 && v < 4194304) ? int_to_float_exact(v) : (float)v;
            }
        }
    }
}

__device__ __forceinline__ void load8_f32(const float* __restrict__ p, int valid, bool vec,
                                          float (&x)[8]) {
    if (vec && valid == 8) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
        x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = j < valid ? __ldg(p + j) : 0.0f;
    }
}

template <int ZT>
__host__ __device__ inline bool z_vec_ok(const void* Z, long long ldz) {
    const int es = ZT == BNN_Z_S16 ? 2 : 4;
    return aligned16(Z) && (ldz * es) % 16 == 0;
}

// ---------------------------------------------------------------------------------------------
// Column sums / sums of squares in fp64 (exact for integer-valued Z; order independent).
// block = 256 threads = 16 column groups (8 columns each) x 16 row lanes over a 256-row slab.
// ---------------------------------------------------------------------------------------------
template <int ZT>
__global__ void __launch_bounds__(256) colstats_kernel(const void* __restrict__ Z, int B, int N,
                                                       long long ldz, double* __restrict__ sums) {
    __shared__ double red[2][8][CS_COLS];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int c0 = blockIdx.x * CS_COLS + tx * 8;
    const int valid = min(8, N - c0);
    const int r0 = blockIdx.y * CS_ROWS;
    const int r1 = min(r0 + CS_ROWS, B);
    const bool vec = z_vec_ok<ZT>(Z, ldz);
    // Per-thread partials over <= 16 rows stay off the fp64 / conversion pipes: integers for the
    // integer-valued Z (exact: |sum| < 2^20, sum of squares in 64 bits), fp32 for the fp32 Z of
    // layer 0 (16 terms).  Only the cross-thread reduction below is fp64.
    double s[8], q[8];
    if (ZT == BNN_Z_F32) {
        float sf[8], qf[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sf[j] = qf[j] = 0.f;
        if (valid > 0) {
#pragma unroll 4
            for (int r = r0 + ty; r < r1; r += 16) {
                float z[8];
                load8_z<ZT>(Z, (long long)r * ldz + c0, valid, vec, z);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    sf[j] += z[j];
                    qf[j] = fmaf(z[j], z[j], qf[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] = (double)sf[j]; q[j] = (double)qf[j]; }
    } else {
        int si[8];
        unsigned long long qi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { si[j] = 0; qi[j] = 0ull; }
        if (valid > 0) {
#pragma unroll 4
            for (int r = r0 + ty; r < r1; r += 16) {
                int z[8];
                load8_zi < ZT == BNN_Z_F32 ? BNN_Z_S32 : ZT > (Z, (long long)r * ldz + c0, valid, vec, z);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    si[j] += z[j];
                    qi[j] += (unsigned long long)((long long)z[j] * (long long)z[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] = (double)si[j]; q[j] = (double)qi[j]; }
    }
    // lanes l and l^16 hold the same columns (rows ty and ty^1): fold, then one row per warp
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], 16);
        q[j] += __shfl_xor_sync(0xffffffffu, q[j], 16);
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 16) == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[0][warp][tx * 8 + j] = s[j];
            red[1][warp][tx * 8 + j] = q[j];
        }
    }
    __syncthreads();
    const int which = threadIdx.x >> 7, col = threadIdx.x & 127;
    double a = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += red[which][i][col];
    const int gc = blockIdx.x * CS_COLS + col;
    if (gc < N) atomicAdd(&sums[(long long)which * N + gc], a);
}

__global__ void bn_stats_finalize_kernel(const double* __restrict__ sums, int N, double count,
                                         float* __restrict__ mu, float* __restrict__ var) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double m = sums[n] / count;
    double v = sums[N + n] / count - m * m;
    if (v < 0) v = 0;
    mu[n] = (float)m;
    var[n] = (float)v;
}

// IEEE division kept out of line so the common path below stays a handful of FMAs (inlined, ptxas
// if-converts it and every element pays for it).
__device__ __noinline__ float ieee_div_slow_path(float x, float d) { return __fdiv_rn(x, d); }

// ---------------------------------------------------------------------------------------------
// BN apply + sign, 64 (rows) x 128 (columns) tiles, 8 columns per thread.  The normalisation follows
// the reference's fp32 operation order (sub, div by sqrt(var+eps), mul gamma, add beta) with explicit
// round-to-nearest intrinsics so no FMA contraction can change a sign decision.  The transposed fp16
// copy (next layer's dW operand) goes through a shared-memory tile of sign bytes and leaves as
// 16-byte stores of 8 consecutive batch rows.
// ---------------------------------------------------------------------------------------------
constexpr int AP_ROWS = 64, AP_COLS = 128;

template <int ZT>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const void* __restrict__ Z, int B, int N, long long ldz, const float* __restrict__ mu,
                const float* __restrict__ var, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, float* __restrict__ out_f32,
                long long ld_f32, int8_t* __restrict__ act_i8, long long ld_i8,
                __half* __restrict__ act_t16, long long ld_t16, uint32_t* __restrict__ act_bits,
                long long ld_bits) {
    // sign bytes [row][column]; the 4-byte word index of a row is XOR-ed with (row / 8) so that the
    // transposed read below (8 rows apart, same column) hits 8 different banks
    __shared__ __align__(16) uint32_t sg[AP_ROWS][AP_COLS / 4];
    const int b0 = blockIdx.y * AP_ROWS, n0 = blockIdx.x * AP_COLS;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int c0 = n0 + tx * 8;
    const int valid = max(0, min(8, N - c0));
    const bool zvec = z_vec_ok<ZT>(Z, ldz);
    const bool fvec = out_f32 && aligned16(out_f32) && (ld_f32 % 4 == 0);
    const bool ivec = act_i8 && ((reinterpret_cast<uintptr_t>(act_i8) & 7u) == 0) && (ld_i8 % 8 == 0);
    // (z - mu) / d with d = sqrt(var + eps) constant per column: q0 = x * RN(1/d), one FMA residual
    // correction gives the correctly rounded quotient (Markstein) unless d's significand is all ones,
    // in which case the column falls back to the IEEE division.  Bit-identical to NumPy's divide.
    float m[8], d[8], rd[8], g[8], be[8];
    uint32_t slow = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool ok = j < valid;
        m[j] = ok ? __ldg(mu + c0 + j) : 0.f;
        d[j] = ok ? __fsqrt_rn(__fadd_rn(__ldg(var + c0 + j), eps)) : 1.f;
        rd[j] = __frcp_rn(d[j]);
        const uint32_t bits = __float_as_uint(d[j]);
        const uint32_t ex = (bits >> 23) & 0xffu;
        // all-ones significand, or anything that is not a comfortably normal positive number
        if ((bits & 0x7fffffu) == 0x7fffffu || (bits >> 31) || ex < 32u || ex > 222u) slow |= 1u << j;
        g[j] = ok ? __ldg(gamma + c0 + j) : 0.f;
        be[j] = ok ? __ldg(beta + c0 + j) : 0.f;
    }
    // issue every row's load before touching any of them: 4 x 16..32 bytes in flight per thread
    float zall[AP_ROWS / 16][8];
#pragma unroll
    for (int i = 0; i < AP_ROWS / 16; ++i) {
        const int gb = b0 + ty + 16 * i;
        if (gb < B && valid > 0) load8_z<ZT>(Z, (long long)gb * ldz + c0, valid, zvec, zall[i]);
    }
#pragma unroll
    for (int i = 0; i < AP_ROWS / 16; ++i) {
        const int r = ty + 16 * i;
        const int gb = b0 + r;
        uint32_t sb = 0;  // sign bits of this thread's 8 columns
        if (gb < B && valid > 0) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x = __fsub_rn(zall[i][j], m[j]);
                float q = __fmul_rn(x, rd[j]);
                q = __fmaf_rn(__fmaf_rn(-q, d[j], x), rd[j], q);
                // the shortcut is valid for normal-range operands; tiny / huge x take the slow path too
                const uint32_t ax = __float_as_uint(x) & 0x7fffffffu;
                if (((slow >> j) & 1u) || (ax != 0u && (ax < 0x10000000u || ax > 0x6f000000u)))
                    q = ieee_div_slow_path(x, d[j]);
                o[j] = __fadd_rn(__fmul_rn(q, g[j]), be[j]);
                sb |= (uint32_t)(j < valid && o[j] >= 0.0f) << j;
            }
            if (out_f32) {
                float* p = out_f32 + (long long)gb * ld_f32 + c0;
                if (fvec && valid == 8) {
                    reinterpret_cast<float4*>(p)[0] = make_float4(o[0], o[1], o[2], o[3]);
                    reinterpret_cast<float4*>(p)[1] = make_float4(o[4], o[5], o[6], o[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < valid) p[j] = o[j];
                }
            }
            if (act_i8) {
                int8_t* p = act_i8 + (long long)gb * ld_i8 + c0;
                if (ivec && valid == 8) {
                    uint2 w;
                    w.x = w.y = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        w.x |= (((sb >> j) & 1u) ? 0x01u : 0xFFu) << (8 * j);
                        w.y |= (((sb >> (j + 4)) & 1u) ? 0x01u : 0xFFu) << (8 * j);
                    }
                    *reinterpret_cast<uint2*>(p) = w;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < valid) p[j] = ((sb >> j) & 1u) ? (int8_t)1 : (int8_t)-1;
                }
            }
        }
        if (act_bits) {
            // 4 adjacent threads own the 32 columns of one word
            uint32_t w = sb << (8 * (tx & 3));
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            if ((tx & 3) == 0 && gb < B && c0 < N) act_bits[(long long)gb * ld_bits + (c0 >> 5)] = w;
        }
        if (act_t16) {
            uint32_t w0 = 0, w1 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                w0 |= ((sb >> j) & 1u) << (8 * j);
                w1 |= ((sb >> (j + 4)) & 1u) << (8 * j);
            }
            const int sw = (r >> 3) & 7;
            sg[r][(2 * tx) ^ sw] = w0;
            sg[r][(2 * tx + 1) ^ sw] = w1;
        }
    }
    if (!act_t16) return;
    __syncthreads();
    const bool tvec = aligned16(act_t16) && (ld_t16 % 8 == 0);
    const uint8_t* sgb = reinterpret_cast<const uint8_t*>(&sg[0][0]);
#pragma unroll
    for (int k = 0; k < (AP_COLS * AP_ROWS / 8) / 256; ++k) {
        // 8 adjacent lanes write the 8 x 16 bytes of one output row: full 128-byte lines
        const int item = threadIdx.x + 256 * k;
        const int bg = item & 7, n = item >> 3;  // group of 8 batch rows, column
        const int gn = n0 + n, gb = b0 + bg * 8;
        if (gn >= N || gb >= B) continue;
        const int word = (n >> 2) ^ bg, byte = n & 3;
        uint32_t h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t lo = sgb[(bg * 8 + 2 * j) * AP_COLS + word * 4 + byte] ? 0x3C00u : 0xBC00u;
            const uint32_t hi = sgb[(bg * 8 + 2 * j + 1) * AP_COLS + word * 4 + byte] ? 0x3C00u : 0xBC00u;
            h[j] = lo | (hi << 16);
        }
        __half* p = act_t16 + (long long)gn * ld_t16 + gb;
        if (tvec && gb + 8 <= B) {
            *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
        } else {
            for (int j = 0; j < 8 && gb + j < B; ++j)
                p[j] = __ushort_as_half((unsigned short)(h[j >> 1] >> (16 * (j & 1))));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Sign thresholds.  out(z) = ((z - mu) / sqrt(var + eps)) * gamma + beta is a composition of
// correctly rounded monotone fp32 operations, hence monotone in z (non-decreasing for gamma >= 0,
// non-increasing otherwise).  So sign(out(z)) == +1  <=>  z * sgn >= thr for a per-column float
// threshold found by bisection over the ordered float bit patterns, evaluating exactly the
// reference's operation sequence.  The sign pass and the fused GEMM epilogue then need one multiply
// and one compare per element and stay bit-identical to evaluating the formula (layers.py:84-86,
// 94-97 followed by activation.py:106-113).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_to_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void bn_sign_thresholds_kernel(const float* __restrict__ mu, const float* __restrict__ var,
                                          const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float eps, int N,
                                          float* __restrict__ thr, float* __restrict__ sgn) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float m = mu[n], d = __fsqrt_rn(__fadd_rn(var[n], eps)), g = gamma[n], be = beta[n];
    const float s = g < 0.0f ? -1.0f : 1.0f;
    auto plus = [&](float w) {  // w = z * s, so this predicate is non-decreasing in w
        const float z = w * s;
        const float o = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(z, m), d), g), be);
        return o >= 0.0f;
    };
    // search domain |z| <= 2^40: far beyond any pre-activation (|z| <= K for +-1 layers) and small
    // enough that (z - mu) / d * gamma cannot overflow to inf and turn inf * 0 into NaN
    uint32_t lo = float_to_key(-1.099511627776e12f), hi = float_to_key(1.099511627776e12f);
    float t;
    if (!plus(key_to_float(hi))) {
        t = INFINITY;  // never +1 (includes NaN parameters: NaN >= 0 is false, like the reference)
    } else if (plus(key_to_float(lo))) {
        t = -INFINITY;  // always +1
    } else {
        while (hi - lo > 1u) {  // invariant: !plus(lo), plus(hi)
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (plus(key_to_float(mid))) hi = mid; else lo = mid;
        }
        t = key_to_float(hi);
    }
    thr[n] = t;
    sgn[n] = s;
}

// Sign pass on thresholds: Z -> int8 +-1, transposed fp16 +-1, packed bits.
// A thread owns an 8 (rows) x 8 (columns) block and never leaves its registers: the 64 sign bits are
// one 8x8 bit matrix, whose rows expand to the int8 operand and whose transpose (three masked
// shift-xor rounds) expands to eight consecutive batch entries of each transposed fp16 row.  Lanes are
// laid out 4 (column blocks) x 8 (row blocks), so that every access is sector-exact without shared
// memory: Z loads cover 64 contiguous bytes per row, int8 stores 32, transposed fp16 stores 128.
// Warp tile 64 x 32, block = 8 warps = 128 rows x 128 columns.
constexpr int SG_ROWS = 128, SG_COLS = 128;

__device__ __forceinline__ unsigned long long transpose_8x8_bits(unsigned long long x) {
    unsigned long long t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;  x ^= t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x ^= t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x ^= t ^ (t << 28);
    return x;
}

template <int ZT>
__global__ void __launch_bounds__(256)
bn_sign_kernel(const void* __restrict__ Z, int B, int N, long long ldz, const float* __restrict__ thr,
               const float* __restrict__ sgn, int8_t* __restrict__ act_i8, long long ld_i8,
               __half* __restrict__ act_t16, long long ld_t16, uint32_t* __restrict__ act_bits,
               long long ld_bits, int8_t* __restrict__ act_t8, int8_t* __restrict__ act_t8x127,
               long long ld_t8, uint8_t* __restrict__ act_f4, long long ld_f4_bytes) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cb = lane & 3, rb = lane >> 2;
    const int c0 = blockIdx.x * SG_COLS + (warp & 3) * 32 + cb * 8;
    const int r0 = blockIdx.y * SG_ROWS + (warp >> 2) * 64 + rb * 8;
    const int valid = max(0, min(8, N - c0));
    const int rvalid = max(0, min(8, B - r0));
    const bool zvec = z_vec_ok<ZT>(Z, ldz);
    float t[8], s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        t[j] = j < valid ? __ldg(thr + c0 + j) : INFINITY;
        s[j] = j < valid ? __ldg(sgn + c0 + j) : 1.0f;
    }
    uint32_t sb[8];  // sign bits of row i, one bit per column
    if (zvec && valid == 8 && rvalid == 8) {
        Raw8 raw[8];  // every row's load is issued before any is consumed
#pragma unroll
        for (int i = 0; i < 8; ++i) raw[i] = load8_raw<ZT>(Z, (long long)(r0 + i) * ldz + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float z[8];
            cvt8_raw<ZT>(raw[i], z);
            uint32_t b = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) b |= (uint32_t)(__fmul_rn(z[j], s[j]) >= t[j]) << j;
            sb[i] = b;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t b = 0;
            if (i < rvalid && valid > 0) {
                float z[8];
                load8_z<ZT>(Z, (long long)(r0 + i) * ldz + c0, valid, false, z);
#pragma unroll
                for (int j = 0; j < 8; ++j) b |= (uint32_t)(j < valid && __fmul_rn(z[j], s[j]) >= t[j]) << j;
            }
            sb[i] = b;
        }
    }
    if (act_i8) {
        const bool ivec = ((reinterpret_cast<uintptr_t>(act_i8) & 7u) == 0) && (ld_i8 % 8 == 0) && valid == 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i >= rvalid) break;
            int8_t* p = act_i8 + (long long)(r0 + i) * ld_i8 + c0;
            if (ivec) {
                *reinterpret_cast<uint2*>(p) = make_uint2(nibble_to_pm1_bytes(sb[i]), nibble_to_pm1_bytes(sb[i] >> 4));
            } else {
                for (int j = 0; j < valid; ++j) p[j] = ((sb[i] >> j) & 1u) ? (int8_t)1 : (int8_t)-1;
            }
        }
    }
    if (act_bits) {
        // 4 adjacent lanes own the 32 columns of one word of every row of the block
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t w = sb[i] << (8 * cb);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            if (cb == 0 && i < rvalid && c0 < N) act_bits[(long long)(r0 + i) * ld_bits + (c0 >> 5)] = w;
        }
    }
    if (act_f4 && valid > 0) {
        // E2M1 nibbles (the A operand of a BNN_DT_F4 product): 8 columns = one 32-bit word per row, 4 adjacent lanes
        // fill 16 contiguous bytes; nibbles past N stay zero (the consumer's K padding)
        const uint32_t keep = valid == 8 ? 0xFFFFFFFFu : ((1u << (4 * valid)) - 1u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i >= rvalid) break;
            *reinterpret_cast<uint32_t*>(act_f4 + (long long)(r0 + i) * ld_f4_bytes + (c0 >> 1)) =
                bits_to_e2m1x8(sb[i]) & keep;
        }
    }
    if (act_t16 || act_t8) {
        unsigned long long m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) m |= (unsigned long long)(sb[i] & 0xFFu) << (8 * i);
        m = transpose_8x8_bits(m);  // byte j: column j, bit i: row i
        if (act_t8) {
            // transposed int8 copies, +-1 and +-127: the A pieces of the digit-split int8 dW product
            const bool t8vec = ((reinterpret_cast<uintptr_t>(act_t8) | reinterpret_cast<uintptr_t>(act_t8x127)) & 7u) == 0 &&
                               (ld_t8 % 8 == 0) && rvalid == 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j >= valid) break;
                const uint32_t cbits = (uint32_t)(m >> (8 * j)) & 0xFFu;
                int8_t* p1 = act_t8 + (long long)(c0 + j) * ld_t8 + r0;
                int8_t* p127 = act_t8x127 + (long long)(c0 + j) * ld_t8 + r0;
                if (t8vec) {
                    *reinterpret_cast<uint2*>(p1) = make_uint2(nibble_to_pm1_bytes(cbits), nibble_to_pm1_bytes(cbits >> 4));
                    *reinterpret_cast<uint2*>(p127) = make_uint2(nibble_to_pm127_bytes(cbits), nibble_to_pm127_bytes(cbits >> 4));
                } else {
                    for (int i = 0; i < rvalid; ++i) {
                        p1[i] = ((cbits >> i) & 1u) ? (int8_t)1 : (int8_t)-1;
                        p127[i] = ((cbits >> i) & 1u) ? (int8_t)kDwMidRadix : (int8_t)-kDwMidRadix;
                    }
                }
            }
        }
    }
    if (act_t16) {
        unsigned long long m = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) m |= (unsigned long long)(sb[i] & 0xFFu) << (8 * i);
        m = transpose_8x8_bits(m);  // byte j: column j, bit i: row i
        const bool tvec = aligned16(act_t16) && (ld_t16 % 8 == 0) && rvalid == 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j >= valid) break;
            const uint32_t cbits = (uint32_t)(m >> (8 * j)) & 0xFFu;
            __half* p = act_t16 + (long long)(c0 + j) * ld_t16 + r0;
            if (tvec) {
                *reinterpret_cast<uint4*>(p) =
                    make_uint4(pair_to_pm1_halves(cbits), pair_to_pm1_halves(cbits >> 2),
                               pair_to_pm1_halves(cbits >> 4), pair_to_pm1_halves(cbits >> 6));
            } else {
                for (int i = 0; i < rvalid; ++i)
                    p[i] = __ushort_as_half(((cbits >> i) & 1u) ? (unsigned short)0x3C00 : (unsigned short)0xBC00);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// softmax + clipped cross-entropy + (p - y)/count, one thread per row (C <= 32).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
softmax_xent_kernel(const float* __restrict__ logits, int B, int C, long long ldl,
                    const int8_t* __restrict__ onehot, float* __restrict__ probs,
                    float* __restrict__ loss_rows, float* __restrict__ grad, float count,
                    double* __restrict__ loss_sum) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    float L = 0.0f;
    if (b < B) {
        float v[32];
        float mx = -INFINITY;
        for (int c = 0; c < C; ++c) {
            v[c] = logits[(long long)b * ldl + c];
            mx = fmaxf(mx, v[c]);
        }
        float sum = 0.0f;
        for (int c = 0; c < C; ++c) {
            v[c] = expf(v[c] - mx);
            sum += v[c];
        }
        for (int c = 0; c < C; ++c) {
            float p = __fdiv_rn(v[c], sum);
            // np.clip(ypred, eps, 1 - eps, out=ypred) with eps = 1e-15: the upper bound rounds to
            // 1.0f, only the lower clip acts (loss.py:32).
            p = fminf(fmaxf(p, 1e-15f), 1.0f);
            if (probs) probs[(long long)b * C + c] = p;
            if (onehot) {
                const float y = (float)onehot[(long long)b * C + c];
                L -= y * logf(p);
                if (grad) grad[(long long)b * C + c] = __fdiv_rn(p - y, count);
            }
        }
        if (loss_rows) loss_rows[b] = L;
    }
    if (loss_sum) {
        double s = warp_sum((double)L);
        if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(loss_sum, s);
    }
}

// ---------------------------------------------------------------------------------------------
// BN backward, pass 1: column reductions of delta and delta*(Z-mu) plus magnitude bounds.
// Same 16 x 16 thread layout as colstats_kernel, 8 columns per thread.
// ---------------------------------------------------------------------------------------------
template <int ZT>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ delta, long long ldd, const void* __restrict__ Z,
                     long long ldz, int B, int N, const float* __restrict__ mu,
                     double* __restrict__ red, uint32_t* __restrict__ red_max) {
    __shared__ double sred[2][8][CS_COLS];
    __shared__ float smax[2][8][CS_COLS];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int c0 = blockIdx.x * CS_COLS + tx * 8;
    const int valid = max(0, min(8, N - c0));
    const int r0 = blockIdx.y * CS_ROWS;
    const int r1 = min(r0 + CS_ROWS, B);
    const bool zvec = z_vec_ok<ZT>(Z, ldz);
    const bool dvec = aligned16(delta) && (ldd % 4 == 0);
    float m[8], md[8], mx[8], f1[8], f2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        m[j] = j < valid ? __ldg(mu + c0 + j) : 0.f;
        md[j] = mx[j] = f1[j] = f2[j] = 0.f;
    }
    if (valid > 0) {
        // <= 16 rows per thread: fp32 partials (no fp64 / conversion-pipe work per element); the
        // cross-thread and cross-block reduction below is fp64
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += 16) {
            float z[8], dl[8];
            load8_f32(delta + (long long)r * ldd + c0, valid, dvec, dl);
            load8_z<ZT>(Z, (long long)r * ldz + c0, valid, zvec, z);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xm = z[j] - m[j];
                f1[j] += dl[j];
                f2[j] = fmaf(dl[j], xm, f2[j]);
                md[j] = fmaxf(md[j], fabsf(dl[j]));
                mx[j] = fmaxf(mx[j], j < valid ? fabsf(xm) : 0.f);
            }
        }
    }
    double s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = (double)f1[j]; s2[j] = (double)f2[j]; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
        s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
        md[j] = fmaxf(md[j], __shfl_xor_sync(0xffffffffu, md[j], 16));
        mx[j] = fmaxf(mx[j], __shfl_xor_sync(0xffffffffu, mx[j], 16));
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 16) == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sred[0][warp][tx * 8 + j] = s1[j];
            sred[1][warp][tx * 8 + j] = s2[j];
            smax[0][warp][tx * 8 + j] = md[j];
            smax[1][warp][tx * 8 + j] = mx[j];
        }
    }
    __syncthreads();
    const int which = threadIdx.x >> 7, col = threadIdx.x & 127;
    double a = 0;
    float mm = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a += sred[which][i][col];
        mm = fmaxf(mm, smax[which][i][col]);
    }
    const int gc = blockIdx.x * CS_COLS + col;
    if (gc < N) {
        atomicAdd(&red[(long long)which * N + gc], a);
        if (mm > 0.f) atomicMax(&red_max[(long long)which * N + gc], __float_as_uint(mm));
    }
}

// ---------------------------------------------------------------------------------------------
// BN backward, finalize: coefficients of delta' = c0*delta + c1*(Z-mu) + c2, in-place SGD on
// gamma/beta, running statistics (layers.py:104-123).  One thread per column, fp64 arithmetic.
// ---------------------------------------------------------------------------------------------
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ red,
                                       const uint32_t* __restrict__ red_max, int N, double count,
                                       const float* __restrict__ mu, const float* __restrict__ var,
                                       float eps, float* __restrict__ gamma, float* __restrict__ beta,
                                       float* __restrict__ mu_run, float* __restrict__ sigma_run,
                                       float lr, float ma, float one_minus_ma,
                                       const double* __restrict__ sums, float* __restrict__ coef,
                                       uint32_t* __restrict__ bound_bits,
                                       float* __restrict__ dgamma_out, float* __restrict__ dbeta_out,
                                       float* __restrict__ col_quant) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    float bound = 0.f;
    if (n < N) {
        const double S1 = red[n], S2 = red[N + n];
        const double g = (double)gamma[n];
        const float varf = var[n];
        const double sinv = 1.0 / sqrt((double)__fadd_rn(varf, eps));
        // dVar = -0.5 * sum(dZnorm * xMu) * stdInv^3, dZnorm = delta * gamma
        const double dVar = -0.5 * g * S2 * sinv * sinv * sinv;
        // mean(-2 xMu) = -2 (mean(Z) - mu): only the rounding residue of mu survives
        const double resid = sums ? (sums[n] / count - (double)mu[n]) : 0.0;
        const double dMu = -g * sinv * S1 + dVar * (-2.0 * resid);
        const double c0 = g * sinv, c1 = 2.0 * dVar / count, c2 = dMu / count;
        // delta' = c0*delta + c1*(Z - mu) + c2 is applied as c0*delta + c1*Z + (c2 - c1*mu)
        coef[n] = (float)c0;
        coef[N + n] = (float)c1;
        coef[2 * N + n] = (float)(c2 - c1 * (double)mu[n]);
        const float dGamma = (float)(S2 * sinv);  // sum(delta * znorm)
        const float dBeta = (float)S1;
        if (dgamma_out) dgamma_out[n] = dGamma;
        if (dbeta_out) dbeta_out[n] = dBeta;
        gamma[n] = __fsub_rn(gamma[n], __fmul_rn(lr, dGamma));
        beta[n] = __fsub_rn(beta[n], __fmul_rn(lr, dBeta));
        mu_run[n] = __fadd_rn(__fmul_rn(ma, mu_run[n]), __fmul_rn(one_minus_ma, mu[n]));
        sigma_run[n] = __fadd_rn(__fmul_rn(ma, sigma_run[n]), __fmul_rn(one_minus_ma, varf));
        const float amax_d = __uint_as_float(red_max[n]);
        const float amax_x = __uint_as_float(red_max[N + n]);
        bound = fabsf((float)c0) * amax_d + fabsf((float)c1) * amax_x + fabsf((float)c2);
        bound *= 1.0009765625f;  // absorb fp32 rounding of the bound itself
        if (col_quant) {
            // fixed-point scale of this column for the int8 dW product: |delta' * s| <= kDwQuantMax, so the
            // rounded value splits into the digits q1 + 256 * (q2 + 127 * q3) (common.cuh).  The scale uses
            // the whole range (not a power of two): 1 / s is rounded once, a relative 6e-8 per column.
            float q = 1.0f;
            if (bound > 0.f && isfinite(bound)) {
                q = __fdiv_rn(kDwQuantMax, bound);
                if (!(q < 1e30f)) q = 1e30f;                   // denormal-sized gradients
            }
            col_quant[n] = q;
            col_quant[N + n] = __fdiv_rn(1.0f, q);
        }
    }
    bound = warp_max(bound);
    if ((threadIdx.x & 31) == 0 && bound > 0.f && isfinite(bound))
        atomicMax(bound_bits, __float_as_uint(bound));
}

// ---------------------------------------------------------------------------------------------
// BN backward, pass 2: delta' and its fp16 hi/lo split in both operand layouts.  64 x 128 tiles,
// 8 columns per thread; the transposed pieces go through shared memory and leave as 16-byte stores.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 pack8_half(const __half (&h)[8]) {
    uint4 v;
    v.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
    v.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
    v.z = (uint32_t)__half_as_ushort(h[4]) | ((uint32_t)__half_as_ushort(h[5]) << 16);
    v.w = (uint32_t)__half_as_ushort(h[6]) | ((uint32_t)__half_as_ushort(h[7]) << 16);
    return v;
}

// Tile of fp16 pieces [row][column] in shared memory; the 16-byte chunk index of a row is XOR-ed with
// (row / 8) & 7 so the transposed read (8 rows apart, same column) is bank-conflict free.
__device__ __forceinline__ int piece_index(int row, int col) {
    return row * AP_COLS + ((((col >> 3) ^ (row >> 3)) & 15) << 3) + (col & 7);
}

// Write a tile's transposed pieces: 8 adjacent lanes cover the 8 x 16 bytes of one output row.
__device__ __forceinline__ void store_transposed_pieces(const __half* sh, const __half* sl, int b0,
                                                        int n0, int B, int N, __half* __restrict__ dt_hi,
                                                        __half* __restrict__ dt_lo, long long ld_t) {
    const bool tvec = aligned16(dt_hi) && aligned16(dt_lo) && (ld_t % 8 == 0);
#pragma unroll
    for (int k = 0; k < (AP_COLS * AP_ROWS / 8) / 256; ++k) {
        const int item = threadIdx.x + 256 * k;
        const int bg = item & 7, n = item >> 3;
        const int gn = n0 + n, gb = b0 + bg * 8;
        if (gn >= N || gb >= B) continue;
        __half h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int idx = piece_index(bg * 8 + j, n);
            h[j] = sh[idx];
            l[j] = sl[idx];
        }
        __half* ph = dt_hi + (long long)gn * ld_t + gb;
        __half* pl = dt_lo + (long long)gn * ld_t + gb;
        if (tvec && gb + 8 <= B) {
            *reinterpret_cast<uint4*>(ph) = pack8_half(h);
            *reinterpret_cast<uint4*>(pl) = pack8_half(l);
        } else {
            for (int j = 0; j < 8 && gb + j < B; ++j) {
                ph[j] = h[j];
                pl[j] = l[j];
            }
        }
    }
}

template <int ZT>
__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_kernel(const float* __restrict__ delta, long long ldd, const void* __restrict__ Z,
                    long long ldz, int B, int N, const float* __restrict__ mu,
                    const float* __restrict__ coef, const float* __restrict__ bound,
                    float* __restrict__ out_f32, long long ld_f32, __half* __restrict__ d_hi,
                    __half* __restrict__ d_lo, long long ld_rm, __half* __restrict__ dt_hi,
                    __half* __restrict__ dt_lo, long long ld_t, float* __restrict__ scale_out) {
    __shared__ __align__(16) __half sh[AP_ROWS * AP_COLS];
    __shared__ __align__(16) __half sl[AP_ROWS * AP_COLS];
    const float s = split_scale_from_bound(__ldg(bound));
    if (scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *scale_out = s;
    const int b0 = blockIdx.y * AP_ROWS, n0 = blockIdx.x * AP_COLS;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int c0 = n0 + tx * 8;
    const int valid = max(0, min(8, N - c0));
    const bool zvec = z_vec_ok<ZT>(Z, ldz);
    const bool dvec = aligned16(delta) && (ldd % 4 == 0);
    const bool fvec = out_f32 && aligned16(out_f32) && (ld_f32 % 4 == 0);
    const bool rvec = d_hi && aligned16(d_hi) && aligned16(d_lo) && (ld_rm % 8 == 0);
    float k0[8], k1[8], k2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool ok = j < valid;
        k0[j] = ok ? __ldg(coef + c0 + j) : 0.f;
        k1[j] = ok ? __ldg(coef + N + c0 + j) : 0.f;
        k2[j] = ok ? __ldg(coef + 2 * N + c0 + j) : 0.f;
    }
    // two rows at a time: 2 x 48 bytes of loads in flight per thread at 3 blocks per SM; Z stays
    // packed (4 registers per row) until it is used
    const bool fast = zvec && dvec && valid == 8;
#pragma unroll
    for (int i0 = 0; i0 < AP_ROWS / 16; i0 += 2) {
        Raw8 zraw[2];
        float dall[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int gb = b0 + ty + 16 * (i0 + u);
            if (fast && gb < B) {
                load8_f32(delta + (long long)gb * ldd + c0, 8, true, dall[u]);
                zraw[u] = load8_raw<ZT>(Z, (long long)gb * ldz + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int r = ty + 16 * (i0 + u);
            const int gb = b0 + r;
            __half h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = l[j] = __ushort_as_half(0);
            if (gb < B && valid > 0) {
                float nd[8], z[8];
                if (fast) {
                    cvt8_raw<ZT>(zraw[u], z);
                } else {  // ragged edge / unaligned pitch
                    load8_f32(delta + (long long)gb * ldd + c0, valid, false, dall[u]);
                    load8_z<ZT>(Z, (long long)gb * ldz + c0, valid, false, z);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    nd[j] = fmaf(k0[j], dall[u][j], fmaf(k1[j], z[j], k2[j]));
                    split_f16(nd[j] * s, h[j], l[j]);
                }
                if (out_f32) {
                    float* p = out_f32 + (long long)gb * ld_f32 + c0;
                    if (fvec && valid == 8) {
                        reinterpret_cast<float4*>(p)[0] = make_float4(nd[0], nd[1], nd[2], nd[3]);
                        reinterpret_cast<float4*>(p)[1] = make_float4(nd[4], nd[5], nd[6], nd[7]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < valid) p[j] = nd[j];
                    }
                }
                if (d_hi) {
                    __half* ph = d_hi + (long long)gb * ld_rm + c0;
                    __half* pl = d_lo + (long long)gb * ld_rm + c0;
                    if (rvec && valid == 8) {
                        *reinterpret_cast<uint4*>(ph) = pack8_half(h);
                        *reinterpret_cast<uint4*>(pl) = pack8_half(l);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < valid) { ph[j] = h[j]; pl[j] = l[j]; }
                    }
                }
            }
            if (dt_hi) {
                *reinterpret_cast<uint4*>(&sh[piece_index(r, tx * 8)]) = pack8_half(h);
                *reinterpret_cast<uint4*>(&sl[piece_index(r, tx * 8)]) = pack8_half(l);
            }
        }
    }
    if (!dt_hi) return;
    __syncthreads();
    store_transposed_pieces(sh, sl, b0, n0, B, N, dt_hi, dt_lo, ld_t);
}

// ---------------------------------------------------------------------------------------------
// BN backward, pass 2, fixed-point form: delta' = c0*delta + c1*Z + c2' as
//   * fp16 hi/lo pieces, row-major (the dX GEMM's A operand), and
//   * three signed-digit int8 matrices, TRANSPOSED [N, B] (the int8 dW GEMM's B pieces):
//       q = rn(delta' * s_n) = q1 + 256 * (q2 + 127 * q3),  q1, q3 in [-128, 127], q2 in [-63, 63].
// Register-only: a thread owns 8 rows x 4 columns, lanes are laid out 8 (column blocks) x 4 (row
// blocks), so delta loads cover 128 contiguous bytes per row, Z loads and row-major stores 64, and
// the transposed 8-byte digit stores 32 per column — all sector-exact, no shared memory.
// Warp tile 32 x 32, block = 8 warps = 128 rows x 64 columns.
// ---------------------------------------------------------------------------------------------
constexpr int QA_ROWS = 128, QA_COLS = 64;

template <int ZT>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_q_kernel(const float* __restrict__ delta, long long ldd, const void* __restrict__ Z,
                      long long ldz, int B, int N, const float* __restrict__ coef,
                      const float* __restrict__ bound, const float* __restrict__ col_quant,
                      __half* __restrict__ d_hi, __half* __restrict__ d_lo, long long ld_rm,
                      int8_t* __restrict__ q1t, int8_t* __restrict__ q3t, int8_t* __restrict__ q2t,
                      long long ld_q, float* __restrict__ scale_out) {
    using ZE = typename ZLoad<ZT>::T;
    const float s = split_scale_from_bound(__ldg(bound));
    if (scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *scale_out = s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cb = lane & 7, rb = lane >> 3;
    const int c0 = blockIdx.x * QA_COLS + (warp & 1) * 32 + cb * 4;
    const int r0 = blockIdx.y * QA_ROWS + (warp >> 1) * 32 + rb * 8;
    const int valid = max(0, min(4, N - c0));
    const int rvalid = max(0, min(8, B - r0));
    if (valid == 0 || rvalid == 0) return;
    const bool fast = z_vec_ok<ZT>(Z, ldz) && aligned16(delta) && (ldd % 4 == 0) && valid == 4 && rvalid == 8;
    float k0[4], k1[4], k2[4], qs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool ok = j < valid;
        k0[j] = ok ? __ldg(coef + c0 + j) : 0.f;
        k1[j] = ok ? __ldg(coef + N + c0 + j) : 0.f;
        k2[j] = ok ? __ldg(coef + 2 * N + c0 + j) : 0.f;
        qs[j] = ok ? __ldg(col_quant + c0 + j) : 1.f;
    }
    // digit bytes of column j over the 8 rows: two 32-bit words per digit matrix
    uint32_t w1[4][2], w2[4][2], w3[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) w1[j][0] = w1[j][1] = w2[j][0] = w2[j][1] = w3[j][0] = w3[j][1] = 0u;
    const bool rvec = d_hi && ((reinterpret_cast<uintptr_t>(d_hi) | reinterpret_cast<uintptr_t>(d_lo)) & 7u) == 0 &&
                      (ld_rm % 4 == 0) && valid == 4;
    float4 dv[8];
    float zv[8][4];
    if (fast) {  // every load of the block is issued before any is consumed
#pragma unroll
        for (int i = 0; i < 8; ++i) dv[i] = __ldg(reinterpret_cast<const float4*>(delta + (long long)(r0 + i) * ldd + c0));
        if (ZT == BNN_Z_S16) {
            uint2 zr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                zr[i] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const ZE*>(Z) + (long long)(r0 + i) * ldz + c0));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                zv[i][0] = int_to_float_exact((int)(short)(zr[i].x & 0xffffu));
                zv[i][1] = int_to_float_exact((int)zr[i].x >> 16);
                zv[i][2] = int_to_float_exact((int)(short)(zr[i].y & 0xffffu));
                zv[i][3] = int_to_float_exact((int)zr[i].y >> 16);
            }
        } else {
            uint4 zr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                zr[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const ZE*>(Z) + (long long)(r0 + i) * ldz + c0));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t w[4] = {zr[i].x, zr[i].y, zr[i].z, zr[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    zv[i][j] = ZT == BNN_Z_F32 ? __uint_as_float(w[j]) : (float)(int)w[j];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i >= rvalid) break;
        const long long r = r0 + i;
        float dl[4], z[4];
        if (fast) {
            dl[0] = dv[i].x; dl[1] = dv[i].y; dl[2] = dv[i].z; dl[3] = dv[i].w;
#pragma unroll
            for (int j = 0; j < 4; ++j) z[j] = zv[i][j];
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                dl[j] = j < valid ? __ldg(delta + r * ldd + c0 + j) : 0.f;
                z[j] = j < valid ? ZLoad<ZT>::get(Z, r * ldz + c0 + j) : 0.f;
            }
        }
        __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float nd = fmaf(k0[j], dl[j], fmaf(k1[j], z[j], k2[j]));
            split_f16(nd * s, h[j], l[j]);
            const int q = __float2int_rn(nd * qs[j]);       // |q| <= kDwQuantMax (+ rounding)
            const int d1 = ((q + 128) & 255) - 128;          // balanced low byte
            const int r8 = (q - d1) >> 8;                    // exact, |r8| <= 16017
            // balanced radix-127 split: r8 = d2 + 127 * d3 with |d2| <= 63 (r8 / 127 is never within
            // 0.003 of a half-integer, so the fp32 rounding below picks the exact nearest multiple)
            const int d3 = __float2int_rn((float)r8 * (1.0f / kDwMidRadix));   // |d3| <= 127
            const int d2 = r8 - kDwMidRadix * d3;
            w1[j][i >> 2] |= (uint32_t)(d1 & 0xFF) << (8 * (i & 3));
            w2[j][i >> 2] |= (uint32_t)(d2 & 0xFF) << (8 * (i & 3));
            w3[j][i >> 2] |= (uint32_t)(d3 & 0xFF) << (8 * (i & 3));
        }
        if (d_hi) {
            __half* ph = d_hi + r * ld_rm + c0;
            __half* pl = d_lo + r * ld_rm + c0;
            if (rvec) {
                *reinterpret_cast<uint2*>(ph) =
                    make_uint2((uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16),
                               (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16));
                *reinterpret_cast<uint2*>(pl) =
                    make_uint2((uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16),
                               (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16));
            } else {
                for (int j = 0; j < valid; ++j) { ph[j] = h[j]; pl[j] = l[j]; }
            }
        }
    }
    const bool qvec = ((reinterpret_cast<uintptr_t>(q1t) | reinterpret_cast<uintptr_t>(q2t) |
                        reinterpret_cast<uintptr_t>(q3t)) & 7u) == 0 && (ld_q % 8 == 0) && rvalid == 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= valid) break;
        const long long off = (long long)(c0 + j) * ld_q + r0;
        if (qvec) {
            *reinterpret_cast<uint2*>(q1t + off) = make_uint2(w1[j][0], w1[j][1]);
            *reinterpret_cast<uint2*>(q2t + off) = make_uint2(w2[j][0], w2[j][1]);
            *reinterpret_cast<uint2*>(q3t + off) = make_uint2(w3[j][0], w3[j][1]);
        } else {
            for (int i = 0; i < rvalid; ++i) {
                q1t[off + i] = (int8_t)((w1[j][i >> 2] >> (8 * (i & 3))) & 0xFF);
                q2t[off + i] = (int8_t)((w2[j][i >> 2] >> (8 * (i & 3))) & 0xFF);
                q3t[off + i] = (int8_t)((w3[j][i >> 2] >> (8 * (i & 3))) & 0xFF);
            }
        }
    }
}

__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  long long n, float lr, int parts, long long part_stride) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float sum = __ldg(g + i);
        for (int s = 1; s < parts; ++s) sum = __fadd_rn(sum, __ldg(g + s * part_stride + i));
        p[i] = __fsub_rn(p[i], __fmul_rn(lr, sum));
    }
}

// rows are grouped: row r belongs to group r / B and is compared with labels[r % B] (one group per
// fault trial, all trials sharing the evaluation labels); correct[group] += hits.
__global__ void __launch_bounds__(256) count_correct_kernel(const float* __restrict__ scores, int B,
                                                            int C, long long lds,
                                                            const int32_t* __restrict__ labels,
                                                            int32_t* __restrict__ correct, int groups) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long rows = (long long)B * groups;
    int hit = 0, grp = 0;
    if (r < rows) {
        grp = (int)(r / B);
        int best = 0;
        float bv = scores[r * lds];
        for (int c = 1; c < C; ++c) {
            const float v = scores[r * lds + c];
            if (v > bv) { bv = v; best = c; }
        }
        hit = (best == labels[r % B]) ? 1 : 0;
    }
    // a warp can straddle two groups: count per group with two ballots
    const int g0 = __shfl_sync(0xffffffffu, grp, 0);
    const unsigned m0 = __ballot_sync(0xffffffffu, hit && grp == g0);
    const unsigned m1 = __ballot_sync(0xffffffffu, hit && grp != g0);
    if ((threadIdx.x & 31) == 0) {
        if (m0) atomicAdd(correct + g0, __popc(m0));
    }
    if (m1) {
        const int src = __ffs(m1) - 1;
        const int g1 = __shfl_sync(0xffffffffu, grp, src);
        if ((threadIdx.x & 31) == 0) atomicAdd(correct + g1, __popc(m1));
    }
}

template <typename F>
int dispatch_z(int z_dtype, F&& f) {
    switch (z_dtype) {
        case BNN_Z_F32: return f(std::integral_constant<int, BNN_Z_F32>{});
        case BNN_Z_S16: return f(std::integral_constant<int, BNN_Z_S16>{});
        case BNN_Z_S32: return f(std::integral_constant<int, BNN_Z_S32>{});
    }
    set_error("unknown z_dtype %d", z_dtype);
    return BNN_ERR_ARG;
}

}  // namespace
}  // namespace bnn

using namespace bnn;

extern "C" int bnn_colstats(const void* Z, int z_dtype, int B, int N, int64_t ldz, double* sums,
                            int accumulate, bnn_stream_t stream) {
    BNN_CHECK_ARG(Z && sums && B > 0 && N > 0 && ldz >= N, "bnn_colstats: bad input");
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) BNN_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * N, s));
    dim3 grid(cdiv(N, CS_COLS), cdiv(B, CS_ROWS));
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_colstats: B too large");
    return dispatch_z(z_dtype, [&](auto zt) {
        colstats_kernel<decltype(zt)::value><<<grid, 256, 0, s>>>(Z, B, N, ldz, sums);
        BNN_LAUNCH_CHECK("colstats_kernel");
        return (int)BNN_OK;
    });
}

extern "C" int bnn_bn_stats_finalize(const double* sums, int N, double count, float* mu, float* var,
                                     bnn_stream_t stream) {
    BNN_CHECK_ARG(sums && mu && var && N > 0 && count > 0, "bnn_bn_stats_finalize: bad input");
    bn_stats_finalize_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(sums, N, count, mu, var);
    BNN_LAUNCH_CHECK("bn_stats_finalize_kernel");
    return BNN_OK;
}

extern "C" int bnn_bn_apply(const void* Z, int z_dtype, int B, int N, int64_t ldz, const float* mu,
                            const float* var, const float* gamma, const float* beta, float eps,
                            float* out_f32, int64_t ld_f32, int8_t* act_i8, int64_t ld_i8,
                            void* act_t16, int64_t ld_t16, uint32_t* act_bits, int64_t ld_bits,
                            bnn_stream_t stream) {
    BNN_CHECK_ARG(Z && mu && var && gamma && beta && B > 0 && N > 0 && ldz >= N,
                  "bnn_bn_apply: bad input");
    BNN_CHECK_ARG(!out_f32 || ld_f32 >= N, "bnn_bn_apply: ld_f32 < N");
    BNN_CHECK_ARG(!act_i8 || ld_i8 >= N, "bnn_bn_apply: ld_i8 < N");
    BNN_CHECK_ARG(!act_t16 || ld_t16 >= B, "bnn_bn_apply: ld_t16 < B");
    BNN_CHECK_ARG(!act_bits || ld_bits >= (N + 31) / 32, "bnn_bn_apply: ld_bits too small");
    dim3 grid(cdiv(N, AP_COLS), cdiv(B, AP_ROWS));
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_bn_apply: B too large");
    cudaStream_t s = (cudaStream_t)stream;
    return dispatch_z(z_dtype, [&](auto zt) {
        bn_apply_kernel<decltype(zt)::value><<<grid, 256, 0, s>>>(
            Z, B, N, ldz, mu, var, gamma, beta, eps, out_f32, ld_f32, act_i8, ld_i8,
            static_cast<__half*>(act_t16), ld_t16, act_bits, ld_bits);
        BNN_LAUNCH_CHECK("bn_apply_kernel");
        return (int)BNN_OK;
    });
}

extern "C" int bnn_bn_sign_thresholds(const float* mu, const float* var, const float* gamma,
                                      const float* beta, float eps, int N, float* thr, float* sgn,
                                      bnn_stream_t stream) {
    BNN_CHECK_ARG(mu && var && gamma && beta && thr && sgn && N > 0, "bnn_bn_sign_thresholds: bad input");
    bn_sign_thresholds_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(mu, var, gamma, beta, eps,
                                                                              N, thr, sgn);
    BNN_LAUNCH_CHECK("bn_sign_thresholds_kernel");
    return BNN_OK;
}

extern "C" int bnn_bn_sign(const void* Z, int z_dtype, int B, int N, int64_t ldz, const float* thr,
                           const float* sgn, int8_t* act_i8, int64_t ld_i8, void* act_t16,
                           int64_t ld_t16, uint32_t* act_bits, int64_t ld_bits, int8_t* act_t8,
                           int8_t* act_t8x127, int64_t ld_t8, void* act_f4, int64_t ld_f4,
                           bnn_stream_t stream) {
    BNN_CHECK_ARG((act_t8 == nullptr) == (act_t8x127 == nullptr) && (!act_t8 || ld_t8 >= B),
                  "bnn_bn_sign: act_t8 / act_t8x127 come as a pair with ld_t8 >= B");
    BNN_CHECK_ARG(Z && thr && sgn && B > 0 && N > 0 && ldz >= N, "bnn_bn_sign: bad input");
    BNN_CHECK_ARG(!act_i8 || ld_i8 >= N, "bnn_bn_sign: ld_i8 < N");
    BNN_CHECK_ARG(!act_t16 || ld_t16 >= B, "bnn_bn_sign: ld_t16 < B");
    BNN_CHECK_ARG(!act_bits || ld_bits >= (N + 31) / 32, "bnn_bn_sign: ld_bits too small");
    BNN_CHECK_ARG(!act_f4 || (aligned16(act_f4) && ld_f4 % 32 == 0 && ld_f4 >= ((N + 31) / 32) * 32),
                  "bnn_bn_sign: act_f4 must be 16-byte aligned with a pitch of whole 32-element words");
    dim3 grid(cdiv(N, SG_COLS), cdiv(B, SG_ROWS));
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_bn_sign: B too large");
    cudaStream_t s = (cudaStream_t)stream;
    return dispatch_z(z_dtype, [&](auto zt) {
        bn_sign_kernel<decltype(zt)::value><<<grid, 256, 0, s>>>(
            Z, B, N, ldz, thr, sgn, act_i8, ld_i8, static_cast<__half*>(act_t16), ld_t16, act_bits,
            ld_bits, act_t8, act_t8x127, ld_t8, static_cast<uint8_t*>(act_f4), ld_f4 / 2);
        BNN_LAUNCH_CHECK("bn_sign_kernel");
        return (int)BNN_OK;
    });
}

extern "C" int bnn_softmax_xent(const float* logits, int B, int C, int64_t ldl, const int8_t* onehot,
                                float* probs, float* loss_rows, float* grad, double count,
                                double* loss_sum, bnn_stream_t stream) {
    BNN_CHECK_ARG(logits && B > 0 && C > 0 && C <= 32 && ldl >= C, "bnn_softmax_xent: bad input");
    BNN_CHECK_ARG(onehot || (!loss_rows && !grad && !loss_sum),
                  "bnn_softmax_xent: loss/grad need the one-hot labels");
    cudaStream_t s = (cudaStream_t)stream;
    if (loss_sum) BNN_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), s));
    softmax_xent_kernel<<<cdiv(B, 128), 128, 0, s>>>(logits, B, C, ldl, onehot, probs, loss_rows,
                                                     grad, (float)count, loss_sum);
    BNN_LAUNCH_CHECK("softmax_xent_kernel");
    return BNN_OK;
}

extern "C" int bnn_bn_bwd_reduce(const float* delta, int64_t ldd, const void* Z, int z_dtype,
                                 int64_t ldz, int B, int N, const float* mu, double* red,
                                 uint32_t* red_max, int accumulate, bnn_stream_t stream) {
    BNN_CHECK_ARG(delta && Z && mu && red && red_max && B > 0 && N > 0 && ldd >= N && ldz >= N,
                  "bnn_bn_bwd_reduce: bad input");
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate) {
        BNN_CUDA(cudaMemsetAsync(red, 0, sizeof(double) * 2 * N, s));
        BNN_CUDA(cudaMemsetAsync(red_max, 0, sizeof(uint32_t) * 2 * N, s));
    }
    dim3 grid(cdiv(N, CS_COLS), cdiv(B, CS_ROWS));
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_bn_bwd_reduce: B too large");
    return dispatch_z(z_dtype, [&](auto zt) {
        bn_bwd_reduce_kernel<decltype(zt)::value><<<grid, 256, 0, s>>>(delta, ldd, Z, ldz, B, N,
                                                                         mu, red, red_max);
        BNN_LAUNCH_CHECK("bn_bwd_reduce_kernel");
        return (int)BNN_OK;
    });
}

extern "C" int bnn_bn_bwd_finalize(const double* red, const uint32_t* red_max, int N, double count,
                                   const float* mu, const float* var, float eps, float* gamma,
                                   float* beta, float* mu_run, float* sigma_run, float lr,
                                   double momentum, const double* sums, float* coef,
                                   uint32_t* bound_bits, float* dgamma_out, float* dbeta_out,
                                   float* col_quant, bnn_stream_t stream) {
    BNN_CHECK_ARG(red && red_max && mu && var && gamma && beta && mu_run && sigma_run && coef &&
                      bound_bits && N > 0 && count > 0,
                  "bnn_bn_bwd_finalize: bad input");
    cudaStream_t s = (cudaStream_t)stream;
    BNN_CUDA(cudaMemsetAsync(bound_bits, 0, sizeof(uint32_t), s));
    // (MA_BETA * mu) + ((1 - MA_BETA) * batch_mu) with Python-float constants cast to fp32
    const float ma = (float)momentum, oma = (float)(1.0 - momentum);
    bn_bwd_finalize_kernel<<<cdiv(N, 128), 128, 0, s>>>(red, red_max, N, count, mu, var, eps, gamma,
                                                        beta, mu_run, sigma_run, lr, ma, oma, sums,
                                                        coef, bound_bits, dgamma_out, dbeta_out,
                                                        col_quant);
    BNN_LAUNCH_CHECK("bn_bwd_finalize_kernel");
    return BNN_OK;
}

extern "C" int bnn_bn_bwd_apply(const float* delta, int64_t ldd, const void* Z, int z_dtype,
                                int64_t ldz, int B, int N, const float* mu, const float* coef,
                                const float* bound, float* out_f32, int64_t ld_f32, void* d_hi,
                                void* d_lo, int64_t ld_rm, void* dt_hi, void* dt_lo, int64_t ld_t,
                                float* scale_out, bnn_stream_t stream) {
    BNN_CHECK_ARG(delta && Z && mu && coef && bound && B > 0 && N > 0 && ldd >= N && ldz >= N,
                  "bnn_bn_bwd_apply: bad input");
    BNN_CHECK_ARG((d_hi == nullptr) == (d_lo == nullptr) && (dt_hi == nullptr) == (dt_lo == nullptr),
                  "bnn_bn_bwd_apply: hi/lo must be given in pairs");
    BNN_CHECK_ARG(!d_hi || ld_rm >= N, "bnn_bn_bwd_apply: ld_rm < N");
    BNN_CHECK_ARG(!dt_hi || ld_t >= B, "bnn_bn_bwd_apply: ld_t < B");
    BNN_CHECK_ARG(!out_f32 || ld_f32 >= N, "bnn_bn_bwd_apply: ld_f32 < N");
    dim3 grid(cdiv(N, AP_COLS), cdiv(B, AP_ROWS));
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_bn_bwd_apply: B too large");
    cudaStream_t s = (cudaStream_t)stream;
    return dispatch_z(z_dtype, [&](auto zt) {
        bn_bwd_apply_kernel<decltype(zt)::value><<<grid, 256, 0, s>>>(
            delta, ldd, Z, ldz, B, N, mu, coef, bound, out_f32, ld_f32, static_cast<__half*>(d_hi),
            static_cast<__half*>(d_lo), ld_rm, static_cast<__half*>(dt_hi),
            static_cast<__half*>(dt_lo), ld_t, scale_out);
        BNN_LAUNCH_CHECK("bn_bwd_apply_kernel");
        return (int)BNN_OK;
    });
}

extern "C" int bnn_bn_bwd_apply_q(const float* delta, int64_t ldd, const void* Z, int z_dtype,
                                  int64_t ldz, int B, int N, const float* coef, const float* bound,
                                  const float* col_quant, void* d_hi, void* d_lo, int64_t ld_rm,
                                  int8_t* q1t, int8_t* q3t, int8_t* q2t, int64_t ld_q,
                                  float* scale_out, bnn_stream_t stream) {
    BNN_CHECK_ARG(delta && Z && coef && bound && col_quant && q1t && q2t && q3t && B > 0 && N > 0 &&
                      ldd >= N && ldz >= N && ld_q >= B,
                  "bnn_bn_bwd_apply_q: bad input");
    BNN_CHECK_ARG((d_hi == nullptr) == (d_lo == nullptr) && (!d_hi || ld_rm >= N),
                  "bnn_bn_bwd_apply_q: d_hi / d_lo come as a pair with ld_rm >= N");
    dim3 grid(cdiv(N, QA_COLS), cdiv(B, QA_ROWS));
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_bn_bwd_apply_q: B too large");
    cudaStream_t s = (cudaStream_t)stream;
    return dispatch_z(z_dtype, [&](auto zt) {
        bn_bwd_apply_q_kernel<decltype(zt)::value><<<grid, 256, 0, s>>>(
            delta, ldd, Z, ldz, B, N, coef, bound, col_quant, static_cast<__half*>(d_hi),
            static_cast<__half*>(d_lo), ld_rm, q1t, q3t, q2t, ld_q, scale_out);
        BNN_LAUNCH_CHECK("bn_bwd_apply_q_kernel");
        return (int)BNN_OK;
    });
}

extern "C" int bnn_sgd_update(float* p, const float* g, int64_t n, float lr, int parts,
                              int64_t part_stride, bnn_stream_t stream) {
    BNN_CHECK_ARG(p && g && n >= 0 && parts >= 1, "bnn_sgd_update: bad input");
    BNN_CHECK_ARG(parts == 1 || part_stride >= n, "bnn_sgd_update: part_stride < n");
    if (n == 0) return BNN_OK;
    const long long blocks = (n + 255) / 256;
    const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    sgd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, n, lr, parts, part_stride);
    BNN_LAUNCH_CHECK("sgd_kernel");
    return BNN_OK;
}

extern "C" int bnn_count_correct(const float* scores, int B, int C, int64_t lds,
                                 const int32_t* labels, int32_t* correct, int groups,
                                 bnn_stream_t stream) {
    BNN_CHECK_ARG(scores && labels && correct && B > 0 && C > 0 && lds >= C && groups >= 1,
                  "bnn_count_correct: bad input");
    BNN_CHECK_ARG(B >= 32 || groups == 1, "bnn_count_correct: grouped mode needs B >= 32");
    count_correct_kernel<<<cdiv((long)B * groups, 256), 256, 0, (cudaStream_t)stream>>>(
        scores, B, C, lds, labels, correct, groups);
    BNN_LAUNCH_CHECK("count_correct_kernel");
    return BNN_OK;
}
