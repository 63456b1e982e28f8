// sign() binarization, bit packing and operand-layout kernels (all HBM-bound).
//   A1  BinaryLayer.binarize   mlpcode/layers.py:275-281   (+1 iff x >= 0)
//   A6  unitstep               mlpcode/activation.py:106-113
// plus the fp32 -> 2 x fp16 split that feeds the real-valued tensor-core GEMMs.
#include "common.cuh"

namespace bnn {
namespace {

// ---------------------------------------------------------------------------------------------
// Row packing: one warp produces one 32-bit word per ballot; coalesced 128-byte reads.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ bool is_plus(T v);
template <>
__device__ __forceinline__ bool is_plus<float>(float v) { return v >= 0.0f; }
template <>
__device__ __forceinline__ bool is_plus<int8_t>(int8_t v) { return v > 0; }

template <typename T>
__global__ void __launch_bounds__(256) pack_rows_kernel(const T* __restrict__ x, int R, int C,
                                                        long long ldx, uint32_t* __restrict__ bits,
                                                        long long ldw) {
    const int words = (C + 31) >> 5;
    const long long total = (long long)R * words;
    const int lane = threadIdx.x & 31;
    long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long wstride = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; w < total; w += wstride) {  // warp-uniform
        const int r = (int)(w / words), cw = (int)(w % words);
        const int c = cw * 32 + lane;
        const bool p = (c < C) && is_plus<T>(x[(long long)r * ldx + c]);
        const uint32_t word = __ballot_sync(0xffffffffu, p);
        if (lane == 0) bits[(long long)r * ldw + cw] = word;
    }
}

// ---------------------------------------------------------------------------------------------
// Transposing weight binarization: W[K,N] fp32 -> K-major +-1 copies [N, K] and packed bits.
// A warp owns 32 (k) x 128 (n): every row is read as one 512-byte segment (float4 per lane), and
// each lane folds the sign of its 4 columns into four 32-bit masks over k — which ARE the packed
// words wt_bits[n][k/32].  No transpose through shared memory is needed: the masks are expanded in
// registers into the 32 consecutive int8 / fp16 values of output row n and leave as 16-byte stores.
// Block = 8 warps = 128 (k) x 256 (n).
// ---------------------------------------------------------------------------------------------
constexpr int BT_K = 128, BT_N = 256;

__global__ void __launch_bounds__(256)
binarize_t_kernel(const float* __restrict__ W, int K, int N, int8_t* __restrict__ wt_i8,
                  long long ld_i8, __half* __restrict__ wt_f16, long long ld_f16,
                  uint32_t* __restrict__ wt_bits, long long ld_bits,
                  uint8_t* __restrict__ wt_f4, long long ld_f4_bytes,
                  uint32_t* __restrict__ absmax_bits) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k0 = blockIdx.y * BT_K + (warp & 3) * 32;
    const int n0 = blockIdx.x * BT_N + (warp >> 2) * 128 + lane * 4;
    if (k0 >= K) return;  // warp-uniform
    const bool vec = (N % 4 == 0) && aligned16(W);
    const int nvalid = max(0, min(4, N - n0));
    const int kvalid = min(32, K - k0);
    uint32_t m[4] = {0u, 0u, 0u, 0u};
    float amax = 0.0f;
    if (nvalid > 0) {
        const float* base = W + (long long)k0 * N + n0;
        if (vec && nvalid == 4 && kvalid == 32) {
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(base + (long long)k * N));
                m[0] |= sign_bit_ge0(v.x) << k;
                m[1] |= sign_bit_ge0(v.y) << k;
                m[2] |= sign_bit_ge0(v.z) << k;
                m[3] |= sign_bit_ge0(v.w) << k;
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
            }
        } else {
            for (int k = 0; k < kvalid; ++k)
                for (int j = 0; j < nvalid; ++j) {
                    const float v = __ldg(base + (long long)k * N + j);
                    m[j] |= sign_bit_ge0(v) << k;
                    amax = fmaxf(amax, fabsf(v));
                }
        }
    }
    if (absmax_bits) {  // max|W| for the backward split comes for free with this read of W
        amax = warp_max(amax);
        // one atomic per warp only while it can still raise the maximum (plain read as a filter)
        if (lane == 0 && __float_as_uint(amax) > *reinterpret_cast<volatile uint32_t*>(absmax_bits))
            atomicMax(absmax_bits, __float_as_uint(amax));
    }
    const bool i8vec = wt_i8 && aligned16(wt_i8) && (ld_i8 % 16 == 0) && kvalid == 32;
    const bool f16vec = wt_f16 && aligned16(wt_f16) && (ld_f16 % 8 == 0) && kvalid == 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j >= nvalid) break;
        const long long n = n0 + j;
        const uint32_t bits = m[j];
        if (wt_bits) wt_bits[n * ld_bits + (k0 >> 5)] = bits;
        if (wt_f4) {   // 32 E2M1 nibbles = one 16-byte word; nibbles past K stay zero (the consumer's K padding)
            const uint32_t keep = kvalid == 32 ? 0xFFFFFFFFu : ((1u << kvalid) - 1u);
            uint4 o;
            o.x = bits_to_e2m1x8(bits) & (spread_bits_to_nibbles(keep) * 0xFu);
            o.y = bits_to_e2m1x8(bits >> 8) & (spread_bits_to_nibbles(keep >> 8) * 0xFu);
            o.z = bits_to_e2m1x8(bits >> 16) & (spread_bits_to_nibbles(keep >> 16) * 0xFu);
            o.w = bits_to_e2m1x8(bits >> 24) & (spread_bits_to_nibbles(keep >> 24) * 0xFu);
            *reinterpret_cast<uint4*>(wt_f4 + n * ld_f4_bytes + (k0 >> 1)) = o;
        }
        if (wt_i8) {
            int8_t* d = wt_i8 + n * ld_i8 + k0;
            if (i8vec) {
                uint4 a, b;
                a.x = nibble_to_pm1_bytes(bits);       a.y = nibble_to_pm1_bytes(bits >> 4);
                a.z = nibble_to_pm1_bytes(bits >> 8);  a.w = nibble_to_pm1_bytes(bits >> 12);
                b.x = nibble_to_pm1_bytes(bits >> 16); b.y = nibble_to_pm1_bytes(bits >> 20);
                b.z = nibble_to_pm1_bytes(bits >> 24); b.w = nibble_to_pm1_bytes(bits >> 28);
                reinterpret_cast<uint4*>(d)[0] = a;
                reinterpret_cast<uint4*>(d)[1] = b;
            } else {
                for (int k = 0; k < kvalid; ++k) d[k] = ((bits >> k) & 1u) ? (int8_t)1 : (int8_t)-1;
            }
        }
        if (wt_f16) {
            __half* d = wt_f16 + n * ld_f16 + k0;
            if (f16vec) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t b8 = bits >> (8 * q);
                    reinterpret_cast<uint4*>(d)[q] =
                        make_uint4(pair_to_pm1_halves(b8), pair_to_pm1_halves(b8 >> 2),
                                   pair_to_pm1_halves(b8 >> 4), pair_to_pm1_halves(b8 >> 6));
                }
            } else {
                for (int k = 0; k < kvalid; ++k)
                    d[k] = __ushort_as_half(((bits >> k) & 1u) ? (unsigned short)0x3C00
                                                               : (unsigned short)0xBC00);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
unpack_bits_kernel(const uint32_t* __restrict__ bits, int R, int C, long long ldw,
                   int8_t* __restrict__ out_i8, long long ld_i8, __half* __restrict__ out_f16,
                   long long ld_f16) {
    const long long total = (long long)R * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / C), c = (int)(i % C);
        const uint32_t b = (bits[(long long)r * ldw + (c >> 5)] >> (c & 31)) & 1u;
        if (out_i8) out_i8[(long long)r * ld_i8 + c] = b ? (int8_t)1 : (int8_t)-1;
        if (out_f16)
            out_f16[(long long)r * ld_f16 + c] =
                __ushort_as_half(b ? (unsigned short)0x3C00 : (unsigned short)0xBC00);
    }
}

// Row gather: dst[i, :] = src[idx[i], :] on raw bytes (any element type).  The mini-batch assembly of an epoch
// (trainX[p, :][k:k+bs], network.py:312-315, 480) without materialising the shuffled data set: idx is the slice of
// the composed permutation, dst the training step's static batch buffer.  One warp per row; 16-byte vectors when
// pointers, pitches and the row length allow, bytes otherwise.  Rows whose index is outside [0, n_src) are
// zero-filled and counted in *bad (the caller checks it once per epoch).
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint8_t* __restrict__ src, long long src_pitch, long long n_src,
                   const long long* __restrict__ idx, int n_rows, long long row_bytes,
                   uint8_t* __restrict__ dst, long long dst_pitch, int vec, int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
        const long long s = __ldg(idx + r);
        const bool ok = s >= 0 && s < n_src;
        if (!ok && lane == 0 && bad) atomicAdd(bad, 1);
        const uint8_t* sp = src + (ok ? s : 0) * src_pitch;
        uint8_t* dp = dst + (long long)r * dst_pitch;
        if (vec) {
            const long long n16 = row_bytes >> 4;
            for (long long j = lane; j < n16; j += 32) {
                const uint4 v = ok ? __ldg(reinterpret_cast<const uint4*>(sp) + j) : make_uint4(0, 0, 0, 0);
                reinterpret_cast<uint4*>(dp)[j] = v;
            }
        } else {
            for (long long j = lane; j < row_bytes; j += 32) dp[j] = ok ? __ldg(sp + j) : (uint8_t)0;
        }
    }
}

// Packed sign bits -> E2M1 nibbles (BNN_DT_F4 operand): one thread turns one 32-bit word into 16 bytes
// (bit j -> nibble j: 1 -> 0x2 = +1.0, 0 -> 0xA = -1.0; columns past C -> 0x0 = +0.0, they pad K).
// Reads 4 bytes, writes one 16-byte vector per thread, both coalesced.
__global__ void __launch_bounds__(256)
bits_to_f4_kernel(const uint32_t* __restrict__ bits, long long R, int C, long long ldw,
                  uint8_t* __restrict__ out, long long ld_bytes) {
    const int words = (C + 31) >> 5;
    const long long total = R * words;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / words;
        const int w = (int)(i - r * words);
        const uint32_t b = __ldg(bits + r * ldw + w);
        const int valid = min(32, C - 32 * w);
        const uint32_t live = valid >= 32 ? 0xFFFFFFFFu : ((1u << valid) - 1u);
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            // live nibble: 0x2 | (bit ? 0 : 0x8); nibble past C: 0
            o[q] = bits_to_e2m1x8(b >> (8 * q));
            if (valid < 32) o[q] &= spread_bits_to_nibbles(live >> (8 * q)) * 0xFu;
        }
        *reinterpret_cast<uint4*>(out + r * ld_bytes + 16 * w) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// max |x| via integer atomicMax on the float bit pattern (non-negative floats order like uints).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n,
                                                     uint32_t* __restrict__ out) {
    float m = 0.0f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n4 = (aligned16(x) ? n / 4 : 0);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (long long j = i; j < n4; j += stride) {
        const float4 v = __ldg(x4 + j);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (long long j = n4 * 4 + i; j < n; j += stride) m = fmaxf(m, fabsf(__ldg(x + j)));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && __float_as_uint(m) > *reinterpret_cast<volatile uint32_t*>(out))
        atomicMax(out, __float_as_uint(m));
}

// ---------------------------------------------------------------------------------------------
// fp32 -> fp16 hi/lo split, row-major and transposed.  64 x 128 tiles, 8 columns per thread: two
// 16-byte loads, 16-byte stores of each piece; the transposed pieces go through shared memory and
// leave as 16-byte stores of 8 consecutive rows.
// ---------------------------------------------------------------------------------------------
constexpr int SP_ROWS = 64, SP_COLS = 128;

__device__ __forceinline__ uint4 pack8(const __half (&h)[8]) {
    uint4 v;
    v.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
    v.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
    v.z = (uint32_t)__half_as_ushort(h[4]) | ((uint32_t)__half_as_ushort(h[5]) << 16);
    v.w = (uint32_t)__half_as_ushort(h[6]) | ((uint32_t)__half_as_ushort(h[7]) << 16);
    return v;
}

__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ x, int R, int C, long long ldx,
                 const float* __restrict__ bound, __half* __restrict__ hi, __half* __restrict__ lo,
                 long long ld_rm, __half* __restrict__ hi_t, __half* __restrict__ lo_t,
                 long long ld_t, float* __restrict__ scale_out) {
    // 16-byte chunk index XOR-ed with (row / 8): conflict-free transposed reads (see bn.cu)
    __shared__ __align__(16) __half sh[SP_ROWS * SP_COLS];
    __shared__ __align__(16) __half sl[SP_ROWS * SP_COLS];
    auto pidx = [](int row, int col) {
        return row * SP_COLS + ((((col >> 3) ^ (row >> 3)) & 15) << 3) + (col & 7);
    };
    const float s = split_scale_from_bound(__ldg(bound));
    if (scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *scale_out = s;
    const int r0 = blockIdx.y * SP_ROWS, c0 = blockIdx.x * SP_COLS + (threadIdx.x & 15) * 8;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int valid = max(0, min(8, C - c0));
    const bool xvec = aligned16(x) && (ldx % 4 == 0);
    const bool rvec = hi && aligned16(hi) && aligned16(lo) && (ld_rm % 8 == 0);
#pragma unroll
    for (int i = 0; i < SP_ROWS / 16; ++i) {
        const int r = ty + 16 * i;
        const int gr = r0 + r;
        __half h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = l[j] = __ushort_as_half(0);
        if (gr < R && valid > 0) {
            const float* p = x + (long long)gr * ldx + c0;
            float v[8];
            if (xvec && valid == 8) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(p));
                const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
                v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = j < valid ? __ldg(p + j) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) split_f16(v[j] * s, h[j], l[j]);
            if (hi) {
                __half* ph = hi + (long long)gr * ld_rm + c0;
                __half* pl = lo + (long long)gr * ld_rm + c0;
                if (rvec && valid == 8) {
                    *reinterpret_cast<uint4*>(ph) = pack8(h);
                    *reinterpret_cast<uint4*>(pl) = pack8(l);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < valid) { ph[j] = h[j]; pl[j] = l[j]; }
                }
            }
        }
        if (hi_t) {
            *reinterpret_cast<uint4*>(&sh[pidx(r, tx * 8)]) = pack8(h);
            *reinterpret_cast<uint4*>(&sl[pidx(r, tx * 8)]) = pack8(l);
        }
    }
    if (!hi_t) return;
    __syncthreads();
    const int n0 = blockIdx.x * SP_COLS;
    const bool tvec = aligned16(hi_t) && aligned16(lo_t) && (ld_t % 8 == 0);
#pragma unroll
    for (int k = 0; k < (SP_COLS * SP_ROWS / 8) / 256; ++k) {
        // 8 adjacent lanes write the 8 x 16 bytes of one transposed row: full 128-byte lines
        const int item = threadIdx.x + 256 * k;
        const int rg = item & 7, n = item >> 3;
        const int gc = n0 + n, gr = r0 + rg * 8;
        if (gc >= C || gr >= R) continue;
        __half h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            h[j] = sh[pidx(rg * 8 + j, n)];
            l[j] = sl[pidx(rg * 8 + j, n)];
        }
        __half* ph = hi_t + (long long)gc * ld_t + gr;
        __half* pl = lo_t + (long long)gc * ld_t + gr;
        if (tvec && gr + 8 <= R) {
            *reinterpret_cast<uint4*>(ph) = pack8(h);
            *reinterpret_cast<uint4*>(pl) = pack8(l);
        } else {
            for (int j = 0; j < 8 && gr + j < R; ++j) { ph[j] = h[j]; pl[j] = l[j]; }
        }
    }
}

}  // namespace
}  // namespace bnn

using namespace bnn;

extern "C" int bnn_pack_sign_rows(const float* x, int R, int C, int64_t ldx, uint32_t* bits,
                                  int64_t ldw, bnn_stream_t stream) {
    BNN_CHECK_ARG(R >= 0 && C >= 0, "bnn_pack_sign_rows: negative shape");
    if (R == 0 || C == 0) return BNN_OK;
    BNN_CHECK_ARG(x && bits, "bnn_pack_sign_rows: null pointer");
    BNN_CHECK_ARG(ldx >= C && ldw >= (C + 31) / 32, "bnn_pack_sign_rows: pitch too small");
    const long long warps = (long long)R * ((C + 31) / 32);
    const int grid = (int)((warps * 32 + 255) / 256 < 148 * 16 ? (warps * 32 + 255) / 256 : 148 * 16);
    pack_rows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x, R, C, ldx, bits, ldw);
    BNN_LAUNCH_CHECK("pack_rows_kernel<float>");
    return BNN_OK;
}

extern "C" int bnn_pack_i8_rows(const int8_t* x, int R, int C, int64_t ldx, uint32_t* bits,
                                int64_t ldw, bnn_stream_t stream) {
    BNN_CHECK_ARG(R >= 0 && C >= 0, "bnn_pack_i8_rows: negative shape");
    if (R == 0 || C == 0) return BNN_OK;
    BNN_CHECK_ARG(x && bits, "bnn_pack_i8_rows: null pointer");
    BNN_CHECK_ARG(ldx >= C && ldw >= (C + 31) / 32, "bnn_pack_i8_rows: pitch too small");
    const long long warps = (long long)R * ((C + 31) / 32);
    const int grid = (int)((warps * 32 + 255) / 256 < 148 * 16 ? (warps * 32 + 255) / 256 : 148 * 16);
    pack_rows_kernel<int8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(x, R, C, ldx, bits, ldw);
    BNN_LAUNCH_CHECK("pack_rows_kernel<int8_t>");
    return BNN_OK;
}

extern "C" int bnn_binarize_weights_t(const float* W, int K, int N, int8_t* wt_i8, int64_t ld_i8,
                                      void* wt_f16, int64_t ld_f16, uint32_t* wt_bits,
                                      int64_t ld_bits, void* wt_f4, int64_t ld_f4, uint32_t* absmax_bits,
                                      bnn_stream_t stream) {
    BNN_CHECK_ARG(W && K > 0 && N > 0, "bnn_binarize_weights_t: bad input");
    BNN_CHECK_ARG(!wt_i8 || ld_i8 >= K, "bnn_binarize_weights_t: ld_i8 < K");
    BNN_CHECK_ARG(!wt_f16 || ld_f16 >= K, "bnn_binarize_weights_t: ld_f16 < K");
    BNN_CHECK_ARG(!wt_bits || ld_bits >= (K + 31) / 32, "bnn_binarize_weights_t: ld_bits too small");
    BNN_CHECK_ARG(!wt_f4 || (aligned16(wt_f4) && ld_f4 % 32 == 0 && ld_f4 >= ((K + 31) / 32) * 32),
                  "bnn_binarize_weights_t: wt_f4 must be 16-byte aligned with a pitch of whole 32-element words");
    dim3 grid(cdiv(N, BT_N), cdiv(K, BT_K));
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_binarize_weights_t: K too large");
    if (absmax_bits)
        BNN_CUDA(cudaMemsetAsync(absmax_bits, 0, sizeof(uint32_t), (cudaStream_t)stream));
    binarize_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        W, K, N, wt_i8, ld_i8, static_cast<__half*>(wt_f16), ld_f16, wt_bits, ld_bits,
        static_cast<uint8_t*>(wt_f4), ld_f4 / 2, absmax_bits);
    BNN_LAUNCH_CHECK("binarize_t_kernel");
    return BNN_OK;
}

extern "C" int bnn_unpack_bits(const uint32_t* bits, int R, int C, int64_t ldw, int8_t* out_i8,
                               int64_t ld_i8, void* out_f16, int64_t ld_f16, bnn_stream_t stream) {
    BNN_CHECK_ARG(bits && R > 0 && C > 0, "bnn_unpack_bits: bad input");
    const long long total = (long long)R * C;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    unpack_bits_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        bits, R, C, ldw, out_i8, ld_i8, static_cast<__half*>(out_f16), ld_f16);
    BNN_LAUNCH_CHECK("unpack_bits_kernel");
    return BNN_OK;
}

extern "C" int bnn_gather_rows(const void* src, int64_t src_pitch_bytes, int64_t n_src_rows, const int64_t* idx,
                               int n_rows, int64_t row_bytes, void* dst, int64_t dst_pitch_bytes, int32_t* bad_count,
                               bnn_stream_t stream) {
    BNN_CHECK_ARG(src && idx && dst && n_rows > 0 && row_bytes > 0 && n_src_rows > 0, "bnn_gather_rows: bad input");
    BNN_CHECK_ARG(src_pitch_bytes >= row_bytes && dst_pitch_bytes >= row_bytes,
                  "bnn_gather_rows: pitch smaller than the row (%lld, %lld < %lld)", (long long)src_pitch_bytes,
                  (long long)dst_pitch_bytes, (long long)row_bytes);
    const int vec = aligned16(src) && aligned16(dst) && src_pitch_bytes % 16 == 0 && dst_pitch_bytes % 16 == 0 &&
                    row_bytes % 16 == 0;
    const int blocks = (n_rows + 7) / 8;   // 8 warps per block, one row per warp
    const int grid = blocks < 148 * 8 ? blocks : 148 * 8;
    gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        static_cast<const uint8_t*>(src), src_pitch_bytes, n_src_rows, reinterpret_cast<const long long*>(idx),
        n_rows, row_bytes, static_cast<uint8_t*>(dst), dst_pitch_bytes, vec, bad_count);
    BNN_LAUNCH_CHECK("gather_rows_kernel");
    return BNN_OK;
}

extern "C" int bnn_bits_to_f4(const uint32_t* bits, int64_t R, int C, int64_t ldw, void* out_f4,
                              int64_t ld_f4, bnn_stream_t stream) {
    BNN_CHECK_ARG(bits && out_f4 && R > 0 && C > 0 && ldw >= (C + 31) / 32, "bnn_bits_to_f4: bad input");
    BNN_CHECK_ARG(aligned16(out_f4) && ld_f4 % 32 == 0 && ld_f4 >= ((C + 31) / 32) * 32,
                  "bnn_bits_to_f4: out_f4 must be 16-byte aligned with a pitch of whole 32-element words (ld_f4=%lld)",
                  (long long)ld_f4);
    const long long total = R * ((C + 31) / 32);
    const long long blocks = (total + 255) / 256;
    const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    bits_to_f4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bits, R, C, ldw, static_cast<uint8_t*>(out_f4),
                                                              ld_f4 / 2);
    BNN_LAUNCH_CHECK("bits_to_f4_kernel");
    return BNN_OK;
}

extern "C" int bnn_absmax_f32(const float* x, int64_t n, uint32_t* absmax_bits, int zero_first,
                              bnn_stream_t stream) {
    BNN_CHECK_ARG(absmax_bits && n >= 0, "bnn_absmax_f32: bad input");
    cudaStream_t s = (cudaStream_t)stream;
    if (zero_first) BNN_CUDA(cudaMemsetAsync(absmax_bits, 0, sizeof(uint32_t), s));
    if (n == 0) return BNN_OK;
    BNN_CHECK_ARG(x, "bnn_absmax_f32: null x");
    const long long blocks = (n / 4 + 255) / 256;
    const int grid = (int)(blocks < 148 * 8 ? (blocks > 0 ? blocks : 1) : 148 * 8);
    absmax_kernel<<<grid, 256, 0, s>>>(x, n, absmax_bits);
    BNN_LAUNCH_CHECK("absmax_kernel");
    return BNN_OK;
}

extern "C" int bnn_split_f16(const float* x, int R, int C, int64_t ldx, const float* bound, void* hi,
                             void* lo, int64_t ld_rm, void* hi_t, void* lo_t, int64_t ld_t,
                             float* scale_out, bnn_stream_t stream) {
    BNN_CHECK_ARG(x && bound && R > 0 && C > 0, "bnn_split_f16: bad input");
    BNN_CHECK_ARG((hi == nullptr) == (lo == nullptr) && (hi_t == nullptr) == (lo_t == nullptr),
                  "bnn_split_f16: hi/lo must be given in pairs");
    BNN_CHECK_ARG(!hi || ld_rm >= C, "bnn_split_f16: ld_rm < C");
    BNN_CHECK_ARG(!hi_t || ld_t >= R, "bnn_split_f16: ld_t < R");
    dim3 grid(cdiv(C, SP_COLS), cdiv(R, SP_ROWS));
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_split_f16: R too large");
    split_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        x, R, C, ldx, bound, static_cast<__half*>(hi), static_cast<__half*>(lo), ld_rm,
        static_cast<__half*>(hi_t), static_cast<__half*>(lo_t), ld_t, scale_out);
    BNN_LAUNCH_CHECK("split_f16_kernel");
    return BNN_OK;
}
