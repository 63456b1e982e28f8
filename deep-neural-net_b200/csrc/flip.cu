// Fault-injection kernels: many trials per launch, masks from a counter-based RNG.
//   A11  ErrorCallback._flipBNNMask     mlpcode/callbacks.py:79-83  (Bernoulli(p) sign flips)
//   A12  ImageErrorCallback.intError    mlpcode/callbacks.py:130-147 (exact-k bit flips per row)
// plus the exact-k weight flips of BASELINE config 4 and utils.normalize (utils.py:178-187).
#include "common.cuh"
#include "fault_math.cuh"

#include <stdlib.h>
#include <string.h>
#include "rng.cuh"

namespace bnn {
namespace {

using rng::U4;

__device__ __forceinline__ uint32_t u4_get(const U4& v, uint32_t i) {
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// One thread per 32-element word (n, w) of one trial.  e = n*K + k indexes the RNG.
__global__ void __launch_bounds__(256)
flip_bernoulli_kernel(const int8_t* __restrict__ wt_i8, long long ld_i8,
                      const uint32_t* __restrict__ wt_bits, long long ld_bits, int N, int K,
                      unsigned long long threshold, uint32_t k0, uint32_t k1, uint32_t stream_id,
                      uint32_t trial0, int n_trials, int8_t* __restrict__ out_i8,
                      long long out_i8_stride, __half* __restrict__ out_f16, long long ld_f16,
                      long long out_f16_stride, uint32_t* __restrict__ out_bits,
                      long long out_bits_stride, uint32_t* __restrict__ mask_bits,
                      long long mask_stride) {
    const int words = (K + 31) >> 5;
    const long long per_trial = (long long)N * words;
    const long long total = per_trial * n_trials;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(idx / per_trial);
        const long long r = idx - (long long)t * per_trial;
        const int n = (int)(r / words), w = (int)(r % words);
        const int kbase = w * 32;
        const int cnt = min(32, K - kbase);
        // clean word
        uint32_t clean;
        if (wt_bits) {
            clean = __ldg(wt_bits + (long long)n * ld_bits + w);
        } else {
            clean = 0;
            for (int j = 0; j < cnt; ++j)
                clean |= (uint32_t)(__ldg(wt_i8 + (long long)n * ld_i8 + kbase + j) > 0) << j;
        }
        // mask word
        uint32_t mask = 0;
        const unsigned long long e0 = (unsigned long long)n * (unsigned long long)K + kbase;
        U4 rnd = {0, 0, 0, 0};
        if (threshold && cnt == 32 && (e0 & 3ull) == 0) {
            // whole word on a counter boundary (K % 4 == 0): 8 Philox calls, four 32-bit compares each — the same
            // draws as the element loop below, without its per-element branch and 64-bit compares
            const bool all = threshold >= (1ull << 32);
            const uint32_t thr32 = all ? 0xFFFFFFFFu : (uint32_t)threshold;
            const unsigned long long q0 = e0 >> 2;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const unsigned long long q = q0 + g;
                const U4 r = rng::philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), stream_id, trial0 + t, k0, k1);
                mask |= ((uint32_t)(r.x < thr32) | ((uint32_t)(r.y < thr32) << 1) | ((uint32_t)(r.z < thr32) << 2) |
                         ((uint32_t)(r.w < thr32) << 3)) << (4 * g);
            }
            if (all) mask = 0xFFFFFFFFu;
        } else
        for (int j = 0; j < (threshold ? cnt : 0); ++j) {  // threshold 0: plain replication
            const unsigned long long e = e0 + j;
            if (j == 0 || (e & 3ull) == 0) {
                const unsigned long long q = e >> 2;
                rnd = rng::philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), stream_id, trial0 + t, k0,
                                         k1);
            }
            if ((unsigned long long)u4_get(rnd, (uint32_t)(e & 3ull)) < threshold) mask |= 1u << j;
        }
        const uint32_t flipped = clean ^ mask;
        if (mask_bits) mask_bits[(long long)t * mask_stride + (long long)n * ld_bits + w] = mask;
        if (out_bits) out_bits[(long long)t * out_bits_stride + (long long)n * ld_bits + w] = flipped;
        if (out_i8) {
            int8_t* o = out_i8 + (long long)t * out_i8_stride + (long long)n * ld_i8 + kbase;
            if (cnt == 32 && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0)) {
                uint32_t pk[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    uint32_t v = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        v |= (((flipped >> (g * 4 + b)) & 1u) ? 0x01u : 0xFFu) << (8 * b);
                    pk[g] = v;
                }
                reinterpret_cast<uint4*>(o)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                reinterpret_cast<uint4*>(o)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
                for (int j = 0; j < cnt; ++j) o[j] = ((flipped >> j) & 1u) ? (int8_t)1 : (int8_t)-1;
            }
        }
        if (out_f16) {
            __half* o = out_f16 + (long long)t * out_f16_stride + (long long)n * ld_f16 + kbase;
            for (int j = 0; j < cnt; ++j)
                o[j] = __ushort_as_half(((flipped >> j) & 1u) ? (unsigned short)0x3C00
                                                               : (unsigned short)0xBC00);
        }
    }
}

// Exact-k: thread (i, trial) flips position P_trial(i) if it falls into this layer.
__global__ void __launch_bounds__(256)
flip_exactk_kernel(int N, int K, unsigned long long total, unsigned long long offset,
                   unsigned long long seed, uint32_t trial0, const int32_t* __restrict__ k_per_trial,
                   int bits, int8_t* __restrict__ io_i8, long long ld_i8, long long io_i8_stride,
                   __half* __restrict__ io_f16, long long ld_f16, long long io_f16_stride,
                   uint32_t* __restrict__ io_bits, long long ld_bits, long long io_bits_stride) {
    const int t = blockIdx.y;
    const int kt = k_per_trial[t];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kt) return;
    const rng::FeistelKey key = rng::feistel_key(seed, rng::kStreamExactK, trial0 + t, 0u);
    const unsigned long long pos = rng::feistel_perm((unsigned long long)i, total, bits, key);
    if (pos < offset || pos >= offset + (unsigned long long)N * K) return;
    const unsigned long long e = pos - offset;
    const int n = (int)(e / K), k = (int)(e % K);
    if (io_i8) {
        int8_t* p = io_i8 + (long long)t * io_i8_stride + (long long)n * ld_i8 + k;
        *p = (int8_t)(-*p);
    }
    if (io_f16) {
        unsigned short* p = reinterpret_cast<unsigned short*>(io_f16) + (long long)t * io_f16_stride +
                            (long long)n * ld_f16 + k;
        *p ^= 0x8000u;
    }
    if (io_bits)
        atomicXor(io_bits + (long long)t * io_bits_stride + (long long)n * ld_bits + (k >> 5),
                  1u << (k & 31));
}

// IEEE-754 bit flips of real-valued weights (ErrorCallback non-BNN branch, callbacks.py:35-59, 100-123).
// Thread i owns chosen element P(i): distinct elements, so the read-modify-write needs no atomics.
// The per-element arithmetic lives in fault_math.cuh (also compiled for the host by the tests).
__global__ void __launch_bounds__(256)
flip_f32_bits_kernel(uint32_t* __restrict__ w, unsigned long long n, unsigned long long n_sel, double p,
                     int mode, unsigned long long seed, uint32_t trial, int ebits) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sel) return;
    const unsigned long long e = fault::f32_flip_element(i, n, ebits, seed, trial);
    const uint32_t v = w[e];
    const uint32_t m = fault::f32_flip_mask(v, e, p, mode, seed, trial);
    if (m) w[e] = v ^ m;
}

__global__ void __launch_bounds__(256)
flip_mask_kernel(const uint8_t* __restrict__ mask_kn, int K, int N, int8_t* __restrict__ io_i8,
                 long long ld_i8, __half* __restrict__ io_f16, long long ld_f16,
                 uint32_t* __restrict__ io_bits, long long ld_bits) {
    const int words = (K + 31) >> 5;
    const long long total = (long long)N * words;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        // consecutive threads take consecutive n for the same word so the mask reads coalesce
        const int w = (int)(idx / N), n = (int)(idx % N);
        const int kbase = w * 32, cnt = min(32, K - kbase);
        uint32_t m = 0;
        for (int j = 0; j < cnt; ++j)
            m |= (uint32_t)(mask_kn[(long long)(kbase + j) * N + n] != 0) << j;
        if (!m) continue;
        if (io_bits) io_bits[(long long)n * ld_bits + w] ^= m;
        for (int j = 0; j < cnt; ++j) {
            if (!((m >> j) & 1u)) continue;
            if (io_i8) {
                int8_t* p = io_i8 + (long long)n * ld_i8 + kbase + j;
                *p = (int8_t)(-*p);
            }
            if (io_f16) {
                unsigned short* p =
                    reinterpret_cast<unsigned short*>(io_f16) + (long long)n * ld_f16 + kbase + j;
                *p ^= 0x8000u;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Image rows: block per (row, trial); the row lives in shared memory while its bits are flipped.
// Bit position j of a row is byte j/8, bit 7-(j%8) (np.unpackbits is MSB first).
// ---------------------------------------------------------------------------------------------
constexpr int IMG_MAX_F = 4096;

// `words` = rows start on 4-byte boundaries (F % 4 == 0, aligned bases): whole-word loads and stores.
// xor_out / minmax: the by-products of bnn_flip_images_ex (operand form and utils.normalize's
// min/max of the corrupted set), taken from the row while it is still in shared memory.
__global__ void __launch_bounds__(128)
flip_images_kernel(const uint8_t* __restrict__ in, int R, int F, int k, int mode,
                   unsigned long long seed, uint32_t trial0, uint8_t* __restrict__ out, int words,
                   uint32_t xor_out, uint32_t* __restrict__ minmax, int lockstep) {
    __shared__ uint32_t row[IMG_MAX_F / 4];
    __shared__ uint32_t orig[IMG_MAX_F / 4];
    __shared__ int pref[IMG_MAX_F / 4 + 1];  // exclusive prefix of candidate counts per 32-bit word
    __shared__ int s_ncand;
    __shared__ uint32_t s_lo[4], s_hi[4];
    const int r = blockIdx.x, t = blockIdx.y;
    const int nwords = (F + 3) >> 2;
    const uint8_t* src = in + (long long)r * F;
    for (int w = threadIdx.x; w < nwords; w += blockDim.x) {
        uint32_t v = 0;
        if (words) {
            v = __ldg(reinterpret_cast<const uint32_t*>(src) + w);
        } else {
            for (int b = 0; b < 4; ++b) {
                const int j = w * 4 + b;
                if (j < F) v |= (uint32_t)src[j] << (8 * b);
            }
        }
        row[w] = v;
        orig[w] = v;
    }
    __syncthreads();
    int ncand = 8 * F;
    if (mode != 2) {
        // candidate = bits equal to (mode == 1); padding bytes beyond F never count
        if (threadIdx.x == 0) {
            int acc = 0;
            for (int w = 0; w < nwords; ++w) {
                pref[w] = acc;
                const int valid_bytes = min(4, F - w * 4);
                const uint32_t vm = valid_bytes == 4 ? 0xFFFFFFFFu : ((1u << (8 * valid_bytes)) - 1u);
                const uint32_t cand = (mode == 1 ? orig[w] : ~orig[w]) & vm;
                acc += __popc(cand);
            }
            pref[nwords] = acc;
            s_ncand = acc;
        }
        __syncthreads();
        ncand = s_ncand;
    }
    const int kk = min(k, ncand);
    if (kk > 0) {
        const rng::FeistelKey key = rng::feistel_key(seed, rng::kStreamImage, trial0 + t, (uint32_t)r);
        const int bits = rng::feistel_bits((unsigned long long)ncand);
        // flip the candidate of the given rank (candidates ranked in MSB-first bit order)
        auto commit = [&](int rank) {
            int byte_idx, bit_in_byte;  // bit_in_byte counted MSB first (0 = 0x80)
            if (mode == 2) {
                byte_idx = rank >> 3;
                bit_in_byte = rank & 7;
            } else {
                // word containing the rank-th candidate (in MSB-first order within each byte)
                int lo = 0, hi = nwords - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (pref[mid] <= rank) lo = mid; else hi = mid - 1;
                }
                int rem = rank - pref[lo];
                const uint32_t cw = (mode == 1 ? orig[lo] : ~orig[lo]);
                byte_idx = -1;
                bit_in_byte = 0;
                for (int b = 0; b < 4 && byte_idx < 0; ++b) {
                    if (lo * 4 + b >= F) break;
                    const uint32_t by = (cw >> (8 * b)) & 0xFFu;
                    const int c = __popc(by);
                    if (rem < c) {
                        for (int q = 0; q < 8; ++q) {
                            if ((by >> (7 - q)) & 1u) {
                                if (rem == 0) { byte_idx = lo * 4 + b; bit_in_byte = q; break; }
                                --rem;
                            }
                        }
                    } else {
                        rem -= c;
                    }
                }
            }
            if (byte_idx >= 0)
                atomicXor(&row[byte_idx >> 2], (0x80u >> bit_in_byte) << (8 * (byte_idx & 3)));
        };
        if (lockstep) {
            // Thread `tid` owns draws tid, tid + 128, ...  Every lane applies the bijection ONCE per trip;
            // a lane whose value lands in range commits it and moves to its next draw.  A warp's trip
            // count is then the longest SUM of walks among its lanes (about 1.4x the mean) instead of
            // the sum over draws of the longest walk (about 3x): same positions, less divergence.
            int i = threadIdx.x;
            unsigned long long x = (unsigned long long)i;
            while (i < kk) {
                x = rng::feistel_once(x, bits, key);
                if (x < (unsigned long long)ncand) {
                    commit((int)x);
                    i += blockDim.x;
                    x = (unsigned long long)i;
                }
            }
        } else {
            for (int i = threadIdx.x; i < kk; i += blockDim.x)
                commit((int)rng::feistel_perm((unsigned long long)i, (unsigned long long)ncand, bits, key));
        }
    }
    __syncthreads();
    uint8_t* dst = out + ((long long)t * R + r) * F;
    const uint32_t x4 = xor_out * 0x01010101u;
    if (words) {
        for (int w = threadIdx.x; w < nwords; w += blockDim.x)
            reinterpret_cast<uint32_t*>(dst)[w] = row[w] ^ x4;
    } else {
        for (int j = threadIdx.x; j < F; j += blockDim.x)
            dst[j] = (uint8_t)((row[j >> 2] >> (8 * (j & 3))) ^ xor_out);
    }
    if (minmax) {
        uint32_t lo = 255u, hi = 0u;
        for (int j = threadIdx.x; j < F; j += blockDim.x) {
            const uint32_t v = (row[j >> 2] >> (8 * (j & 3))) & 0xFFu;
            lo = min(lo, v);
            hi = max(hi, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if ((threadIdx.x & 31) == 0) {
            s_lo[threadIdx.x >> 5] = lo;
            s_hi[threadIdx.x >> 5] = hi;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
                lo = min(lo, s_lo[w]);
                hi = max(hi, s_hi[w]);
            }
            // R blocks share one pair of words per trial: only a block that improves on the value
            // it can already see issues the atomic
            uint32_t* mm = minmax + 2 * t;
            if (lo < *reinterpret_cast<volatile uint32_t*>(mm)) atomicMin(mm, lo);
            if (hi > *reinterpret_cast<volatile uint32_t*>(mm + 1)) atomicMax(mm + 1, hi);
        }
    }
}

__global__ void minmax_init_kernel(uint32_t* __restrict__ minmax, int n_trials) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_trials) {
        minmax[2 * t] = 255u;
        minmax[2 * t + 1] = 0u;
    }
}

// One warp per output unit n: S_n = sum_k Wb[k, n] from the packed bits, then the per-trial thresholds
// (see bnn_u8_input_thresholds in the header for the algebra).
__global__ void __launch_bounds__(256)
u8_input_thresholds_kernel(const uint32_t* __restrict__ wt_bits, long long ld_bits,
                           long long bits_stride, int N, int K, const float* __restrict__ thr,
                           const float* __restrict__ sgn, const uint32_t* __restrict__ minmax,
                           int minmax_stride, int n_trials, float new_min, float new_max, int bias,
                           float* __restrict__ thr_out, long long ld_thr) {
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const int nw = (K + 31) >> 5;
    const float tz = __ldg(thr + n);
    const double s = (double)__ldg(sgn + n);
    if (bits_stride == 0) {
        // one weight matrix, one min / max per trial (image faults): S_n once, lanes take trials
        int ones = 0;
        for (int w = lane; w < nw; w += 32) ones += __popc(__ldg(wt_bits + (long long)n * ld_bits + w));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ones += __shfl_xor_sync(0xffffffffu, ones, o);
        const double S = (double)(2 * ones - K);
        for (int t = lane; t < n_trials; t += 32)
            thr_out[(long long)t * ld_thr + n] = fault::u8_input_threshold(
                tz, s, S, (double)minmax[(long long)t * minmax_stride],
                (double)minmax[(long long)t * minmax_stride + 1], (double)new_min, (double)new_max,
                (double)bias);
    } else {
        // one weight matrix per trial (weight faults): S_n of every trial's own (flipped) column
        for (int t = 0; t < n_trials; ++t) {
            const uint32_t* row = wt_bits + (long long)t * bits_stride + (long long)n * ld_bits;
            int ones = 0;
            for (int w = lane; w < nw; w += 32) ones += __popc(__ldg(row + w));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ones += __shfl_xor_sync(0xffffffffu, ones, o);
            if (lane == 0)
                thr_out[(long long)t * ld_thr + n] = fault::u8_input_threshold(
                    tz, s, (double)(2 * ones - K), (double)minmax[(long long)t * minmax_stride],
                    (double)minmax[(long long)t * minmax_stride + 1], (double)new_min, (double)new_max,
                    (double)bias);
        }
    }
}

__global__ void __launch_bounds__(256) normalize_u8_kernel(const uint8_t* __restrict__ in, long long n,
                                                           const float* __restrict__ mn,
                                                           const float* __restrict__ mx, float new_min,
                                                           float new_max, float* __restrict__ out) {
    const float lo = __ldg(mn), hi = __ldg(mx);
    const double span = (double)hi - (double)lo;
    const float nrange = __fsub_rn(new_max, new_min);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // utils.py:178-187: ((x - min) * (newMax - newMin)) / (max - min) + newMin; the division is
        // by a float64 scalar (NumPy 2 promotes it), the result is rounded back to fp32
        const float a = __fmul_rn(__fsub_rn((float)in[i], lo), nrange);
        out[i] = __fadd_rn((float)((double)a / span), new_min);
    }
}

__global__ void __launch_bounds__(256) minmax_u8_kernel(const uint8_t* __restrict__ in, long long n,
                                                        uint32_t* __restrict__ minmax) {
    uint32_t lo = 255u, hi = 0u;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t v = in[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[0], lo);
        atomicMax(&minmax[1], hi);
    }
}

__global__ void u8_to_f32_minmax_kernel(const uint32_t* __restrict__ mm, float* __restrict__ out) {
    out[0] = (float)mm[0];
    out[1] = (float)mm[1];
}

}  // namespace
}  // namespace bnn

using namespace bnn;

extern "C" int bnn_flip_weights_bernoulli(const int8_t* wt_i8, int64_t ld_i8, const uint32_t* wt_bits,
                                          int64_t ld_bits, int N, int K, uint64_t threshold,
                                          uint64_t seed, uint32_t stream_id, uint32_t trial0,
                                          int n_trials, int8_t* out_i8, int64_t out_i8_stride,
                                          void* out_f16, int64_t ld_f16, int64_t out_f16_stride,
                                          uint32_t* out_bits, int64_t out_bits_stride,
                                          uint32_t* mask_bits, int64_t mask_stride,
                                          bnn_stream_t stream) {
    BNN_CHECK_ARG(N > 0 && K > 0 && n_trials > 0, "bnn_flip_weights_bernoulli: bad shape");
    BNN_CHECK_ARG(wt_i8 || wt_bits, "bnn_flip_weights_bernoulli: need wt_i8 or wt_bits");
    BNN_CHECK_ARG(threshold <= (1ull << 32), "bnn_flip_weights_bernoulli: threshold > 2^32");
    BNN_CHECK_ARG(!(out_bits || mask_bits || wt_bits) || ld_bits >= (K + 31) / 32,
                  "bnn_flip_weights_bernoulli: ld_bits too small");
    BNN_CHECK_ARG(!(out_i8 || wt_i8) || ld_i8 >= K, "bnn_flip_weights_bernoulli: ld_i8 < K");
    BNN_CHECK_ARG(!out_f16 || ld_f16 >= K, "bnn_flip_weights_bernoulli: ld_f16 < K");
    const long long total = (long long)N * ((K + 31) / 32) * n_trials;
    const long long blocks = (total + 255) / 256;
    const int grid = (int)(blocks < 148 * 32 ? blocks : 148 * 32);
    flip_bernoulli_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        wt_i8, ld_i8, wt_bits, ld_bits, N, K, threshold, (uint32_t)seed, (uint32_t)(seed >> 32),
        stream_id, trial0, n_trials, out_i8, out_i8_stride, static_cast<__half*>(out_f16), ld_f16,
        out_f16_stride, out_bits, out_bits_stride, mask_bits, mask_stride);
    BNN_LAUNCH_CHECK("flip_bernoulli_kernel");
    return BNN_OK;
}

extern "C" int bnn_flip_weights_exactk(int N, int K, uint64_t total, uint64_t offset, uint64_t seed,
                                       uint32_t trial0, int n_trials, const int32_t* k_per_trial,
                                       int k_max, int8_t* io_i8, int64_t ld_i8, int64_t io_i8_stride,
                                       void* io_f16, int64_t ld_f16, int64_t io_f16_stride,
                                       uint32_t* io_bits, int64_t ld_bits, int64_t io_bits_stride,
                                       bnn_stream_t stream) {
    BNN_CHECK_ARG(N > 0 && K > 0 && n_trials > 0 && k_per_trial, "bnn_flip_weights_exactk: bad input");
    BNN_CHECK_ARG(offset + (uint64_t)N * K <= total, "bnn_flip_weights_exactk: layer outside [0,total)");
    BNN_CHECK_ARG((uint64_t)k_max <= total, "bnn_flip_weights_exactk: k_max > total");
    BNN_CHECK_ARG(n_trials <= 65535, "bnn_flip_weights_exactk: at most 65535 trials per launch");
    if (k_max <= 0) return BNN_OK;
    dim3 grid(cdiv(k_max, 256), n_trials);
    flip_exactk_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        N, K, total, offset, seed, trial0, k_per_trial, rng::feistel_bits(total), io_i8, ld_i8,
        io_i8_stride, static_cast<__half*>(io_f16), ld_f16, io_f16_stride, io_bits, ld_bits,
        io_bits_stride);
    BNN_LAUNCH_CHECK("flip_exactk_kernel");
    return BNN_OK;
}

extern "C" int bnn_flip_weights_mask(const uint8_t* mask_kn, int K, int N, int8_t* io_i8,
                                     int64_t ld_i8, void* io_f16, int64_t ld_f16, uint32_t* io_bits,
                                     int64_t ld_bits, bnn_stream_t stream) {
    BNN_CHECK_ARG(mask_kn && K > 0 && N > 0, "bnn_flip_weights_mask: bad input");
    const long long total = (long long)N * ((K + 31) / 32);
    const long long blocks = (total + 255) / 256;
    const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    flip_mask_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        mask_kn, K, N, io_i8, ld_i8, static_cast<__half*>(io_f16), ld_f16, io_bits, ld_bits);
    BNN_LAUNCH_CHECK("flip_mask_kernel");
    return BNN_OK;
}

extern "C" int bnn_flip_images_ex(const uint8_t* in, int R, int F, int k, int mode, uint64_t seed,
                                  uint32_t trial0, int n_trials, uint8_t* out, int xor_out,
                                  uint32_t* minmax, bnn_stream_t stream) {
    BNN_CHECK_ARG(in && out && R > 0 && F > 0 && n_trials > 0, "bnn_flip_images: bad input");
    BNN_CHECK_ARG(F <= IMG_MAX_F, "bnn_flip_images: F > %d", IMG_MAX_F);
    BNN_CHECK_ARG(mode >= 0 && mode <= 2, "bnn_flip_images: mode must be 0, 1 or 2");
    BNN_CHECK_ARG(k >= 0, "bnn_flip_images: negative k");
    BNN_CHECK_ARG(n_trials <= 65535, "bnn_flip_images: at most 65535 trials per launch");
    BNN_CHECK_ARG(xor_out >= 0 && xor_out <= 255, "bnn_flip_images_ex: xor_out must be a byte value");
    cudaStream_t s = (cudaStream_t)stream;
    if (minmax) {
        minmax_init_kernel<<<cdiv(n_trials, 128), 128, 0, s>>>(minmax, n_trials);
        BNN_LAUNCH_CHECK("minmax_init_kernel");
    }
    const int words = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 3u) == 0) &&
                      ((reinterpret_cast<uintptr_t>(out) & 3u) == 0);
    dim3 grid(R, n_trials);
    // BNN_FLIP_WALK=perdraw keeps round 1's loop (one complete walk per draw); same output
    const char* walk = getenv("BNN_FLIP_WALK");  // read per call: tests switch it within one process
    const int lockstep = (walk && strcmp(walk, "perdraw") == 0) ? 0 : 1;
    flip_images_kernel<<<grid, 128, 0, s>>>(in, R, F, k, mode, seed, trial0, out, words,
                                            (uint32_t)xor_out, minmax, lockstep);
    BNN_LAUNCH_CHECK("flip_images_kernel");
    return BNN_OK;
}

extern "C" int bnn_flip_images(const uint8_t* in, int R, int F, int k, int mode, uint64_t seed,
                               uint32_t trial0, int n_trials, uint8_t* out, bnn_stream_t stream) {
    return bnn_flip_images_ex(in, R, F, k, mode, seed, trial0, n_trials, out, 0, nullptr, stream);
}

extern "C" int bnn_u8_input_thresholds(const uint32_t* wt_bits, int64_t ld_bits, int64_t bits_stride,
                                       int N, int K, const float* thr, const float* sgn,
                                       const uint32_t* minmax, int minmax_stride, int n_trials,
                                       float new_min, float new_max, int bias, float* thr_out,
                                       int64_t ld_thr, bnn_stream_t stream) {
    BNN_CHECK_ARG(wt_bits && thr && sgn && minmax && thr_out && N > 0 && K > 0 && n_trials > 0,
                  "bnn_u8_input_thresholds: bad input");
    BNN_CHECK_ARG(ld_bits >= (K + 31) / 32 && ld_thr >= N, "bnn_u8_input_thresholds: pitch too small");
    BNN_CHECK_ARG(new_max > new_min, "bnn_u8_input_thresholds: new_max must exceed new_min");
    BNN_CHECK_ARG(bias == 0 || bias == 128, "bnn_u8_input_thresholds: bias is 0 (uint8 operand) or 128");
    BNN_CHECK_ARG(bits_stride == 0 || bits_stride >= (int64_t)N * ld_bits,
                  "bnn_u8_input_thresholds: bits_stride is 0 (shared weights) or >= N*ld_bits");
    BNN_CHECK_ARG(minmax_stride == 0 || minmax_stride >= 2,
                  "bnn_u8_input_thresholds: minmax_stride is 0 (shared images) or >= 2");
    u8_input_thresholds_kernel<<<cdiv(N, 8), 256, 0, (cudaStream_t)stream>>>(
        wt_bits, ld_bits, bits_stride, N, K, thr, sgn, minmax, minmax_stride, n_trials, new_min, new_max,
        bias, thr_out, ld_thr);
    BNN_LAUNCH_CHECK("u8_input_thresholds_kernel");
    return BNN_OK;
}

extern "C" int bnn_normalize_u8(const uint8_t* in, int64_t n, const float* mn, const float* mx,
                                float new_min, float new_max, float* out, bnn_stream_t stream) {
    BNN_CHECK_ARG(in && mn && mx && out && n >= 0, "bnn_normalize_u8: bad input");
    if (n == 0) return BNN_OK;
    const long long blocks = (n + 255) / 256;
    const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    normalize_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, n, mn, mx, new_min, new_max, out);
    BNN_LAUNCH_CHECK("normalize_u8_kernel");
    return BNN_OK;
}

extern "C" int bnn_minmax_u8(const uint8_t* in, int64_t n, uint32_t* minmax, float* minmax_f32,
                             bnn_stream_t stream) {
    BNN_CHECK_ARG(in && minmax && n > 0, "bnn_minmax_u8: bad input");
    cudaStream_t s = (cudaStream_t)stream;
    static const uint32_t init[2] = {255u, 0u};
    BNN_CUDA(cudaMemcpyAsync(minmax, init, sizeof(init), cudaMemcpyHostToDevice, s));
    const long long blocks = (n + 255) / 256;
    const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    minmax_u8_kernel<<<grid, 256, 0, s>>>(in, n, minmax);
    BNN_LAUNCH_CHECK("minmax_u8_kernel");
    if (minmax_f32) {
        u8_to_f32_minmax_kernel<<<1, 1, 0, s>>>(minmax, minmax_f32);
        BNN_LAUNCH_CHECK("u8_to_f32_minmax_kernel");
    }
    return BNN_OK;
}

extern "C" int bnn_flip_f32_bits(float* w, int64_t n, int64_t n_sel, double p, int mode, uint64_t seed,
                                 uint32_t trial, bnn_stream_t stream) {
    BNN_CHECK_ARG(w && n > 0 && n < (1ll << 32), "bnn_flip_f32_bits: need 0 < n < 2^32 elements");
    BNN_CHECK_ARG(n_sel >= 0 && n_sel <= n, "bnn_flip_f32_bits: n_sel outside [0, n]");
    BNN_CHECK_ARG(p >= 0.0 && p <= 1.0, "bnn_flip_f32_bits: p outside [0, 1]");
    BNN_CHECK_ARG(mode >= 0 && mode <= 2, "bnn_flip_f32_bits: mode must be 0, 1 or 2");
    if (n_sel == 0) return BNN_OK;
    const long long blocks = (n_sel + 255) / 256;
    flip_f32_bits_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<uint32_t*>(w), (unsigned long long)n, (unsigned long long)n_sel, p, mode, seed,
        trial, rng::feistel_bits((unsigned long long)n));
    BNN_LAUNCH_CHECK("flip_f32_bits_kernel");
    return BNN_OK;
}
