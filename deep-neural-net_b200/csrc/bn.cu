// Batch-norm forward/backward, sign activation, softmax + cross-entropy, SGD and accuracy kernels.
// All HBM-bound: each reads its operands once and writes each result once.
//   A4/A5  BatchNormLayer.forward    mlpcode/layers.py:74-99
//   A6     unitstep / softmax        mlpcode/activation.py:106-113, 44-53
//   A7     hard_tanh_derivative      mlpcode/activation.py:122-131 (identity on +-1 inputs)
//   A8     BatchNormLayer.backwards  mlpcode/layers.py:101-127
//   A10    cross_entropy_loss/_derivative  mlpcode/loss.py:10-36, 39-68
#include "common.cuh"

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

constexpr int CS_COLS = 128;  // columns per block (32 lanes x 4)
constexpr int CS_ROWS = 256;  // rows per block

// ---------------------------------------------------------------------------------------------
// Column sums / sums of squares in fp64 (exact for integer-valued Z; order independent).
// block = (32, 8); thread owns 4 adjacent columns and every 8th row of a 256-row slab.
// ---------------------------------------------------------------------------------------------
template <int ZT>
__global__ void __launch_bounds__(256) colstats_kernel(const void* __restrict__ Z, int B, int N,
                                                       long long ldz, double* __restrict__ sums) {
    __shared__ double red[2][8][CS_COLS];
    const int lane = threadIdx.x, y = threadIdx.y;
    const int c0 = blockIdx.x * CS_COLS + lane * 4;
    const int r0 = blockIdx.y * CS_ROWS;
    const int r1 = min(r0 + CS_ROWS, B);
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    for (int r = r0 + y; r < r1; r += 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (c0 + j < N) {
                const double v = (double)ZLoad<ZT>::get(Z, (long long)r * ldz + c0 + j);
                s[j] += v;
                q[j] += v * v;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        red[0][y][lane * 4 + j] = s[j];
        red[1][y][lane * 4 + j] = q[j];
    }
    __syncthreads();
    const int t = y * 32 + lane;  // 256 threads -> 2 x 128 outputs
    const int which = t >> 7, col = t & 127;
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

// ---------------------------------------------------------------------------------------------
// BN apply + sign, 64 x 64 tiles.  The normalisation follows the reference's fp32 operation
// order (sub, div by sqrt(var+eps), mul gamma, add beta) with explicit round-to-nearest
// intrinsics so no FMA contraction can change a sign decision.
// ---------------------------------------------------------------------------------------------
template <int ZT>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const void* __restrict__ Z, int B, int N, long long ldz, const float* __restrict__ mu,
                const float* __restrict__ var, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, float* __restrict__ out_f32,
                long long ld_f32, int8_t* __restrict__ act_i8, long long ld_i8,
                __half* __restrict__ act_t16, long long ld_t16, uint32_t* __restrict__ act_bits,
                long long ld_bits) {
    __shared__ uint8_t sg[64][68];
    const int b0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tn = threadIdx.x & 63, tb = threadIdx.x >> 6;
    const int gn = n0 + tn;
    const bool col_ok = gn < N;
    float m = 0.f, d = 1.f, g = 0.f, be = 0.f;
    if (col_ok) {
        m = __ldg(mu + gn);
        d = __fsqrt_rn(__fadd_rn(__ldg(var + gn), eps));
        g = __ldg(gamma + gn);
        be = __ldg(beta + gn);
    }
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int r = tb + 4 * i;
        const int gb = b0 + r;
        bool plus = false;
        if (col_ok && gb < B) {
            const float z = ZLoad<ZT>::get(Z, (long long)gb * ldz + gn);
            const float o = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(z, m), d), g), be);
            plus = o >= 0.0f;
            if (out_f32) out_f32[(long long)gb * ld_f32 + gn] = o;
            if (act_i8) act_i8[(long long)gb * ld_i8 + gn] = plus ? (int8_t)1 : (int8_t)-1;
        }
        if (act_bits) {
            const uint32_t word = __ballot_sync(0xffffffffu, plus);
            if ((threadIdx.x & 31) == 0 && gb < B && (n0 + (tn & 32)) < N)
                act_bits[(long long)gb * ld_bits + ((n0 + (tn & 32)) >> 5)] = word;
        }
        sg[r][tn] = plus ? 1 : 0;
    }
    if (!act_t16) return;
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int c = tb + 4 * i;
        const int gb = b0 + tn, gc = n0 + c;
        if (gb < B && gc < N)
            act_t16[(long long)gc * ld_t16 + gb] =
                __ushort_as_half(sg[tn][c] ? (unsigned short)0x3C00 : (unsigned short)0xBC00);
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
// ---------------------------------------------------------------------------------------------
template <int ZT>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ delta, long long ldd, const void* __restrict__ Z,
                     long long ldz, int B, int N, const float* __restrict__ mu,
                     double* __restrict__ red, uint32_t* __restrict__ red_max) {
    __shared__ double sred[2][8][CS_COLS];
    __shared__ float smax[2][8][CS_COLS];
    const int lane = threadIdx.x, y = threadIdx.y;
    const int c0 = blockIdx.x * CS_COLS + lane * 4;
    const int r0 = blockIdx.y * CS_ROWS;
    const int r1 = min(r0 + CS_ROWS, B);
    float m[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = (c0 + j < N) ? __ldg(mu + c0 + j) : 0.f;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    float md[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0};
    for (int r = r0 + y; r < r1; r += 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (c0 + j < N) {
                const float dl = __ldg(delta + (long long)r * ldd + c0 + j);
                const float xm = ZLoad<ZT>::get(Z, (long long)r * ldz + c0 + j) - m[j];
                s1[j] += (double)dl;
                s2[j] += (double)dl * (double)xm;
                md[j] = fmaxf(md[j], fabsf(dl));
                mx[j] = fmaxf(mx[j], fabsf(xm));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sred[0][y][lane * 4 + j] = s1[j];
        sred[1][y][lane * 4 + j] = s2[j];
        smax[0][y][lane * 4 + j] = md[j];
        smax[1][y][lane * 4 + j] = mx[j];
    }
    __syncthreads();
    const int t = y * 32 + lane;
    const int which = t >> 7, col = t & 127;
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
                                       float* __restrict__ dgamma_out, float* __restrict__ dbeta_out) {
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
        coef[n] = (float)c0;
        coef[N + n] = (float)c1;
        coef[2 * N + n] = (float)c2;
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
    }
    bound = warp_max(bound);
    if ((threadIdx.x & 31) == 0 && bound > 0.f && isfinite(bound))
        atomicMax(bound_bits, __float_as_uint(bound));
}

// ---------------------------------------------------------------------------------------------
// BN backward, pass 2: delta' and its fp16 hi/lo split in both operand layouts (64 x 64 tiles).
// ---------------------------------------------------------------------------------------------
template <int ZT>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ delta, long long ldd, const void* __restrict__ Z,
                    long long ldz, int B, int N, const float* __restrict__ mu,
                    const float* __restrict__ coef, const float* __restrict__ bound,
                    float* __restrict__ out_f32, long long ld_f32, __half* __restrict__ d_hi,
                    __half* __restrict__ d_lo, long long ld_rm, __half* __restrict__ dt_hi,
                    __half* __restrict__ dt_lo, long long ld_t, float* __restrict__ scale_out) {
    __shared__ __half sh[64][66];
    __shared__ __half sl[64][66];
    const float s = split_scale_from_bound(__ldg(bound));
    if (scale_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *scale_out = s;
    const int b0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int tn = threadIdx.x & 63, tb = threadIdx.x >> 6;
    const int gn = n0 + tn;
    const bool col_ok = gn < N;
    float m = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    if (col_ok) {
        m = __ldg(mu + gn);
        c0 = __ldg(coef + gn);
        c1 = __ldg(coef + N + gn);
        c2 = __ldg(coef + 2 * N + gn);
    }
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int r = tb + 4 * i;
        const int gb = b0 + r;
        __half h = __ushort_as_half(0), l = __ushort_as_half(0);
        if (col_ok && gb < B) {
            const float dl = __ldg(delta + (long long)gb * ldd + gn);
            const float xm = ZLoad<ZT>::get(Z, (long long)gb * ldz + gn) - m;
            const float nd = fmaf(c0, dl, fmaf(c1, xm, c2));
            if (out_f32) out_f32[(long long)gb * ld_f32 + gn] = nd;
            split_f16(nd * s, h, l);
            if (d_hi) {
                d_hi[(long long)gb * ld_rm + gn] = h;
                d_lo[(long long)gb * ld_rm + gn] = l;
            }
        }
        sh[r][tn] = h;
        sl[r][tn] = l;
    }
    if (!dt_hi) return;
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const int c = tb + 4 * i;
        const int gb = b0 + tn, gc = n0 + c;
        if (gb < B && gc < N) {
            dt_hi[(long long)gc * ld_t + gb] = sh[tn][c];
            dt_lo[(long long)gc * ld_t + gb] = sl[tn][c];
        }
    }
}

__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  long long n, float lr) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = __fsub_rn(p[i], __fmul_rn(lr, __ldg(g + i)));
}

__global__ void __launch_bounds__(256) count_correct_kernel(const float* __restrict__ scores, int B,
                                                            int C, long long lds,
                                                            const int32_t* __restrict__ labels,
                                                            int32_t* __restrict__ correct) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    int hit = 0;
    if (b < B) {
        int best = 0;
        float bv = scores[(long long)b * lds];
        for (int c = 1; c < C; ++c) {
            const float v = scores[(long long)b * lds + c];
            if (v > bv) { bv = v; best = c; }
        }
        hit = (best == labels[b]) ? 1 : 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(correct, __popc(m));
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
    dim3 grid(cdiv(N, CS_COLS), cdiv(B, CS_ROWS)), block(32, 8);
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_colstats: B too large");
    return dispatch_z(z_dtype, [&](auto zt) {
        colstats_kernel<decltype(zt)::value><<<grid, block, 0, s>>>(Z, B, N, ldz, sums);
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
    dim3 grid(cdiv(N, 64), cdiv(B, 64));
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
    dim3 grid(cdiv(N, CS_COLS), cdiv(B, CS_ROWS)), block(32, 8);
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_bn_bwd_reduce: B too large");
    return dispatch_z(z_dtype, [&](auto zt) {
        bn_bwd_reduce_kernel<decltype(zt)::value><<<grid, block, 0, s>>>(delta, ldd, Z, ldz, B, N,
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
                                   bnn_stream_t stream) {
    BNN_CHECK_ARG(red && red_max && mu && var && gamma && beta && mu_run && sigma_run && coef &&
                      bound_bits && N > 0 && count > 0,
                  "bnn_bn_bwd_finalize: bad input");
    cudaStream_t s = (cudaStream_t)stream;
    BNN_CUDA(cudaMemsetAsync(bound_bits, 0, sizeof(uint32_t), s));
    // (MA_BETA * mu) + ((1 - MA_BETA) * batch_mu) with Python-float constants cast to fp32
    const float ma = (float)momentum, oma = (float)(1.0 - momentum);
    bn_bwd_finalize_kernel<<<cdiv(N, 128), 128, 0, s>>>(red, red_max, N, count, mu, var, eps, gamma,
                                                        beta, mu_run, sigma_run, lr, ma, oma, sums,
                                                        coef, bound_bits, dgamma_out, dbeta_out);
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
    dim3 grid(cdiv(N, 64), cdiv(B, 64));
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

extern "C" int bnn_sgd_update(float* p, const float* g, int64_t n, float lr, bnn_stream_t stream) {
    BNN_CHECK_ARG(p && g && n >= 0, "bnn_sgd_update: bad input");
    if (n == 0) return BNN_OK;
    const long long blocks = (n + 255) / 256;
    const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    sgd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, n, lr);
    BNN_LAUNCH_CHECK("sgd_kernel");
    return BNN_OK;
}

extern "C" int bnn_count_correct(const float* scores, int B, int C, int64_t lds,
                                 const int32_t* labels, int32_t* correct, bnn_stream_t stream) {
    BNN_CHECK_ARG(scores && labels && correct && B > 0 && C > 0 && lds >= C,
                  "bnn_count_correct: bad input");
    count_correct_kernel<<<cdiv(B, 256), 256, 0, (cudaStream_t)stream>>>(scores, B, C, lds, labels,
                                                                         correct);
    BNN_LAUNCH_CHECK("count_correct_kernel");
    return BNN_OK;
}
