// CUDA-core GEMM with the same contract as gemm_tc.cu (bnn_gemm, impl = BNN_GEMM_SIMT).
// It exists to cross-check the tcgen05 kernel on the device at sizes the CPU oracle cannot reach
// and to serve tiny shapes; it is shared-memory tiled (64x64x32, 4x4 outputs per thread) but makes
// no attempt at peak throughput.  int8 operands accumulate in int32 (exact), fp16 pieces in fp32.
#include "common.cuh"

namespace bnn {
namespace {

constexpr int TM = 64, TN = 64, TK = 32;

template <bool kI8>
struct Elem;
template <>
struct Elem<true> {
    using in_t = int8_t;
    using acc_t = int;
    __device__ static acc_t cvt(in_t v) { return (int)v; }
};
template <>
struct Elem<false> {
    using in_t = __half;
    using acc_t = float;
    __device__ static acc_t cvt(in_t v) { return __half2float(v); }
};

struct SimtParams {
    int M, N, K, n_pass;
    int pa[BNN_GEMM_MAX_PASS], pb[BNN_GEMM_MAX_PASS];
    const void* A[2];
    const void* B[2];
    long long lda, ldb, stride_a, stride_b;
    void* D;
    long long ldd, stride_d;
    const float* sa;
    const float* sb;
    float host_scale;
    int epilogue;
    const float* thr;
    const float* sgn;
    int thr_stride;
    int a_unsigned;
};

template <bool kI8>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams p) {
    using E = Elem<kI8>;
    using in_t = typename E::in_t;
    using acc_t = typename E::acc_t;
    __shared__ acc_t As[TK][TM + 1];
    __shared__ acc_t Bs[TK][TN + 1];
    const int b = blockIdx.z;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    acc_t acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0;

    for (int pass = 0; pass < p.n_pass; ++pass) {
        const in_t* A = reinterpret_cast<const in_t*>(p.A[p.pa[pass]]) + (long long)b * p.stride_a;
        const in_t* B = reinterpret_cast<const in_t*>(p.B[p.pb[pass]]) + (long long)b * p.stride_b;
        for (int k0 = 0; k0 < p.K; k0 += TK) {
            for (int i = threadIdx.x; i < TM * TK; i += 256) {
                const int r = i / TK, c = i % TK;
                const int gm = m0 + r, gk = k0 + c;
                acc_t av = (acc_t)0;
                if (gm < p.M && gk < p.K) {
                    av = E::cvt(A[(long long)gm * p.lda + gk]);
                    if (kI8 && p.a_unsigned) av = (acc_t)((int)av & 0xFF);  // the byte read as uint8
                }
                As[c][r] = av;
                const int gn = n0 + r;
                Bs[c][r] = (gn < p.N && gk < p.K) ? E::cvt(B[(long long)gn * p.ldb + gk]) : (acc_t)0;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < TK; ++k) {
                acc_t a[4], bb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * bb[j];
            }
            __syncthreads();
        }
    }
    float inv = 1.0f;
    if (p.sa) inv /= *p.sa;
    if (p.sb) inv /= *p.sb;
    const float scale = p.host_scale * inv;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= p.N) continue;
            const long long off = (long long)b * p.stride_d + (long long)gm * p.ldd + gn;
            switch (p.epilogue) {
                case BNN_EPI_STORE_S16:
                    reinterpret_cast<int16_t*>(p.D)[off] = (int16_t)(int)acc[i][j];
                    break;
                case BNN_EPI_STORE_S32:
                    reinterpret_cast<int32_t*>(p.D)[off] = (int)acc[i][j];
                    break;
                case BNN_EPI_SIGN_I8: {
                    const float a = kI8 ? (float)acc[i][j] : __fmul_rn((float)acc[i][j], scale);
                    reinterpret_cast<int8_t*>(p.D)[off] =
                        (__fmul_rn(a, p.sgn[gn]) >= p.thr[(long long)b * p.thr_stride + gn]) ? (int8_t)1
                                                                                               : (int8_t)-1;
                } break;
                case BNN_EPI_SGD_F32: {
                    float* d = reinterpret_cast<float*>(p.D) + off;
                    *d = __fsub_rn(*d, __fmul_rn((float)acc[i][j], scale));
                } break;
                default:
                    reinterpret_cast<float*>(p.D)[off] = __fmul_rn((float)acc[i][j], scale);
            }
        }
    }
}

}  // namespace

int gemm_simt_dispatch(const bnn_gemm_desc& g, cudaStream_t stream) {
    SimtParams p;
    p.M = g.M; p.N = g.N; p.K = g.K; p.n_pass = g.n_pass;
    for (int i = 0; i < BNN_GEMM_MAX_PASS; ++i) { p.pa[i] = g.pa[i]; p.pb[i] = g.pb[i]; }
    p.A[0] = g.A[0]; p.A[1] = g.A[1] ? g.A[1] : g.A[0];
    p.B[0] = g.B[0]; p.B[1] = g.B[1] ? g.B[1] : g.B[0];
    p.lda = g.lda; p.ldb = g.ldb; p.stride_a = g.stride_a; p.stride_b = g.stride_b;
    p.D = g.D; p.ldd = g.ldd; p.stride_d = g.stride_d;
    p.sa = g.sa; p.sb = g.sb; p.host_scale = g.host_scale; p.epilogue = g.epilogue;
    p.thr = g.epi_thr; p.sgn = g.epi_sgn;
    p.thr_stride = g.epi_thr_stride; p.a_unsigned = g.a_unsigned;
    dim3 grid(cdiv(g.N, TN), cdiv(g.M, TM), g.batch);
    BNN_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535, "bnn_gemm(simt): grid too large");
    if (g.in_dtype == BNN_DT_I8)
        gemm_simt_kernel<true><<<grid, 256, 0, stream>>>(p);
    else
        gemm_simt_kernel<false><<<grid, 256, 0, stream>>>(p);
    BNN_LAUNCH_CHECK("gemm_simt_kernel");
    return BNN_OK;
}

}  // namespace bnn
