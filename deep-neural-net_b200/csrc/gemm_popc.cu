// Variant A of the +-1 x +-1 GEMM: XOR + popc on the CUDA cores over packed sign bits.
//   D[m,n] = K - 2 * popc(a_bits[m,:] ^ b_bits[n,:])        (== X.dot(sign(W)), layers.py:220,
//   for X, W in {+-1}; bit 1 <-> +1, callbacks.py:68-72)
// 128x128 output tile per CTA, 8x8 per thread, 16 words (512 bits of K) per shared-memory stage,
// register-prefetched double buffering.  Words are staged transposed ([word][row]) so the inner
// loop reads them with conflict-free 128-bit LDS.  Pad bits beyond K are zero in both operands, so
// they never contribute to the XOR.
#include "common.cuh"

namespace bnn {
namespace {

constexpr int PT = 128;      // tile edge (rows of A and rows of B)
constexpr int PW = 16;       // words per stage
constexpr int PLD = PT + 4;  // padded row length of the transposed stage

__device__ __forceinline__ void load_stage(const uint32_t* __restrict__ g, long long ld, int rows,
                                           int row0, int words, int w0, uint32_t (&r)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int i = threadIdx.x + 256 * j;
        const int row = row0 + (i >> 4), w = w0 + (i & 15);
        r[j] = (row < rows && w < words) ? __ldg(g + (long long)row * ld + w) : 0u;
    }
}
__device__ __forceinline__ void store_stage(uint32_t (*s)[PLD], const uint32_t (&r)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int i = threadIdx.x + 256 * j;
        s[i & 15][i >> 4] = r[j];
    }
}

template <typename out_t>
__global__ void __launch_bounds__(256) gemm_popc_kernel(const uint32_t* __restrict__ A, long long lda,
                                                        const uint32_t* __restrict__ B, long long ldb,
                                                        int M, int N, int K, out_t* __restrict__ D,
                                                        long long ldd) {
    __shared__ __align__(16) uint32_t As[2][PW][PLD];
    __shared__ __align__(16) uint32_t Bs[2][PW][PLD];
    const int words = (K + 31) >> 5;
    const int m0 = blockIdx.y * PT, n0 = blockIdx.x * PT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    int acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0;

    uint32_t ra[8], rb[8];
    load_stage(A, lda, M, m0, words, 0, ra);
    load_stage(B, ldb, N, n0, words, 0, rb);
    store_stage(As[0], ra);
    store_stage(Bs[0], rb);
    __syncthreads();

    const int stages = (words + PW - 1) / PW;
    for (int st = 0; st < stages; ++st) {
        const int cur = st & 1;
        if (st + 1 < stages) {
            load_stage(A, lda, M, m0, words, (st + 1) * PW, ra);
            load_stage(B, ldb, N, n0, words, (st + 1) * PW, rb);
        }
#pragma unroll 4
        for (int k = 0; k < PW; ++k) {
            // rows ty*4..+3 and 64+ty*4..+3 ; cols tx*4..+3 and 64+tx*4..+3
            const uint4 a0 = *reinterpret_cast<const uint4*>(&As[cur][k][ty * 4]);
            const uint4 a1 = *reinterpret_cast<const uint4*>(&As[cur][k][64 + ty * 4]);
            const uint4 b0 = *reinterpret_cast<const uint4*>(&Bs[cur][k][tx * 4]);
            const uint4 b1 = *reinterpret_cast<const uint4*>(&Bs[cur][k][64 + tx * 4]);
            const uint32_t a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const uint32_t b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] += __popc(a[i] ^ b[j]);
        }
        if (st + 1 < stages) {
            store_stage(As[cur ^ 1], ra);
            store_stage(Bs[cur ^ 1], rb);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n < N) D[(long long)m * ldd + n] = (out_t)(K - 2 * acc[i][j]);
        }
    }
}

}  // namespace
}  // namespace bnn

extern "C" int bnn_gemm_popc(const uint32_t* a_bits, int64_t lda_w, const uint32_t* b_bits,
                             int64_t ldb_w, int M, int N, int K, void* D, int64_t ldd,
                             int out_dtype, bnn_stream_t stream) {
    using namespace bnn;
    BNN_CHECK_ARG(M > 0 && N > 0 && K > 0, "bnn_gemm_popc: empty shape M=%d N=%d K=%d", M, N, K);
    const int words = (K + 31) / 32;
    BNN_CHECK_ARG(lda_w >= words && ldb_w >= words, "bnn_gemm_popc: word pitch < ceil(K/32)");
    BNN_CHECK_ARG(out_dtype == BNN_EPI_STORE_S32 || out_dtype == BNN_EPI_STORE_S16,
                  "bnn_gemm_popc: out_dtype must be S32 or S16");
    BNN_CHECK_ARG(out_dtype == BNN_EPI_STORE_S32 || K < 32768, "bnn_gemm_popc: K too large for int16");
    dim3 grid(cdiv(N, PT), cdiv(M, PT));
    BNN_CHECK_ARG(grid.y <= 65535, "bnn_gemm_popc: M too large");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (out_dtype == BNN_EPI_STORE_S32)
        gemm_popc_kernel<int32_t><<<grid, 256, 0, s>>>(a_bits, lda_w, b_bits, ldb_w, M, N, K,
                                                       static_cast<int32_t*>(D), ldd);
    else
        gemm_popc_kernel<int16_t><<<grid, 256, 0, s>>>(a_bits, lda_w, b_bits, ldb_w, M, N, K,
                                                       static_cast<int16_t*>(D), ldd);
    BNN_LAUNCH_CHECK("gemm_popc_kernel");
    return BNN_OK;
}
