// Library-level entry points of the C ABI: error text, version, device query, GEMM dispatch.
#include "common.cuh"

#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace bnn {

namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return BNN_ERR_CUDA;
}

int sm_count() {
    // per-device cache; racing first calls from several host threads store the same value (relaxed atomics)
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

// K steps (of 128 bytes) per fp32 accumulation chunk of the tcgen05 GEMM; see gemm_tc.cu.
int gemm_chunk_iters() {  // read on every call (no cached state to race on; the tests vary it in one process)
    const char* e = getenv("BNN_GEMM_CHUNK");
    const int n = e ? atoi(e) : 16;
    return n < 1 ? 1 : (n > 4096 ? 4096 : n);
}

int gemm_tc_dispatch(const bnn_gemm_desc& g, cudaStream_t stream);
int gemm_simt_dispatch(const bnn_gemm_desc& g, cudaStream_t stream);

}  // namespace bnn

using namespace bnn;

extern "C" const char* bnn_last_error(void) { return g_err; }

extern "C" int bnn_version(void) { return BNN_B200_VERSION; }

extern "C" int bnn_device_query(int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
    int dev = 0;
    BNN_CUDA(cudaGetDevice(&dev));
    int sms = 0, maj = 0, mnr = 0;
    BNN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    BNN_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    BNN_CUDA(cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count_host) *sm_count_host = sms;
    if (cc_major_host) *cc_major_host = maj;
    if (cc_minor_host) *cc_minor_host = mnr;
    return BNN_OK;
}

extern "C" int bnn_zero_async(void* ptr, int64_t bytes, bnn_stream_t stream) {
    BNN_CHECK_ARG(bytes >= 0 && (ptr != nullptr || bytes == 0), "bnn_zero_async: bad input");
    if (bytes) BNN_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, static_cast<cudaStream_t>(stream)));
    return BNN_OK;
}

extern "C" int bnn_gemm(const bnn_gemm_desc* g, int impl, bnn_stream_t stream) {
    BNN_CHECK_ARG(g != nullptr, "bnn_gemm: null descriptor");
    BNN_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0 && g->batch >= 1,
                  "bnn_gemm: empty shape M=%d N=%d K=%d batch=%d", g->M, g->N, g->K, g->batch);
    BNN_CHECK_ARG(g->in_dtype == BNN_DT_I8 || g->in_dtype == BNN_DT_F16 || g->in_dtype == BNN_DT_F4,
                  "bnn_gemm: bad in_dtype %d", g->in_dtype);
    BNN_CHECK_ARG(g->epilogue >= BNN_EPI_STORE_F32 && g->epilogue <= BNN_EPI_SIGN_F4,
                  "bnn_gemm: bad epilogue %d", g->epilogue);
    const bool sign_epi = g->epilogue == BNN_EPI_SIGN_I8 || g->epilogue == BNN_EPI_SIGN_F4;
    BNN_CHECK_ARG((g->in_dtype != BNN_DT_F4 && g->epilogue != BNN_EPI_SIGN_F4) || impl == BNN_GEMM_TCGEN05,
                  "bnn_gemm: E2M1 operands and the nibble epilogue are tcgen05-path features");
    if (g->epilogue == BNN_EPI_SCATTER_F32) {
        const bnn_peer_scatter* sc = g->scatter;
        BNN_CHECK_ARG(sc != nullptr, "bnn_gemm: BNN_EPI_SCATTER_F32 needs desc.scatter");
        BNN_CHECK_ARG(sc->world >= 1 && sc->world <= BNN_MAX_RANKS && sc->rank >= 0 &&
                          sc->rank < sc->world,
                      "bnn_gemm: scatter world %d / rank %d out of range", sc->world, sc->rank);
        BNN_CHECK_ARG(sc->rows_per_owner > 0 && sc->rows_per_owner % 128 == 0 &&
                          (int64_t)sc->rows_per_owner * sc->world >= g->M,
                      "bnn_gemm: scatter rows_per_owner %d must be a multiple of 128 covering M=%d",
                      sc->rows_per_owner, g->M);
        for (int r = 0; r < sc->world; ++r)
            BNN_CHECK_ARG(sc->stage[r] != 0, "bnn_gemm: scatter stage[%d] is null", r);
        BNN_CHECK_ARG(g->batch == 1 && impl == BNN_GEMM_TCGEN05 &&
                          (g->in_dtype == BNN_DT_F16 || g->hi_from_pass > 0),
                      "bnn_gemm: BNN_EPI_SCATTER_F32 is a single tcgen05 product (fp16-split, or "
                      "digit-split int8)");
    }
    BNN_CHECK_ARG(!sign_epi || (g->epi_thr && g->epi_sgn),
                  "bnn_gemm: BNN_EPI_SIGN_I8 / SIGN_F4 need epi_thr and epi_sgn");
    BNN_CHECK_ARG(!sign_epi || (aligned16(g->epi_thr) && aligned16(g->epi_sgn)),
                  "bnn_gemm: epi_thr / epi_sgn must be 16-byte aligned");
    BNN_CHECK_ARG(g->epi_thr_stride == 0 ||
                      (sign_epi && g->epi_thr_stride >= g->N && g->epi_thr_stride % 4 == 0),
                  "bnn_gemm: epi_thr_stride %d must be 0 or a multiple of 4 >= N (sign epilogues only)",
                  g->epi_thr_stride);
    BNN_CHECK_ARG((g->a_unsigned == 0 || g->a_unsigned == 1) &&
                      (g->a_unsigned == 0 || (g->in_dtype == BNN_DT_I8 && g->hi_from_pass == 0)),
                  "bnn_gemm: a_unsigned is 0 or 1 and goes with plain int8-path products");
    BNN_CHECK_ARG(g->epilogue != BNN_EPI_STORE_S16 || !g->a_unsigned ||
                      255 * (int64_t)g->K * g->n_pass < 32768,
                  "bnn_gemm: K too large for an int16 result of a uint8 operand");
    BNN_CHECK_ARG(g->in_dtype != BNN_DT_F16 ||
                      (g->epilogue != BNN_EPI_STORE_S32 && g->epilogue != BNN_EPI_STORE_S16),
                  "bnn_gemm: integer stores need int8 or E2M1 operands");
    BNN_CHECK_ARG(g->epilogue != BNN_EPI_STORE_S16 || g->K * (int64_t)g->n_pass < 32768,
                  "bnn_gemm: K too large for an int16 result");
    BNN_CHECK_ARG(g->n_pass >= 1 && g->n_pass <= BNN_GEMM_MAX_PASS, "bnn_gemm: n_pass %d out of range",
                  g->n_pass);
    BNN_CHECK_ARG(g->A[0] && g->B[0] && (g->D || g->epilogue == BNN_EPI_SCATTER_F32),
                  "bnn_gemm: null operand");
    for (int p = 0; p < g->n_pass; ++p) {
        BNN_CHECK_ARG(g->pa[p] >= 0 && g->pa[p] <= 1 && g->pb[p] >= 0 && g->pb[p] <= 2,
                      "bnn_gemm: pass %d refers to a piece outside A{0,1} / B{0,1,2}", p);
        BNN_CHECK_ARG(g->A[g->pa[p]] && (g->pb[p] == 2 ? g->B2 : g->B[g->pb[p]]),
                      "bnn_gemm: pass %d uses a null piece", p);
    }
    BNN_CHECK_ARG(g->stats >= BNN_STATS_NONE && g->stats <= BNN_STATS_BN_BWD, "bnn_gemm: bad stats %d",
                  g->stats);
    if (g->stats != BNN_STATS_NONE) {
        BNN_CHECK_ARG(impl == BNN_GEMM_TCGEN05, "bnn_gemm: fused column statistics are a tcgen05 epilogue");
        BNN_CHECK_ARG(g->stats_sums != nullptr, "bnn_gemm: stats_sums is null");
        if (g->stats == BNN_STATS_BN_BWD) {
            BNN_CHECK_ARG(g->stats_max && g->stats_z && g->stats_mu && g->stats_ldz >= g->N,
                          "bnn_gemm: BNN_STATS_BN_BWD needs stats_max, stats_z (ldz >= N) and stats_mu");
            BNN_CHECK_ARG(g->stats_z_dtype >= BNN_Z_F32 && g->stats_z_dtype <= BNN_Z_S32,
                          "bnn_gemm: bad stats_z_dtype %d", g->stats_z_dtype);
        }
    }
    BNN_CHECK_ARG((g->hi_from_pass == 0 && g->col_scale == nullptr && g->B2 == nullptr) ||
                      impl == BNN_GEMM_TCGEN05,
                  "bnn_gemm: B2 / hi_from_pass / col_scale are tcgen05-path features");
    BNN_CHECK_ARG(g->ldd >= g->N, "bnn_gemm: ldd < N");
    BNN_CHECK_ARG(g->batch == 1 || g->stride_d >= (int64_t)g->M * g->ldd,
                  "bnn_gemm: batched output needs stride_d >= M*ldd");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (impl == BNN_GEMM_TCGEN05) return gemm_tc_dispatch(*g, s);
    if (impl == BNN_GEMM_SIMT) return gemm_simt_dispatch(*g, s);
    set_error("bnn_gemm: unknown impl %d", impl);
    return BNN_ERR_ARG;
}
