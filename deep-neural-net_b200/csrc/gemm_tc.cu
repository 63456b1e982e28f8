// tcgen05 GEMM for sm_100a: TMA-fed, TMEM-accumulated, warp-specialised, persistent.
//
//   D[b][M,N] = epilogue( sum_p A[pa[p]][b][M,K] * B[pb[p]][b][N,K]^T )
//
// One kernel serves every matrix product on the binarized-layer hot path (reference call sites:
// X.dot(weight) mlpcode/layers.py:220, input.T.dot(delta) :260, delta.dot(weights.T) :266, and the
// fused `weights -= lr*dw` :268):
//   * +-1 x +-1 forward:  int8 operands, tcgen05.mma kind::i8, int32 accumulators (bit-exact);
//   * real-valued products: fp16 hi/lo split operands, kind::f16, fp32 accumulators; the passes of
//     the split (hi*hi, hi*lo, lo*hi) are extra trips round the K loop into the SAME accumulator.
//
// CTA = 8 warps:  w0 lane0  TMA producer        (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//                 w1 lane0  MMA issuer          (tcgen05.mma, tcgen05.commit -> mbarriers)
//                 w2        TMEM allocator
//                 w4..w7    epilogue            (tcgen05.ld -> registers -> global), one TMEM lane
//                                                quarter each, overlapped with the next tile's
//                                                main loop through a 2-deep TMEM accumulator ring.
// Tile 128 x BLOCK_N, K step = 128 bytes (one swizzle atom row), 4 UMMAs (32 bytes of K) per step.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <mutex>

namespace bnn {

namespace {

constexpr int kBlockM = 128;
constexpr int kKBytes = 128;   // bytes of K per pipeline stage row (= swizzle span)
constexpr int kUmmaKBytes = 32;
constexpr int kThreads = 256;
constexpr int kAccStages = 2;

template <int BLOCK_N>
struct TileCfg {
    static constexpr int kStages = (BLOCK_N >= 256) ? 4 : ((BLOCK_N >= 128) ? 6 : 8);
    static constexpr int kABytes = kBlockM * kKBytes;
    static constexpr int kBBytes = BLOCK_N * kKBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = (kAccStages * BLOCK_N) < 32 ? 32 : (kAccStages * BLOCK_N);
    static constexpr int kBarBytes = (2 * kStages + 2 * kAccStages) * 8 + 16;
    static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // + align slack
};

struct TcParams {
    int M, N, batch;
    int kblocks;  // K steps per pass
    int n_pass;
    int pa[BNN_GEMM_MAX_PASS];
    int pb[BNN_GEMM_MAX_PASS];
    int m_tiles, n_tiles;
    int epilogue;
    int vec_ok;  // D base/pitch allow 16-byte vector access
    int a_batched, b_batched;
    void* D;
    long long ldd, stride_d;
    const float* sa;
    const float* sb;
    float host_scale;
};

__device__ __forceinline__ void tile_coords(const TcParams& p, int tile, int& b, int& m_blk,
                                            int& n_blk) {
    const int per_batch = p.m_tiles * p.n_tiles;
    b = tile / per_batch;
    int r = tile - b * per_batch;
    // groups of 8 M-blocks sweep all N-blocks: a wave of CTAs shares a small set of A and B tiles
    const int group = 8 * p.n_tiles;
    const int g = r / group;
    const int first_m = g * 8;
    const int gsz = min(p.m_tiles - first_m, 8);
    const int rr = r - g * group;
    m_blk = first_m + rr % gsz;
    n_blk = rr / gsz;
}

template <bool kI8, int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
               const TcParams p) {
    using Cfg = TileCfg<BLOCK_N>;
    constexpr int S = Cfg::kStages;
    constexpr int kKElems = kI8 ? kKBytes : kKBytes / 2;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + S;
    uint64_t* tfull_bar = empty_bar + S;
    uint64_t* tempty_bar = tfull_bar + kAccStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + kAccStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.batch * p.m_tiles * p.n_tiles;
    const int k_iters = p.n_pass * p.kblocks;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&mapA0);
        ptx::prefetch_tensormap(&mapB0);
        if (p.n_pass > 1) {
            ptx::prefetch_tensormap(&mapA1);
            ptx::prefetch_tensormap(&mapB1);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < kAccStages; ++a) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&tempty_bar[a], 128);
        }
        ptx::fence_mbar_init();
        ptx::fence_proxy_async();
    }
    if (warp == 2) ptx::tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------- TMA producer
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int b, m_blk, n_blk;
                tile_coords(p, tile, b, m_blk, n_blk);
                const int ba = p.a_batched ? b : 0;
                const int bb = p.b_batched ? b : 0;
                for (int pass = 0; pass < p.n_pass; ++pass) {
                    const CUtensorMap* ma = p.pa[pass] ? &mapA1 : &mapA0;
                    const CUtensorMap* mb = p.pb[pass] ? &mapB1 : &mapB0;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
                        uint8_t* sa = smem + s * Cfg::kStageBytes;
                        uint8_t* sb = sa + Cfg::kABytes;
                        ptx::mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
                        ptx::tma_load_3d(sa, ma, &full_bar[s], kb * kKElems, m_blk * kBlockM, ba);
                        ptx::tma_load_3d(sb, mb, &full_bar[s], kb * kKElems, n_blk * BLOCK_N, bb);
                        if (++s == S) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc = kI8 ? ptx::umma_idesc(2u, 1u, kBlockM, BLOCK_N)
                                           : ptx::umma_idesc(1u, 0u, kBlockM, BLOCK_N);
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                ptx::mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
                ptx::tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int it = 0; it < k_iters; ++it) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after_sync();
                    const uint32_t a_addr = ptx::smem_u32(smem + s * Cfg::kStageBytes);
                    const uint64_t da = ptx::umma_desc_k_sw128(a_addr);
                    const uint64_t db = ptx::umma_desc_k_sw128(a_addr + Cfg::kABytes);
#pragma unroll
                    for (int j = 0; j < kKBytes / kUmmaKBytes; ++j) {
                        const uint64_t adv = (uint64_t)((j * kUmmaKBytes) >> 4);
                        const uint32_t accum = (it > 0 || j > 0) ? 1u : 0u;
                        if (kI8)
                            ptx::mma_i8(d_tmem, da + adv, db + adv, idesc, accum);
                        else
                            ptx::mma_f16(d_tmem, da + adv, db + adv, idesc, accum);
                    }
                    ptx::mma_commit(&empty_bar[s]);  // frees the smem stage once the MMAs retire
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
                ptx::mma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
                if (++acc == kAccStages) { acc = 0; acc_ph ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ----------------------------------------------------------------- epilogue
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        float inv = 1.0f;
        if (p.sa) inv /= __ldg(p.sa);
        if (p.sb) inv /= __ldg(p.sb);
        const float scale = p.host_scale * inv;  // lr * 2^-e : exact
        int acc = 0;
        uint32_t acc_ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int b, m_blk, n_blk;
            tile_coords(p, tile, b, m_blk, n_blk);
            const int row = m_blk * kBlockM + q * 32 + lane;
            const bool row_ok = row < p.M;
            ptx::mbar_wait(&tfull_bar[acc], acc_ph);
            ptx::tc_fence_after_sync();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            char* dbase = reinterpret_cast<char*>(p.D);
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                const int col0 = n_blk * BLOCK_N + c * 32;
                if (col0 >= p.N) break;  // uniform across the CTA
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)(c * 32), v);
                ptx::tmem_ld_wait();
                if (!row_ok) continue;
                const long long off = (long long)b * p.stride_d + (long long)row * p.ldd + col0;
                const bool full = (col0 + 32 <= p.N) && p.vec_ok;
                if (p.epilogue == BNN_EPI_STORE_S16) {
                    int16_t* d = reinterpret_cast<int16_t*>(dbase) + off;
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 o;
                            o.x = (v[j + 0] & 0xffffu) | (v[j + 1] << 16);
                            o.y = (v[j + 2] & 0xffffu) | (v[j + 3] << 16);
                            o.z = (v[j + 4] & 0xffffu) | (v[j + 5] << 16);
                            o.w = (v[j + 6] & 0xffffu) | (v[j + 7] << 16);
                            *reinterpret_cast<uint4*>(d + j) = o;
                        }
                    } else {
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.N) d[j] = (int16_t)(int)v[j];
                    }
                } else if (p.epilogue == BNN_EPI_STORE_S32) {
                    int32_t* d = reinterpret_cast<int32_t*>(dbase) + off;
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<uint4*>(d + j) =
                                make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < p.N) d[j] = (int)v[j];
                    }
                } else {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float a = kI8 ? (float)(int)v[j] : __uint_as_float(v[j]);
                        f[j] = __fmul_rn(a, scale);
                    }
                    float* d = reinterpret_cast<float*>(dbase) + off;
                    if (p.epilogue == BNN_EPI_SGD_F32) {
                        // weights -= lr * dw  (layers.py:268): two roundings, like NumPy
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 w = *reinterpret_cast<float4*>(d + j);
                                w.x = __fsub_rn(w.x, f[j + 0]);
                                w.y = __fsub_rn(w.y, f[j + 1]);
                                w.z = __fsub_rn(w.z, f[j + 2]);
                                w.w = __fsub_rn(w.w, f[j + 3]);
                                *reinterpret_cast<float4*>(d + j) = w;
                            }
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < p.N) d[j] = __fsub_rn(d[j], f[j]);
                        }
                    } else {
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4*>(d + j) =
                                    make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < p.N) d[j] = f[j];
                        }
                    }
                }
            }
            ptx::tc_fence_before_sync();
            ptx::mbar_arrive(&tempty_bar[acc]);
            if (++acc == kAccStages) { acc = 0; acc_ph ^= 1u; }
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// Operand [batch][rows, K] with pitch ld (elements) -> 3-D map (K, rows, batch), box (128 B, box_rows, 1)
int make_operand_map(CUtensorMap* map, const void* base, bool i8, long long rows, long long K,
                     long long ld, long long batch, long long stride, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable");
        return BNN_ERR_CUDA;
    }
    const int es = i8 ? 1 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
    cuuint64_t strides[2] = {(cuuint64_t)(ld * es),
                             (cuuint64_t)((stride > 0 ? stride : rows * ld) * es)};
    cuuint32_t box[3] = {(cuuint32_t)(kKBytes / es), (cuuint32_t)box_rows, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(map, i8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                     const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d): rows=%lld K=%lld ld=%lld batch=%lld",
                  (int)r, rows, K, ld, batch);
        return BNN_ERR_CUDA;
    }
    return BNN_OK;
}

template <bool kI8, int BLOCK_N>
int launch_tc(const bnn_gemm_desc& g, cudaStream_t stream) {
    using Cfg = TileCfg<BLOCK_N>;
    static bool attr_set = false;
    if (!attr_set) {
        BNN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<kI8, BLOCK_N>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg::kSmemBytes));
        attr_set = true;
    }
    CUtensorMap mA[2], mB[2];
    const int nA = (g.A[1] != nullptr) ? 2 : 1;
    const int nB = (g.B[1] != nullptr) ? 2 : 1;
    for (int i = 0; i < 2; ++i) {
        const void* a = g.A[i < nA ? i : 0];
        const void* b = g.B[i < nB ? i : 0];
        int rc = make_operand_map(&mA[i], a, kI8, g.M, g.K, g.lda, g.stride_a ? g.batch : 1,
                                  g.stride_a, kBlockM);
        if (rc) return rc;
        rc = make_operand_map(&mB[i], b, kI8, g.N, g.K, g.ldb, g.stride_b ? g.batch : 1,
                              g.stride_b, BLOCK_N);
        if (rc) return rc;
    }
    TcParams p;
    p.M = g.M;
    p.N = g.N;
    p.batch = g.batch;
    const int kelems = kI8 ? kKBytes : kKBytes / 2;
    p.kblocks = cdiv(g.K, kelems);
    p.n_pass = g.n_pass;
    for (int i = 0; i < BNN_GEMM_MAX_PASS; ++i) {
        p.pa[i] = g.pa[i];
        p.pb[i] = g.pb[i];
    }
    p.m_tiles = cdiv(g.M, kBlockM);
    p.n_tiles = cdiv(g.N, BLOCK_N);
    p.epilogue = g.epilogue;
    const int esz = (g.epilogue == BNN_EPI_STORE_S16) ? 2 : 4;
    p.vec_ok = aligned16(g.D) && ((g.ldd * esz) % 16 == 0) && ((g.stride_d * esz) % 16 == 0);
    p.a_batched = g.stride_a != 0;
    p.b_batched = g.stride_b != 0;
    p.D = g.D;
    p.ldd = g.ldd;
    p.stride_d = g.stride_d;
    p.sa = g.sa;
    p.sb = g.sb;
    p.host_scale = g.host_scale;
    const long long tiles = (long long)p.batch * p.m_tiles * p.n_tiles;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    gemm_tc_kernel<kI8, BLOCK_N><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(mA[0], mA[1], mB[0],
                                                                             mB[1], p);
    BNN_LAUNCH_CHECK("gemm_tc_kernel");
    return BNN_OK;
}

}  // namespace

int gemm_tc_dispatch(const bnn_gemm_desc& g, cudaStream_t stream) {
    const bool i8 = g.in_dtype == BNN_DT_I8;
    const int es = i8 ? 1 : 2;
    for (int i = 0; i < 2; ++i) {
        if (g.A[i]) BNN_CHECK_ARG(aligned16(g.A[i]), "bnn_gemm: A[%d] not 16-byte aligned", i);
        if (g.B[i]) BNN_CHECK_ARG(aligned16(g.B[i]), "bnn_gemm: B[%d] not 16-byte aligned", i);
    }
    BNN_CHECK_ARG((g.lda * es) % 16 == 0 && (g.ldb * es) % 16 == 0,
                  "bnn_gemm: operand pitches must be multiples of 16 bytes (lda=%lld ldb=%lld)",
                  (long long)g.lda, (long long)g.ldb);
    BNN_CHECK_ARG((g.stride_a * es) % 16 == 0 && (g.stride_b * es) % 16 == 0,
                  "bnn_gemm: batch strides must be multiples of 16 bytes");
    BNN_CHECK_ARG(g.lda >= g.K && g.ldb >= g.K, "bnn_gemm: pitch smaller than K");
    if (g.N <= 32) return i8 ? launch_tc<true, 32>(g, stream) : launch_tc<false, 32>(g, stream);
    if (g.N <= 128) return i8 ? launch_tc<true, 128>(g, stream) : launch_tc<false, 128>(g, stream);
    return i8 ? launch_tc<true, 256>(g, stream) : launch_tc<false, 256>(g, stream);
}

}  // namespace bnn
