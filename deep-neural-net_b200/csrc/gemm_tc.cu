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
// CTA = 12 warps:  w0 lane0  TMA producer       (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//                  w1 lane0  MMA issuer         (tcgen05.mma, tcgen05.commit -> mbarriers)
//                  w2        TMEM allocator     (w3 idle)
//                  w4..w11   epilogue           (tcgen05.ld -> registers -> global): warp w owns TMEM
//                                                lane quarter w%4 and column HALF (w-4)/4 of the tile.
// setmaxnreg moves registers from the role warp group (56 each) to the two epilogue warp groups (224 each).
// Tile 128 x BLOCK_N, K step = 128 bytes (one swizzle atom row), 4 UMMAs (32 bytes of K) per step.
//
// Chunked accumulation.  Measured on B200: the TMEM fp32 accumulator TRUNCATES on every MMA, a bias of
// ~3e-8 (relative) per accumulation step that grows linearly with K (2.5e-5 after the 1024 steps of a
// K = 8192, 2-pass dW product; an fp32 SGEMM is at 1e-6).  The MMA issuer therefore closes the
// accumulator every `chunk_iters` K steps and moves to the other half of a 2-deep TMEM ring, while the
// epilogue warps drain the finished chunk and add it to a register-resident running sum with
// round-to-nearest fp32 adds.  Draining a chunk (128 KB of TMEM) is hidden behind the next chunk's MMAs.
// Measured at config-2 shapes against an fp64 product (tests/test_kernels_gpu.py): chunk = 4 / 8 / 16 /
// none -> 3.9e-7 / 7e-7 / 1.2e-6 / 2.5e-5 relative error (cuBLAS SGEMM: 1.2e-6) for +2.4 / +0.9 / +0.5 / 0 %
// of the dominant launches' time; the default of 16 K steps matches an fp32 SGEMM.
// Integer accumulation is exact, so kind::i8 products use a single chunk.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <atomic>
#include <mutex>
#include <stdlib.h>

namespace bnn {

int gemm_chunk_iters();  // api.cu: K steps per accumulation chunk (env BNN_GEMM_CHUNK, default 16)

namespace {

constexpr int kBlockM = 128;
constexpr int kKBytes = 128;   // bytes of K per pipeline stage row (= swizzle span)
constexpr int kUmmaKBytes = 32;
constexpr int kRoleThreads = 128;       // warp group 0: TMA, MMA, TMEM alloc, idle
constexpr int kEpiWarps = 8;            // warp groups 1-2
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kRoleThreads + kEpiThreads;
constexpr int kAccStages = 2;
// fp32 results leave through a per-warp shared-memory transpose so that every store instruction
// writes whole 128-byte row segments (the TMEM layout gives each lane one ROW, and row-per-lane
// stores are 32 scattered 16-byte pieces: tolerable into local L2, crippling over NVLink).
constexpr int kXposeRows = 16;    // half a warp's rows per phase
constexpr int kXposeStride = 36;  // floats per staged row (32 + 4: conflict-free 128-bit access)
// per-warp staging buffer: [0, 1024) the sign epilogues' thresholds, [1024, 3328) row staging of the integer
// epilogues (the fp32 epilogues stage their rows in [0, 2304))
constexpr int kRowStageOff = 1024;
constexpr int kRowStageBytes = kXposeRows * kXposeStride * 4;   // 2304
constexpr int kXposeWarpBytes = kRowStageOff + kRowStageBytes;  // 3328

// kCta2: the CTA-pair form keeps only HALF of the B tile in each CTA, so a stage is 32 KB instead of 48 KB and
// six of them fit where the single-CTA form holds four.
// kF4: block-scaled FP4 operands (kind::mxf4).  32 TMEM columns after the two accumulator stages hold the
// constant unit scale factors, so the widest tile is 224 columns (2 x 224 + 32 <= 512).
template <int BLOCK_N, bool kCta2 = false, bool kF4 = false>
struct TileCfg {
    static constexpr int kABytes = kBlockM * kKBytes;
    static constexpr int kBBytes = (kCta2 ? BLOCK_N / 2 : BLOCK_N) * kKBytes;   // per CTA
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = kCta2 ? 6 : ((BLOCK_N >= 200) ? 4 : ((BLOCK_N >= 128) ? 6 : 8));
    static constexpr int kSfCol = kAccStages * BLOCK_N;     // first scale-factor column (kF4)
    static constexpr int kSfCols = 32;
    static constexpr int kColsNeeded = kAccStages * BLOCK_N + (kF4 ? kSfCols : 0);
    static constexpr int kTmemCols = kColsNeeded <= 32 ? 32 : (kColsNeeded <= 64 ? 64 : (kColsNeeded <= 128 ? 128 :
                                     (kColsNeeded <= 256 ? 256 : 512)));
    static_assert(kColsNeeded <= 512, "accumulator stages + scale factors exceed TMEM");
    static constexpr int kBarBytes = (2 * kStages + 2 * kAccStages) * 8 + 16;
    static constexpr int kXposeBytes = kEpiWarps * kXposeWarpBytes;  // store staging + thresholds
    static constexpr int kSmemBytes =
        kStages * kStageBytes + kBarBytes + 16 + kXposeBytes + 1024;  // + align slack
    static constexpr int kColsPerThread = BLOCK_N / 2;                 // per epilogue thread
    static constexpr int kGroup = kColsPerThread < 16 ? kColsPerThread : 16;  // columns per tcgen05.ld
    static constexpr int kGroups = kColsPerThread / kGroup;
    static_assert(kColsPerThread % kGroup == 0, "tile width must split into whole tcgen05.ld groups");
};

struct TcParams {
    int M, N, batch;
    int kblocks;      // K steps per pass
    int n_pass;
    int chunk_iters;  // K steps per accumulation chunk
    int pa[BNN_GEMM_MAX_PASS];
    int pb[BNN_GEMM_MAX_PASS];
    int m_tiles, n_tiles;
    int group_m, group_n;  // tile walk (tile_coords)
    int mc;                // cluster-multicast form (launch_tc): CTA pairs share the B tile, see the kernel
    int debug;             // BNN_GEMM_DEBUG (timing experiments only): 1 = no TMA loads, 2 = no global stores
    int epilogue;
    int vec_ok;  // D base/pitch allow 16-byte vector access
    int a_batched, b_batched;
    void* D;
    long long ldd, stride_d;
    const float* sa;
    const float* sb;
    float host_scale;
    const float* thr;  // BNN_EPI_SIGN_I8
    const float* sgn;
    int thr_stride;    // elements between the threshold rows of consecutive batches (0 = shared)
    int a_unsigned;    // kind::i8: the A operand holds uint8 (u8 x s8 product)
    // BNN_EPI_SCATTER_F32: row tiles go to the staging buffer of the rank that owns them
    unsigned long long peer[BNN_MAX_RANKS];
    int rows_per_owner;
    int rank;
    int m_rot;  // row-tile rotation: this rank starts with the tiles it owns itself
    // digit-split int8 product: passes [0, hi_from_pass) accumulate the LOW accumulator, the rest the
    // HIGH one; the epilogue value is low + 256 * high (float).  0 = off.
    int hi_from_pass;
    const float* col_scale;  // optional per-column factor applied with `scale` (fp32 outputs)
    // column reductions fused into the epilogue (BNN_STATS_*)
    double* st_sums;       // [2, N] fp64 accumulators (pre-zeroed by the caller)
    uint32_t* st_max;      // [2, N] float bit patterns (BNN_STATS_BN_BWD)
    const void* st_Z;      // BNN_STATS_BN_BWD: pre-activations [M, st_ldz] of the batch-norm layer
    long long st_ldz;
    int st_zdtype;
    const float* st_mu;
};

// Tile walk.  The output is cut into SUPER-COLUMNS of p.group_n column tiles; inside one, groups of p.group_m row
// tiles sweep its column tiles with the row index running fastest.  CTAs that run concurrently (consecutive tile
// numbers) therefore share a small set of A and B tiles, and — what decides the DRAM traffic — the B tiles of a
// super-column (group_n x BLOCK_N rows x K x pieces) stay L2-resident while all of A streams past them once:
// DRAM reads = |A| x (number of super-columns) + |B|, instead of |A| + |B| x (row groups) when the whole of B
// does not fit in L2 (profiles/r02_summary.md: 4096-wide dX read 0.83-0.95 GB against 0.34 GB algorithmic).
__device__ __forceinline__ void tile_coords(const TcParams& p, int tile, int& b, int& m_blk,
                                            int& n_blk) {
    const int per_batch = p.m_tiles * p.n_tiles;
    b = tile / per_batch;
    int r = tile - b * per_batch;
    const int sc_tiles = p.m_tiles * p.group_n;       // tiles of a full super-column
    const int sc = r / sc_tiles;
    const int first_n = sc * p.group_n;
    const int ncols = min(p.n_tiles - first_n, p.group_n);
    r -= sc * sc_tiles;
    const int group = p.group_m * ncols;
    const int g = r / group;
    const int first_m = g * p.group_m;
    const int gsz = min(p.m_tiles - first_m, p.group_m);
    const int rr = r - g * group;
    m_blk = first_m + rr % gsz;
    n_blk = first_n + rr / gsz;
    if (p.m_rot) {  // ranks of a data-parallel group walk the row tiles from different starting points
        m_blk += p.m_rot;
        if (m_blk >= p.m_tiles) m_blk -= p.m_tiles;
    }
}

// Store / update one group of G consecutive columns of one output row from registers.
// sign epilogues: thr_smem is the shared-memory address of this warp's staged thresholds for the thread's column
// range (thr at +0, orientation at +kSignStageBytes/2; columns past N hold +inf / 1), cl the group's offset in it.
constexpr uint32_t kSignStageBytes = 1024;   // 2 x 128 floats: fits the warp's transpose buffer (2304 bytes)
// 16 sign decisions of one row: bit j = (acc_j * s_j >= thr_j) against the thresholds staged in shared memory.
template <bool kI8, bool kF4, int G>
__device__ __forceinline__ uint32_t sign_bits_group(const uint32_t (&v)[G], float scale, uint32_t thr_smem, int cl) {
    constexpr bool kIntCmp = kI8 && !kF4;
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < G; j += 4) {
        const uint4 tu = ptx::lds_v4(thr_smem + (uint32_t)((cl + j) * 4));
        const uint4 su = ptx::lds_v4(thr_smem + 512u + (uint32_t)((cl + j) * 4));
        const uint32_t tw[4] = {tu.x, tu.y, tu.z, tu.w}, sw[4] = {su.x, su.y, su.z, su.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (kIntCmp) {
                bits |= (uint32_t)((int)v[j + q] * (int)sw[q] >= (int)tw[q]) << (j + q);
            } else {
                const float a = kF4 ? __uint_as_float(v[j + q]) : __fmul_rn(__uint_as_float(v[j + q]), scale);
                bits |= (uint32_t)(__fmul_rn(a, __uint_as_float(sw[q])) >= __uint_as_float(tw[q])) << (j + q);
            }
        }
    }
    return bits;
}

// Integer epilogues of FULL-width tiles leave through shared memory too (round 2): a lane owns one output ROW, so
// direct stores are 32 pieces of 8-32 bytes per instruction, each in a different row — timing experiments put them
// at 2.6 us per tile whatever the output width (profiles/r02_summary.md), a quarter of an int8 product and more of an
// FP4 one.  Here every lane packs its row (ROWB bytes: int8 / nibble decisions or int16 sums) into registers, the
// rows of a phase are written to the warp's buffer, and all lanes copy them out so that consecutive lanes cover
// consecutive 16-byte (8-byte for 56-byte rows) pieces of one row: whole 64-256 byte row segments per instruction.
template <bool kI8, bool kF4, int EPI, int CPT>
__device__ __forceinline__ void staged_packed_rows(const TcParams& p, uint32_t xbuf, uint8_t* g_row0,
                                                   long long pitch_bytes, int rows_valid, const uint32_t (&sum)[CPT],
                                                   float scale, int lane) {
    constexpr int ROWB = EPI == BNN_EPI_STORE_S16 ? 2 * CPT : (EPI == BNN_EPI_SIGN_I8 ? CPT : CPT / 2);
    constexpr int W = ROWB / 4;                                   // packed words per row
    constexpr int PIECE = (ROWB % 16 == 0) ? 16 : 8;
    constexpr int PPR = ROWB / PIECE;                             // pieces per row
    // staged row pitch: the row plus padding that spreads a phase's rows over all banks
    constexpr int STRIDE = (PIECE == 8) ? ROWB + 16 : (((ROWB + 16) / 4) % 32 == 0 ? ROWB + 32 : ROWB + 16);
    constexpr int RP = (32 * STRIDE <= kRowStageBytes) ? 32 : ((16 * STRIDE <= kRowStageBytes) ? 16 : 8);
    static_assert(8 * STRIDE <= kRowStageBytes && ROWB % 8 == 0, "a phase of 8 rows must fit the staging buffer");
    const uint32_t sbuf = xbuf + (uint32_t)kRowStageOff;
    // ---- pack this lane's row into registers (sign epilogues; int16 rows are packed while they are staged)
    uint32_t pk[(EPI == BNN_EPI_STORE_S16) ? 1 : W];
    if constexpr (EPI != BNN_EPI_STORE_S16) {
        constexpr int G = 16;
#pragma unroll
        for (int g = 0; g < CPT / G; ++g) {
            uint32_t v[G];
#pragma unroll
            for (int j = 0; j < G; ++j) v[j] = sum[g * G + j];
            const uint32_t bits = sign_bits_group<kI8, kF4, G>(v, scale, xbuf, g * G);
            if constexpr (EPI == BNN_EPI_SIGN_F4) {
                pk[2 * g] = bits_to_e2m1x8(bits);
                pk[2 * g + 1] = bits_to_e2m1x8(bits >> 8);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) pk[4 * g + q] = nibble_to_pm1_bytes(bits >> (4 * q));
            }
        }
    }
#pragma unroll
    for (int ph = 0; ph < 32 / RP; ++ph) {
        if (lane / RP == ph) {
            const uint32_t w = sbuf + (uint32_t)((lane % RP) * STRIDE);
            if constexpr (EPI == BNN_EPI_STORE_S16) {
#pragma unroll
                for (int j = 0; j < CPT; j += 8)
                    ptx::sts_v4(w + (uint32_t)(2 * j), (sum[j] & 0xffffu) | (sum[j + 1] << 16),
                                (sum[j + 2] & 0xffffu) | (sum[j + 3] << 16), (sum[j + 4] & 0xffffu) | (sum[j + 5] << 16),
                                (sum[j + 6] & 0xffffu) | (sum[j + 7] << 16));
            } else if constexpr (PIECE == 16) {
#pragma unroll
                for (int j = 0; j < W; j += 4) ptx::sts_v4(w + (uint32_t)(4 * j), pk[j], pk[j + 1], pk[j + 2], pk[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < W; j += 2) ptx::sts_v2(w + (uint32_t)(4 * j), pk[j], pk[j + 1]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < (RP * PPR + 31) / 32; ++i) {
            const int pc = i * 32 + lane;
            const int r = pc / PPR, c = pc - r * PPR;
            const int row = ph * RP + r;
            if (pc < RP * PPR && row < rows_valid) {
                uint8_t* d = g_row0 + (long long)row * pitch_bytes + c * PIECE;
                if constexpr (PIECE == 16) {
                    const uint4 u = ptx::lds_v4(sbuf + (uint32_t)(r * STRIDE + c * 16));
                    *reinterpret_cast<uint4*>(d) = u;
                } else {
                    const uint2 u = ptx::lds_v2(sbuf + (uint32_t)(r * STRIDE + c * 8));
                    *reinterpret_cast<uint2*>(d) = u;
                }
            }
        }
        __syncwarp();
    }
}

// kIntCmp (kind::i8 sign epilogues): the staged thresholds are INTEGERS — ceil(thr) and the orientation as +-1 —
// and the decision is acc * s >= ceil(thr) on the int32 accumulator, identical to the float form for every
// integer acc and free of the int -> float conversions (quarter-rate pipe) that the float form needs.
// kFloatAcc (kind::mxf4 sign epilogues): v holds the fp32 accumulator bits (exact integers), compared as floats.
template <bool kI8, bool kF4, int EPI, int G>
__device__ __forceinline__ void store_group(const TcParams& p, char* dbase, long long off, int col0,
                                            const uint32_t (&v)[G], float scale, uint32_t thr_smem, int cl) {
    constexpr bool kIntCmp = kI8 && !kF4;
    const bool full = (col0 + G <= p.N) && p.vec_ok;
    if (EPI == BNN_EPI_SIGN_I8 || EPI == BNN_EPI_SIGN_F4) {
        // batch-norm + sign as one compare against the per-column threshold (bn.cu).  The thresholds come from
        // shared memory (staged once per tile with coalesced loads): fetched per group straight from global
        // memory, 56 dependent L2 round trips per tile made this epilogue 3x longer than an FP4 main loop.
        int8_t* d = reinterpret_cast<int8_t*>(dbase) + off;
        const uint32_t bits = sign_bits_group<kI8, kF4, G>(v, scale, thr_smem, cl);
        if (EPI == BNN_EPI_SIGN_F4) {
            // E2M1 nibbles, two columns per byte (low nibble = even column): +1 = 0x2, -1 = 0xA, columns past
            // N = 0x0 (+0.0: the consumer's K padding).  Whole 8-column words; the host guarantees the pitch
            // covers them (ldd % 32 == 0) and 16-byte alignment of D.
            uint8_t* d4 = reinterpret_cast<uint8_t*>(dbase) + (off >> 1);
            const int nvalid = p.N - col0;   // columns of this group inside the matrix
            uint32_t w[G / 8];
#pragma unroll
            for (int j = 0; j < G; j += 8) {
                uint32_t word = bits_to_e2m1x8(bits >> j);   // decision bit set -> 0x2 (+1), clear -> 0xA (-1)
                const int live = nvalid - j;                                        // nibbles past N -> 0
                if (live < 8) word &= live <= 0 ? 0u : (0xFFFFFFFFu >> (32 - 4 * live));
                w[j / 8] = word;
            }
            if (G == 16) *reinterpret_cast<uint2*>(d4) = make_uint2(w[0], w[G / 8 - 1]);
            else
#pragma unroll
                for (int j = 0; j < G / 8; ++j) *reinterpret_cast<uint32_t*>(d4 + 4 * j) = w[j];
        } else if (full) {
#pragma unroll
            for (int j = 0; j < G; j += 16) {
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    w[q] = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        w[q] |= (((bits >> (j + 4 * q + b)) & 1u) ? 0x01u : 0xFFu) << (8 * b);
                }
                *reinterpret_cast<uint4*>(d + j) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < G; ++j)
                if (col0 + j < p.N) d[j] = ((bits >> j) & 1u) ? (int8_t)1 : (int8_t)-1;
        }
    } else if (EPI == BNN_EPI_STORE_S16) {
        int16_t* d = reinterpret_cast<int16_t*>(dbase) + off;
        if (full) {
#pragma unroll
            for (int j = 0; j < G; j += 8) {
                uint4 o;
                o.x = (v[j + 0] & 0xffffu) | (v[j + 1] << 16);
                o.y = (v[j + 2] & 0xffffu) | (v[j + 3] << 16);
                o.z = (v[j + 4] & 0xffffu) | (v[j + 5] << 16);
                o.w = (v[j + 6] & 0xffffu) | (v[j + 7] << 16);
                *reinterpret_cast<uint4*>(d + j) = o;
            }
        } else {
#pragma unroll
            for (int j = 0; j < G; ++j)
                if (col0 + j < p.N) d[j] = (int16_t)(int)v[j];
        }
    } else if (EPI == BNN_EPI_STORE_S32) {
        int32_t* d = reinterpret_cast<int32_t*>(dbase) + off;
        if (full) {
#pragma unroll
            for (int j = 0; j < G; j += 4)
                *reinterpret_cast<uint4*>(d + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < G; ++j)
                if (col0 + j < p.N) d[j] = (int)v[j];
        }
    } else {
        float* d = reinterpret_cast<float*>(dbase) + off;
        constexpr bool sgd = EPI == BNN_EPI_SGD_F32;
        if (full) {
#pragma unroll
            for (int j = 0; j < G; j += 4) {
                float4 f;
                f.x = __fmul_rn(kI8 ? (float)(int)v[j + 0] : __uint_as_float(v[j + 0]), scale);
                f.y = __fmul_rn(kI8 ? (float)(int)v[j + 1] : __uint_as_float(v[j + 1]), scale);
                f.z = __fmul_rn(kI8 ? (float)(int)v[j + 2] : __uint_as_float(v[j + 2]), scale);
                f.w = __fmul_rn(kI8 ? (float)(int)v[j + 3] : __uint_as_float(v[j + 3]), scale);
                if (sgd) {  // weights -= lr * dw (layers.py:268): two roundings, like NumPy
                    const float4 w = *reinterpret_cast<const float4*>(d + j);
                    f.x = __fsub_rn(w.x, f.x);
                    f.y = __fsub_rn(w.y, f.y);
                    f.z = __fsub_rn(w.z, f.z);
                    f.w = __fsub_rn(w.w, f.w);
                }
                *reinterpret_cast<float4*>(d + j) = f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < G; ++j) {
                if (col0 + j < p.N) {
                    const float f = __fmul_rn(kI8 ? (float)(int)v[j] : __uint_as_float(v[j]), scale);
                    d[j] = sgd ? __fsub_rn(d[j], f) : f;
                }
            }
        }
    }
}

// One warp's 32 rows x CPT columns (registers: lane = row) pass through the warp's transpose buffer:
// two phases of 16 rows; per phase and 32-column group the 16 source lanes write their 8 float4, then
//   * kStore: all 32 lanes read 4 float4 each so that 8 adjacent lanes cover the 128 bytes of one row
//     segment (EPI: STORE_F32 / SCATTER_F32 store, SGD_F32 subtracts from the destination);
//   * STATS: lane c reads column c of the staged rows and accumulates the column reductions that
//     the batch-norm kernels would otherwise re-read the whole matrix for:
//       BNN_STATS_COLSUM  sum v, sum v^2                       (bnn_colstats,      layers.py:81-82)
//       BNN_STATS_BN_BWD  sum v, sum v*(Z-mu), max|v|, max|Z-mu| (bnn_bn_bwd_reduce, layers.py:107-115)
//     32-row partials (exact integers for int8 products, fp32 otherwise) go to the fp64 column
//     accumulators with one atomicAdd per column and warp.
// The value v is what the epilogue stores: acc * scale for fp32 outputs, the raw int32 sum for
// integer outputs.
// (one group of GC <= 32 columns starting at column g0 of the warp's CPT; g0 is a constant once the caller's loop
// is unrolled, so sum[g0 + j] stays in registers)
template <bool kI8, int EPI, int CPT, int STATS, bool kStore, int GC>
__device__ __forceinline__ void staged_group(const TcParams& p, uint32_t xbuf, float* dst_row0,
                                             int row_w0, int col_w0, const uint32_t (&sum)[CPT],
                                             float scale, int lane, const int g0) {
    constexpr int LPR = GC / 4;               // lanes per row segment
    constexpr int RPI = 32 / LPR;             // rows per store instruction
    constexpr bool sgd = EPI == BNN_EPI_SGD_F32;
    constexpr bool kIntVal = EPI == BNN_EPI_STORE_S16 || EPI == BNN_EPI_STORE_S32;  // v is an integer
    const int rows_valid = p.M - row_w0;
    {
        // per-group column accumulators of lane c (STATS only)
        int si = 0;
        unsigned long long qi = 0ull;
        float f1 = 0.f, f2 = 0.f, md = 0.f, mx = 0.f;
        const int col = col_w0 + g0 + lane;  // this lane's column in the stats phase
        float mu = 0.f;
        if (STATS == BNN_STATS_BN_BWD && lane < GC && col < p.N) mu = __ldg(p.st_mu + col);
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
            // BN_BWD: this lane's column of Z for the 16 rows of the phase is requested first, so the
            // loads are in flight while the tile is staged and stored
            // (raw bits only: converting here would make every load wait for the one before it).
            // Code size matters — this function is unrolled 8 times per tile and the kernel must stay
            // within the instruction cache — so the fast path assumes 16 valid rows and a valid
            // column (pointer increments, no predicates); ragged bottom tiles take a rolled loop.
            int zr[kXposeRows];
            bool zfast = false;
            int rmax = 0;
            const char* zp = nullptr;
            long long zstride = 0;
            if (STATS == BNN_STATS_BN_BWD) {
                const bool colok = lane < GC && col < p.N;
                rmax = colok ? max(0, min(kXposeRows, rows_valid - ph * kXposeRows)) : 0;
                const int esz = p.st_zdtype == BNN_Z_S16 ? 2 : 4;
                zstride = p.st_ldz * esz;
                zp = reinterpret_cast<const char*>(p.st_Z) +
                     ((long long)(row_w0 + ph * kXposeRows) * p.st_ldz + col) * esz;
                zfast = __all_sync(0xffffffffu, rmax == kXposeRows);
                if (zfast) {
                    const char* q = zp;
                    if (esz == 2) {
#pragma unroll
                        for (int r = 0; r < kXposeRows; ++r, q += zstride)
                            zr[r] = (int)__ldg(reinterpret_cast<const int16_t*>(q));
                    } else {  // int32 and fp32 travel as 32-bit patterns
#pragma unroll
                        for (int r = 0; r < kXposeRows; ++r, q += zstride)
                            zr[r] = __ldg(reinterpret_cast<const int32_t*>(q));
                    }
                }
            }
            if ((lane >> 4) == ph) {
                const uint32_t w = xbuf + (uint32_t)((lane & 15) * kXposeStride * 4);
#pragma unroll
                for (int j = 0; j < GC; j += 4) {
                    uint4 u;
                    if (kIntVal) {
                        u = make_uint4(sum[g0 + j], sum[g0 + j + 1], sum[g0 + j + 2],
                                       sum[g0 + j + 3]);
                    } else {
                        // digit-split int8 products already hold float bits (low + 256 * high)
                        const bool fbits = !kI8 || p.hi_from_pass != 0;
                        float4 cs = make_float4(1.f, 1.f, 1.f, 1.f);
                        if (p.col_scale)  // per-column power of two (exact), then the scalar factor
                            cs = __ldg(reinterpret_cast<const float4*>(p.col_scale + col_w0 + g0 + j));
                        const uint32_t r0 = sum[g0 + j + 0], r1 = sum[g0 + j + 1],
                                       r2 = sum[g0 + j + 2], r3 = sum[g0 + j + 3];
                        u.x = __float_as_uint(__fmul_rn(__fmul_rn(fbits ? __uint_as_float(r0) : (float)(int)r0, cs.x), scale));
                        u.y = __float_as_uint(__fmul_rn(__fmul_rn(fbits ? __uint_as_float(r1) : (float)(int)r1, cs.y), scale));
                        u.z = __float_as_uint(__fmul_rn(__fmul_rn(fbits ? __uint_as_float(r2) : (float)(int)r2, cs.z), scale));
                        u.w = __float_as_uint(__fmul_rn(__fmul_rn(fbits ? __uint_as_float(r3) : (float)(int)r3, cs.w), scale));
                    }
                    ptx::sts_v4(w + (uint32_t)(j * 4), u.x, u.y, u.z, u.w);
                }
            }
            __syncwarp();
            if (kStore) {
#pragma unroll
                for (int i = 0; i < kXposeRows / RPI; ++i) {
                    const int r = lane / LPR + RPI * i;       // row within the phase
                    const int c = (lane % LPR) * 4;           // column within the group
                    const int row = ph * kXposeRows + r;      // row within the warp's 32
                    const uint4 fu = ptx::lds_v4(xbuf + (uint32_t)((r * kXposeStride + c) * 4));
                    float4 f = make_float4(__uint_as_float(fu.x), __uint_as_float(fu.y),
                                           __uint_as_float(fu.z), __uint_as_float(fu.w));
                    if (row < rows_valid) {
                        float* d = dst_row0 + (long long)row * p.ldd + g0 + c;
                        if (sgd) {  // weights -= lr * dw (layers.py:268): two roundings, like NumPy
                            const float4 wv = *reinterpret_cast<const float4*>(d);
                            f.x = __fsub_rn(wv.x, f.x);
                            f.y = __fsub_rn(wv.y, f.y);
                            f.z = __fsub_rn(wv.z, f.z);
                            f.w = __fsub_rn(wv.w, f.w);
                        }
                        *reinterpret_cast<float4*>(d) = f;
                    }
                }
            }
            if (STATS != BNN_STATS_NONE && lane < GC && col < p.N) {
                if (STATS == BNN_STATS_COLSUM) {
                    // rows past M hold zeros (TMA zero-fill), so no row guard is needed for the sums
#pragma unroll
                    for (int r = 0; r < kXposeRows; ++r) {
                        const uint32_t u = ptx::lds_b32(xbuf + (uint32_t)((r * kXposeStride + lane) * 4));
                        if (kIntVal) {
                            const int v = (int)u;
                            si += v;
                            qi += (unsigned long long)((long long)v * (long long)v);
                        } else {
                            const float v = __uint_as_float(u);
                            f1 += v;
                            f2 = fmaf(v, v, f2);
                        }
                    }
                } else {
                    const bool zfloat = p.st_zdtype == BNN_Z_F32;
                    if (zfast) {
#pragma unroll
                        for (int r = 0; r < kXposeRows; ++r) {
                            const float d = __uint_as_float(
                                ptx::lds_b32(xbuf + (uint32_t)((r * kXposeStride + lane) * 4)));
                            const float z = zfloat ? __int_as_float(zr[r]) : (float)zr[r];
                            const float xm = z - mu;
                            f1 += d;
                            f2 = fmaf(d, xm, f2);
                            md = fmaxf(md, fabsf(d));
                            mx = fmaxf(mx, fabsf(xm));
                        }
                    } else {
#pragma unroll 1
                        for (int r = 0; r < rmax; ++r) {
                            const char* q = zp + (long long)r * zstride;
                            float z;
                            if (p.st_zdtype == BNN_Z_S16) z = (float)__ldg(reinterpret_cast<const int16_t*>(q));
                            else if (zfloat) z = __ldg(reinterpret_cast<const float*>(q));
                            else z = (float)__ldg(reinterpret_cast<const int32_t*>(q));
                            const float d = __uint_as_float(
                                ptx::lds_b32(xbuf + (uint32_t)((r * kXposeStride + lane) * 4)));
                            const float xm = z - mu;
                            f1 += d;
                            f2 = fmaf(d, xm, f2);
                            md = fmaxf(md, fabsf(d));
                            mx = fmaxf(mx, fabsf(xm));
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (STATS != BNN_STATS_NONE && lane < GC && col < p.N) {
            double a, b;
            if (STATS == BNN_STATS_COLSUM && kIntVal) { a = (double)si; b = (double)qi; }
            else { a = (double)f1; b = (double)f2; }
            atomicAdd(p.st_sums + col, a);
            atomicAdd(p.st_sums + p.N + col, b);
            if (STATS == BNN_STATS_BN_BWD) {
                if (md > 0.f) atomicMax(p.st_max + col, __float_as_uint(md));
                if (mx > 0.f) atomicMax(p.st_max + p.N + col, __float_as_uint(mx));
            }
        }
    }
}

template <bool kI8, int EPI, int CPT, int STATS, bool kStore>
__device__ __forceinline__ void staged_rows(const TcParams& p, uint32_t xbuf, float* dst_row0,
                                            int row_w0, int col_w0, const uint32_t (&sum)[CPT],
                                            float scale, int lane) {
    constexpr int GC = CPT < 32 ? CPT : 32;   // columns per group
    constexpr int kTail = CPT % GC;           // 224-wide tiles: 112 columns per warp = 3 x 32 + 16
#pragma unroll
    for (int g = 0; g < CPT / GC; ++g)
        staged_group<kI8, EPI, CPT, STATS, kStore, GC>(p, xbuf, dst_row0, row_w0, col_w0, sum, scale, lane, g * GC);
    if constexpr (kTail != 0)
        staged_group<kI8, EPI, CPT, STATS, kStore, kTail>(p, xbuf, dst_row0, row_w0, col_w0, sum, scale, lane,
                                                          (CPT / GC) * GC);
}

// kCta2: the CTA-pair form (cta_group::2, opt-in with BNN_GEMM_CTA2=1, see launch_tc_pair).  Two CTAs of a
// cluster share one 256 x BLOCK_N tile: each loads its own 128 rows of A and HALF of the B tile, the leader
// issues UMMAs of M = 256 that read both halves, each CTA's TMEM receives its own 128 rows and is drained
// by that CTA's epilogue warps exactly as in the single-CTA form.  The B tile then crosses L2 -> shared
// memory once per pair instead of once per CTA.  p.m_tiles counts PAIRS of row tiles in this form.
// kF4 (with kI8 = true: the epilogues see exact integers): E2M1 operands on kind::mxf4.block_scale with unit scale
// factors — every +-1 x +-1 product (training forward with fused column sums, sweeps, inference) at twice the
// kind::i8 MMA rate and half the operand bytes.
template <bool kI8, int BLOCK_N, int EPI, int STATS, bool kCta2 = false, bool kF4 = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
               const __grid_constant__ CUtensorMap mapB2, const TcParams p) {
    using Cfg = TileCfg<BLOCK_N, kCta2, kF4>;
    constexpr int S = Cfg::kStages;
    constexpr int kKElems = kI8 ? kKBytes : kKBytes / 2;   // of the tensor map's element type (bytes for i8 / f4)
    static_assert(!kF4 || (kI8 && !kCta2 && STATS != BNN_STATS_BN_BWD), "FP4 form: integer epilogues, single CTA");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + S;
    uint64_t* tfull_bar = empty_bar + S;
    uint64_t* tempty_bar = tfull_bar + kAccStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + kAccStages);
    float* xpose = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + Cfg::kBarBytes + 16 -
                                            ((S * Cfg::kStageBytes + Cfg::kBarBytes) & 15));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.batch * p.m_tiles * p.n_tiles;
    // Multicast form (p.mc, run-time, single-CTA MMAs): the two CTAs of a cluster work on vertically adjacent
    // 128 x BLOCK_N tiles, which read the SAME B tile.  Each CTA fetches its own A tile and HALF of the B tile, and
    // the half is delivered to both CTAs by TMA multicast — 32 KB instead of 48 KB leave the L2 per CTA and K step
    // (profiles/r02_summary.md: the int8 / FP4 products are bound by operand delivery).  A stage may be refilled only
    // when BOTH CTAs have consumed it (the peer writes into it), so the empty barriers count two arrivals and every
    // MMA issuer commits onto both CTAs' barriers.  p.m_tiles counts PAIRS of row tiles in this form.
    const bool mc = !kCta2 && p.mc != 0;
    const uint32_t cta_rank = (kCta2 || mc) ? ptx::cluster_ctarank() : 0u;
    const bool paired = kCta2 || mc;
// a CTA's (pair: a CTA pair's) walk over the output tiles
#define BNN_TILE_LOOP                                                                       \
    for (int tile = paired ? (blockIdx.x >> 1) : blockIdx.x; tile < total_tiles;            \
         tile += paired ? (gridDim.x >> 1) : gridDim.x)
    const int k_iters = p.n_pass * p.kblocks;
    // digit-split products have exactly two "chunks": the low and the high accumulator
    const int split_iters = p.hi_from_pass * p.kblocks;
    const int n_chunks = split_iters ? 2 : (k_iters + p.chunk_iters - 1) / p.chunk_iters;

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
            ptx::mbar_init(&full_bar[s], kCta2 ? 2 : 1);  // pair: the leader's arrive.expect_tx + the peer
            ptx::mbar_init(&empty_bar[s], mc ? 2 : 1);    // multicast: both CTAs' MMAs have read the stage
        }
        for (int a = 0; a < kAccStages; ++a) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&tempty_bar[a], kCta2 ? 2 * kEpiWarps : kEpiWarps);  // pair: both CTAs' warps
        }
        ptx::fence_mbar_init();
        ptx::fence_proxy_async();
    }
    if (warp == 2) ptx::tmem_alloc_t<kCta2, Cfg::kTmemCols>(tmem_ptr);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync_if<kCta2>();  // pair: the peer's barriers are initialised before anyone signals them
    if (mc) ptx::cluster_sync();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr;
    if constexpr (kF4) {
        // unit scale factors (UE8M0 0x7F = 2^0) in every byte of the 32 columns behind the accumulators, all
        // 128 lanes: warps 4-7 own the four lane quarters
        if (warp >= 4 && warp < 8) {
            for (int c = 0; c < Cfg::kSfCols; ++c)
                ptx::tmem_st_32x32b_x1(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(Cfg::kSfCol + c),
                                       0x7F7F7F7Fu);
            ptx::tmem_st_wait();
        }
        ptx::tc_fence_before_sync();
        __syncthreads();
        ptx::tc_fence_after_sync();
    }

    if (warp < 4) {
        // give the registers this warp group does not need to the epilogue warp groups.  The budget
        // (128 x 56 + 256 x 224 = 64512) leaves 1024 registers of the SM's 64K unclaimed: asking for
        // the full 65536 made tcgen05-era hardware spin forever in USETMAXREG.TRY_ALLOC.
        ptx::setmaxnreg_dec<56>();
        if (warp == 0 && lane == 0) {
            // ------------------------------------------------------------- TMA producer
            int s = 0;
            uint32_t ph = 0;
            BNN_TILE_LOOP {
                int b, m_blk, n_blk;
                tile_coords(p, tile, b, m_blk, n_blk);
                if (paired) m_blk = 2 * m_blk + (int)cta_rank;
                const int ba = p.a_batched ? b : 0;
                const int bb = p.b_batched ? b : 0;
                for (int pass = 0; pass < p.n_pass; ++pass) {
                    const CUtensorMap* ma = p.pa[pass] ? &mapA1 : &mapA0;
                    const CUtensorMap* mb = p.pb[pass] == 2 ? &mapB2 : (p.pb[pass] ? &mapB1 : &mapB0);
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
                        uint8_t* sa = smem + s * Cfg::kStageBytes;
                        uint8_t* sb = sa + Cfg::kABytes;
                        // pair: both CTAs' bytes are counted on the LEADER's barrier, each CTA loads its own
                        // 128 rows of A and its half of the B tile, the peer then arrives on that barrier
                        if (p.debug & 1) {   // timing experiment: the MMAs run on whatever the stage holds
                            if (!kCta2 || cta_rank == 0) ptx::mbar_arrive(&full_bar[s]);
                            else ptx::mbar_arrive_remote(&full_bar[s], 0u);
                            if (++s == S) { s = 0; ph ^= 1u; }
                            continue;
                        }
                        if (!kCta2 || cta_rank == 0)
                            ptx::mbar_arrive_expect_tx(&full_bar[s],
                                                       kCta2 ? 2u * Cfg::kStageBytes : Cfg::kStageBytes);
                        ptx::tma_load_3d_t<kCta2>(sa, ma, &full_bar[s], kb * kKElems, m_blk * kBlockM, ba);
                        if (mc)   // this CTA's half of the B tile, to the same place in both CTAs
                            ptx::tma_load_3d_mc(sb + (int)cta_rank * (BLOCK_N / 2) * kKBytes, mb, &full_bar[s],
                                                kb * kKElems, n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / 2), bb,
                                                (uint16_t)3);
                        else
                        ptx::tma_load_3d_t<kCta2>(sb, mb, &full_bar[s], kb * kKElems,
                                                  kCta2 ? n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / 2)
                                                        : n_blk * BLOCK_N, bb);
                        if (kCta2 && cta_rank != 0) ptx::mbar_arrive_remote(&full_bar[s], 0u);
                        if (++s == S) { s = 0; ph ^= 1u; }
                    }
                }
            }
        } else if (warp == 1 && lane == 0 && (!kCta2 || cta_rank == 0)) {
            // ------------------------------------------------------------- MMA issuer (pair: the leader)
            uint32_t idesc = kF4 ? ptx::umma_idesc_mxf4(kBlockM, BLOCK_N)
                           : (kI8 ? ptx::umma_idesc(2u, 1u, kCta2 ? 2 * kBlockM : kBlockM, BLOCK_N)
                                  : ptx::umma_idesc(1u, 0u, kCta2 ? 2 * kBlockM : kBlockM, BLOCK_N));
            const uint32_t sf_tmem = tmem_base + (uint32_t)Cfg::kSfCol;
            // kind::i8 operand formats: 1 = signed, 0 = unsigned 8-bit; A and B are independent fields
            if (kI8 && !kF4 && p.a_unsigned) idesc &= ~(7u << 7);
            int s = 0;
            uint32_t ph = 0;
            int acc = 0;
            uint32_t acc_ph = 0;
            BNN_TILE_LOOP {
                int it = 0;
                for (int c = 0; c < n_chunks; ++c) {
                    ptx::mbar_wait(&tempty_bar[acc], acc_ph ^ 1u);
                    ptx::tc_fence_after_sync();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                    const int it_end = split_iters ? (c == 0 ? split_iters : k_iters)
                                                   : min(it + p.chunk_iters, k_iters);
                    for (bool first = true; it < it_end; ++it) {
                        ptx::mbar_wait(&full_bar[s], ph);
                        ptx::tc_fence_after_sync();
                        const uint32_t a_addr = ptx::smem_u32(smem + s * Cfg::kStageBytes);
                        const uint64_t da = ptx::umma_desc_k_sw128(a_addr);
                        const uint64_t db = ptx::umma_desc_k_sw128(a_addr + Cfg::kABytes);
#pragma unroll
                        for (int j = 0; j < kKBytes / kUmmaKBytes; ++j) {
                            const uint64_t adv = (uint64_t)((j * kUmmaKBytes) >> 4);
                            const uint32_t accum = (!first || j > 0) ? 1u : 0u;
                            if (kF4)
                                ptx::mma_mxf4(d_tmem, da + adv, db + adv, idesc, accum, sf_tmem, sf_tmem);
                            else if (kI8)
                                ptx::mma_i8_t<kCta2>(d_tmem, da + adv, db + adv, idesc, accum);
                            else
                                ptx::mma_f16_t<kCta2>(d_tmem, da + adv, db + adv, idesc, accum);
                        }
                        first = false;
                        if (mc) ptx::mma_commit_mc(&empty_bar[s], (uint16_t)3);   // in both CTAs
                        else
                        ptx::mma_commit_t<kCta2>(&empty_bar[s]);  // frees the smem stage once the MMAs retire
                        if (++s == S) { s = 0; ph ^= 1u; }        // (pair: in both CTAs)
                    }
                    ptx::mma_commit_t<kCta2>(&tfull_bar[acc]);  // chunk complete -> epilogue (pair: both CTAs')
                    if (++acc == kAccStages) { acc = 0; acc_ph ^= 1u; }
                }
            }
        }
    } else {
        // ----------------------------------------------------------------- epilogue (8 warps)
        ptx::setmaxnreg_inc<224>();
        constexpr int CPT = Cfg::kColsPerThread, G = Cfg::kGroup, NG = Cfg::kGroups;
        const int q = warp & 3;          // TMEM lane quarter this warp may access
        const int h = (warp - 4) >> 2;   // column half of the tile
        float inv = 1.0f;
        if (p.sa) inv /= __ldg(p.sa);
        if (p.sb) inv /= __ldg(p.sb);
        const float scale = p.host_scale * inv;  // lr * 2^-e : exact
        int acc = 0;
        uint32_t acc_ph = 0;
        BNN_TILE_LOOP {
            uint32_t sum[CPT];  // running sum of the chunks (fp32 bit patterns, or int32)
            int b, m_blk, n_blk;
            tile_coords(p, tile, b, m_blk, n_blk);
            if (paired) m_blk = 2 * m_blk + (int)cta_rank;
            const int row = m_blk * kBlockM + q * 32 + lane;
            const int colbase = n_blk * BLOCK_N + h * CPT;
            const int row_w0 = m_blk * kBlockM + q * 32;  // first row of this warp
            const uint32_t xbuf = ptx::smem_u32(xpose) + (uint32_t)((warp - 4) * kXposeWarpBytes);
            if constexpr (EPI == BNN_EPI_SIGN_I8 || EPI == BNN_EPI_SIGN_F4) {
                // stage this warp's thresholds (thr, orientation) of the tile's column range while the MMAs run:
                // lane l fetches columns 4l .. 4l+3 (coalesced), columns past N get +inf / 1
                static_assert(2 * CPT * 4 <= (int)kSignStageBytes && (int)kSignStageBytes <= kRowStageOff,
                              "threshold staging must fit its part of the warp's buffer");
                const float* tb = p.thr + (long long)b * p.thr_stride;
                __syncwarp();   // the previous tile's reads of the buffer are done
                if (4 * lane < CPT) {
                    const int c0 = colbase + 4 * lane;
                    float4 tv, sv;
                    if (c0 + 4 <= p.N) {   // bases are 16-byte aligned (checked by bnn_gemm), strides multiples of 4
                        tv = __ldg(reinterpret_cast<const float4*>(tb + c0));
                        sv = __ldg(reinterpret_cast<const float4*>(p.sgn + c0));
                    } else {
                        tv.x = c0 + 0 < p.N ? __ldg(tb + c0 + 0) : INFINITY;  sv.x = c0 + 0 < p.N ? __ldg(p.sgn + c0 + 0) : 1.f;
                        tv.y = c0 + 1 < p.N ? __ldg(tb + c0 + 1) : INFINITY;  sv.y = c0 + 1 < p.N ? __ldg(p.sgn + c0 + 1) : 1.f;
                        tv.z = c0 + 2 < p.N ? __ldg(tb + c0 + 2) : INFINITY;  sv.z = c0 + 2 < p.N ? __ldg(p.sgn + c0 + 2) : 1.f;
                        tv.w = c0 + 3 < p.N ? __ldg(tb + c0 + 3) : INFINITY;  sv.w = c0 + 3 < p.N ? __ldg(p.sgn + c0 + 3) : 1.f;
                    }
                    if constexpr (kI8 && !kF4) {
                        // integer form: acc * s >= t  <=>  acc * (+-1) >= ceil(t) for integer acc (|acc| < 2^31 - 1);
                        // cvt.rpi saturates (+inf -> INT_MAX: never, -inf -> INT_MIN: always), NaN -> never
                        auto up = [](float t) { return t != t ? 0x7FFFFFFF : __float2int_ru(t); };
                        auto or1 = [](float x) { return x < 0.f ? -1 : 1; };
                        ptx::sts_v4(xbuf + (uint32_t)(16 * lane), (uint32_t)up(tv.x), (uint32_t)up(tv.y),
                                    (uint32_t)up(tv.z), (uint32_t)up(tv.w));
                        ptx::sts_v4(xbuf + kSignStageBytes / 2 + (uint32_t)(16 * lane), (uint32_t)or1(sv.x),
                                    (uint32_t)or1(sv.y), (uint32_t)or1(sv.z), (uint32_t)or1(sv.w));
                    } else {
                        ptx::sts_v4(xbuf + (uint32_t)(16 * lane), __float_as_uint(tv.x), __float_as_uint(tv.y),
                                    __float_as_uint(tv.z), __float_as_uint(tv.w));
                        ptx::sts_v4(xbuf + kSignStageBytes / 2 + (uint32_t)(16 * lane), __float_as_uint(sv.x),
                                    __float_as_uint(sv.y), __float_as_uint(sv.z), __float_as_uint(sv.w));
                    }
                }
                __syncwarp();
            }
            for (int c = 0; c < n_chunks; ++c) {
                if (STATS == BNN_STATS_BN_BWD && c == n_chunks - 1 && row < p.M) {
                    // pull this row's slice of Z into L2 one chunk ahead of the statistics phase,
                    // whose column-wise reads would otherwise each pay a DRAM round trip
                    const int esz = p.st_zdtype == BNN_Z_S16 ? 2 : 4;
                    const char* zrow = reinterpret_cast<const char*>(p.st_Z) +
                                       ((long long)row * p.st_ldz + colbase) * esz;
                    for (int o = 0; o < CPT * esz; o += 128)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(zrow + o));
                }
                ptx::mbar_wait(&tfull_bar[acc], acc_ph);
                ptx::tc_fence_after_sync();
                const uint32_t t_row =
                    tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + h * CPT);
                if (c == 0) {
                    // first chunk lands directly in the running sum (columns past N hold zeros)
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        uint32_t v[G];
                        ptx::tmem_ld_cols<G>(t_row + (uint32_t)(g * G), v);
                        // FP4: fp32 accumulators holding exact integers (|sum| <= K < 2^24).  The integer stores
                        // convert here; the sign epilogues compare the floats as they are.
                        constexpr bool kCvt = kF4 && EPI != BNN_EPI_SIGN_I8 && EPI != BNN_EPI_SIGN_F4;
                        if constexpr (kCvt) ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < G; ++j)
                            sum[g * G + j] = kCvt ? (uint32_t)__float2int_rn(__uint_as_float(v[j])) : v[j];
                    }
                    ptx::tmem_ld_wait();
                } else {
                    // ptxas would hoist all eight tcgen05.ld's above the adds (8 x 16 temporaries on
                    // top of the 128-register running sum spill); a data-dependent branch after each
                    // group is a barrier for the warp-synchronous loads and keeps 16 temporaries
                    // live at a time.  The tested bit pattern (a NaN payload no accumulator can
                    // produce, far outside any int32 sum) never occurs.
#pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        uint32_t v[G];
                        ptx::tmem_ld_cols<G>(t_row + (uint32_t)(g * G), v);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < G; ++j)
                            sum[g * G + j] =
                                kI8 ? (split_iters
                                           ? __float_as_uint(fmaf((float)(int)v[j], 256.0f,
                                                                  (float)(int)sum[g * G + j]))
                                           : sum[g * G + j] + v[j])
                                    : __float_as_uint(__fadd_rn(__uint_as_float(sum[g * G + j]),
                                                                __uint_as_float(v[j])));
                        if (sum[g * G + G - 1] == 0x7FC12345u) __trap();
                    }
                }
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_leader<kCta2>(&tempty_bar[acc]);
                if (++acc == kAccStages) { acc = 0; acc_ph ^= 1u; }
            }
            constexpr bool kF32Out = EPI == BNN_EPI_STORE_F32 || EPI == BNN_EPI_SGD_F32 ||
                                     EPI == BNN_EPI_SCATTER_F32;
            if (p.debug & 2) continue;   // timing experiment: accumulators drained, nothing stored
            char* dbase = reinterpret_cast<char*>(p.D);
            long long rowoff = (long long)b * p.stride_d + (long long)row * p.ldd;
            if (EPI == BNN_EPI_SCATTER_F32) {
                // reduce-scatter fused into the epilogue: this partial tile is stored straight
                // into slice `rank` of the owner's staging buffer (peer memory over NVLink)
                const int owner = (m_blk * kBlockM) / p.rows_per_owner;
                dbase = reinterpret_cast<char*>(p.peer[owner]);
                rowoff = ((long long)p.rank * p.rows_per_owner +
                          (row - owner * p.rows_per_owner)) * p.ldd;
            }
            // the host only asks for STATS on shapes where every tile is full-width and vector-aligned
            if (kF32Out && ((p.vec_ok && colbase + CPT <= p.N) || STATS != BNN_STATS_NONE)) {  // (digit-split
                // and column-scaled products are admitted by the host only when every tile is full)
                // full-width tile: whole 128-byte row segments per store instruction
                float* dst = reinterpret_cast<float*>(dbase) + (rowoff - (long long)lane * p.ldd) + colbase;
                staged_rows<kI8, EPI, CPT, STATS, true>(p, xbuf, dst, row_w0, colbase, sum, scale, lane);
            } else if ((EPI == BNN_EPI_SIGN_I8 || EPI == BNN_EPI_SIGN_F4 || EPI == BNN_EPI_STORE_S16) && G == 16 &&
                       p.vec_ok && colbase + CPT <= p.N) {
                // full-width tile of an integer epilogue: rows staged, whole row segments per store instruction
                constexpr int kNum = EPI == BNN_EPI_STORE_S16 ? 4 : (EPI == BNN_EPI_SIGN_I8 ? 2 : 1);   // half bytes
                uint8_t* g0 = reinterpret_cast<uint8_t*>(dbase) +
                              (((long long)b * p.stride_d + (long long)row_w0 * p.ldd + colbase) * kNum) / 2;
                if constexpr (G == 16 && (EPI == BNN_EPI_SIGN_I8 || EPI == BNN_EPI_SIGN_F4 || EPI == BNN_EPI_STORE_S16))
                    staged_packed_rows<kI8, kF4, EPI, CPT>(p, xbuf, g0, (p.ldd * kNum) / 2, p.M - row_w0, sum, scale,
                                                           lane);
            } else if (row < p.M) {
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const int col0 = colbase + g * G;
                    if (col0 < p.N) {
                        uint32_t v[G];
#pragma unroll
                        for (int j = 0; j < G; ++j) v[j] = sum[g * G + j];
                        store_group<kI8, kF4, EPI, G>(p, dbase, rowoff + col0, col0, v, scale, xbuf, g * G);
                        asm volatile("" ::: "memory");  // keep the groups' live ranges disjoint
                    }
                }
            }
            if (!kF32Out && STATS != BNN_STATS_NONE) {
                // integer outputs keep their direct stores; the column sums take the staged route
                __syncwarp();
                staged_rows<kI8, EPI, CPT, STATS, false>(p, xbuf, nullptr, row_w0, colbase, sum, scale, lane);
            }
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::cluster_sync_if<kCta2>();  // pair: neither CTA leaves while the other may still reach into it
    if (mc) ptx::cluster_sync();
    if (warp == 2) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc_t<kCta2, Cfg::kTmemCols>(tmem_base);
    }
#undef BNN_TILE_LOOP
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

// BNN_GEMM_DEBUG (read per call): bit mask of timing experiments that break the RESULT on purpose — 1: the producer
// signals its barriers without issuing the TMA loads (tensor-pipe rate without operand traffic), 2: the epilogue
// drains the accumulators without storing.  Used by tools/bench_gemm_forms.py; never set in product runs.
int gemm_debug_flags() {
    const char* e = getenv("BNN_GEMM_DEBUG");
    return e ? atoi(e) : 0;
}

// Tile-walk parameters (tile_coords): super-columns whose B tiles (all pieces the passes read) take at most
// ~40 MB — a third of the 126 MB L2, which holds A's stream and the output beside them — balanced over the width.
// BNN_GEMM_GN / BNN_GEMM_GM (read once) override for experiments.
void pick_tile_walk(const bnn_gemm_desc& g, int es, int block_n, int m_tiles, int n_tiles, int& group_m,
                    int& group_n) {
    static const int env_gn = [] { const char* e = getenv("BNN_GEMM_GN"); return e ? atoi(e) : 0; }();
    static const int env_gm = [] { const char* e = getenv("BNN_GEMM_GM"); return e ? atoi(e) : 0; }();
    bool used[3] = {false, false, false};
    for (int i = 0; i < g.n_pass && i < BNN_GEMM_MAX_PASS; ++i) used[g.pb[i] < 3 ? g.pb[i] : 0] = true;
    const int pieces = (int)used[0] + (int)used[1] + (int)used[2];
    const double tile_bytes = (double)block_n * (double)g.K * es * (pieces > 0 ? pieces : 1);
    int gn = (int)(40.0e6 / tile_bytes);
    gn = gn < 1 ? 1 : (gn > n_tiles ? n_tiles : gn);
    const int cols = cdiv(n_tiles, gn);
    gn = cdiv(n_tiles, cols);                       // balanced super-columns
    if (g.batch > 1 || g.stride_b != 0) gn = n_tiles;  // batched products: every batch has its own B
    group_n = env_gn > 0 ? (env_gn > n_tiles ? n_tiles : env_gn) : gn;
    group_m = env_gm > 0 ? env_gm : 8;
    if (group_m > m_tiles) group_m = m_tiles;
    if (group_m < 1) group_m = 1;
}

// BNN_GEMM_MC (read on every call): the cluster-multicast form of launch_tc — two CTAs on vertically adjacent
// tiles share the B tile through TMA multicast.  1 = on for every eligible product, 0 = off; default: see launch_tc.
int multicast_mode() {
    const char* e = getenv("BNN_GEMM_MC");
    return e ? atoi(e) : -1;
}

// BNN_GEMM_CTA2=1 (read on every call, like BNN_GEMM_CHUNK: the tests flip both inside one process) routes
// 256-wide products with more than one row tile to the CTA-pair form.  Round 2, on a B200: the opt-in tests pass
// (tests/test_gemm_pair_gpu.py) and the training step runs 1.7x SLOWER with it (profiles/r02_summary.md), so it
// stays opt-in.
bool pair_form_requested() {
    const char* e = getenv("BNN_GEMM_CTA2");
    return e != nullptr && atoi(e) != 0;
}

// The CTA-pair launch (gemm_tc_kernel<..., kCta2 = true>): clusters of two CTAs, 256 x 256 tiles per pair.
// Differences from launch_tc: the B maps fetch half tiles (128 rows per CTA), m_tiles counts PAIRS of row
// tiles (an odd last tile leaves the peer CTA with rows past M: TMA zero-fills them and the epilogue's row
// guards drop them), the grid is an even number of CTAs launched with a cluster dimension of 2.
template <bool kI8, int EPI, int STATS>
int launch_tc_pair(const bnn_gemm_desc& g, cudaStream_t stream) {
    constexpr int BLOCK_N = 256;
    using Cfg = TileCfg<BLOCK_N, true>;
    auto kernel = gemm_tc_kernel<kI8, BLOCK_N, EPI, STATS, true>;
    static std::atomic<bool> attr_set[64];  // racing first calls set the same attribute twice: harmless
    int dev = 0;
    BNN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        BNN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    CUtensorMap mA[3], mB[3];
    const int nA = (g.A[1] != nullptr) ? 2 : 1;
    const void* Bp[3] = {g.B[0], g.B[1] ? g.B[1] : g.B[0], g.B2 ? g.B2 : g.B[0]};
    for (int i = 0; i < 3; ++i) {
        int rc = make_operand_map(&mA[i], g.A[i < nA ? i : 0], kI8, g.M, g.K, g.lda,
                                  g.stride_a ? g.batch : 1, g.stride_a, kBlockM);
        if (rc) return rc;
        rc = make_operand_map(&mB[i], Bp[i], kI8, g.N, g.K, g.ldb, g.stride_b ? g.batch : 1, g.stride_b,
                              BLOCK_N / 2);
        if (rc) return rc;
    }
    TcParams p;
    p.M = g.M;
    p.N = g.N;
    p.batch = g.batch;
    p.kblocks = cdiv(g.K, kI8 ? kKBytes : kKBytes / 2);
    p.n_pass = g.n_pass;
    p.chunk_iters = kI8 ? (p.kblocks * g.n_pass) : gemm_chunk_iters();
    for (int i = 0; i < BNN_GEMM_MAX_PASS; ++i) {
        p.pa[i] = g.pa[i];
        p.pb[i] = g.pb[i];
    }
    p.m_tiles = cdiv(cdiv(g.M, kBlockM), 2);  // pairs of row tiles
    p.n_tiles = cdiv(g.N, BLOCK_N);
    pick_tile_walk(g, kI8 ? 1 : 2, BLOCK_N, p.m_tiles, p.n_tiles, p.group_m, p.group_n);
    p.mc = 0;
    p.debug = gemm_debug_flags();
    p.epilogue = g.epilogue;
    const int esz = (g.epilogue == BNN_EPI_STORE_S16) ? 2 : (g.epilogue == BNN_EPI_SIGN_I8 ? 1 : 4);
    p.vec_ok = aligned16(g.D) && ((g.ldd * esz) % 16 == 0) && ((g.stride_d * esz) % 16 == 0);
    p.a_batched = g.stride_a != 0;
    p.b_batched = g.stride_b != 0;
    p.D = g.D;
    p.ldd = g.ldd;
    p.stride_d = g.stride_d;
    p.sa = g.sa;
    p.sb = g.sb;
    p.host_scale = g.host_scale;
    p.thr = g.epi_thr;
    p.sgn = g.epi_sgn;
    p.thr_stride = g.epi_thr_stride;
    p.a_unsigned = g.a_unsigned;
    p.rows_per_owner = 0;
    p.rank = 0;
    p.m_rot = 0;
    for (int r = 0; r < BNN_MAX_RANKS; ++r) p.peer[r] = 0;
    p.hi_from_pass = g.hi_from_pass;
    p.col_scale = g.col_scale;
    const bool f32out = EPI == BNN_EPI_STORE_F32 || EPI == BNN_EPI_SGD_F32;
    if (g.hi_from_pass != 0 || g.col_scale != nullptr) {
        if (!(f32out && p.vec_ok && g.N % BLOCK_N == 0 && g.batch == 1 &&
              (g.hi_from_pass == 0 || (kI8 && g.hi_from_pass > 0 && g.hi_from_pass < g.n_pass)) &&
              (!g.col_scale || aligned16(g.col_scale)))) {
            set_error("bnn_gemm: digit-split / column-scaled products need int8 operands (split), an fp32 "
                      "epilogue, N %% %d == 0, 16-byte aligned output and col_scale, batch == 1 (N=%d)",
                      BLOCK_N, g.N);
            return BNN_ERR_UNSUPPORTED;
        }
    }
    p.st_sums = g.stats_sums;
    p.st_max = g.stats_max;
    p.st_Z = g.stats_z;
    p.st_ldz = g.stats_ldz;
    p.st_zdtype = g.stats_z_dtype;
    p.st_mu = g.stats_mu;
    if (STATS != BNN_STATS_NONE && !(p.vec_ok && g.N % BLOCK_N == 0 && g.batch == 1)) {
        set_error("bnn_gemm: fused column statistics need N %% %d == 0, a 16-byte aligned output "
                  "and batch == 1 (N=%d)", BLOCK_N, g.N);
        return BNN_ERR_UNSUPPORTED;
    }
    const long long pairs = (long long)p.batch * p.m_tiles * p.n_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)sm_count() & ~1u, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // The tile walk is static (cluster c takes tiles c, c + clusters, ...): a cluster that cannot be
    // co-resident would start only after another has finished ALL its tiles and double the run time
    // (measured: 74 clusters requested on 148 SMs ran 1.9x slower than the single-CTA form).  So the grid is
    // what the device can hold at once — GPCs do not all expose an even number of usable SMs to clusters.
    static std::atomic<int> max_clusters_dev[64];
    int max_clusters = (dev >= 0 && dev < 64) ? max_clusters_dev[dev].load() : 0;
    if (max_clusters <= 0) {
        BNN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
        if (max_clusters <= 0) {
            set_error("bnn_gemm: the device cannot hold a single CTA pair of this kernel");
            return BNN_ERR_UNSUPPORTED;
        }
        if (dev >= 0 && dev < 64) max_clusters_dev[dev] = max_clusters;
    }
    if (getenv("BNN_GEMM_CTA2_DEBUG")) fprintf(stderr, "bnn_gemm pair: %d co-resident clusters\n", max_clusters);
    const unsigned grid = 2u * (unsigned)(pairs < max_clusters ? pairs : max_clusters);
    cfg.gridDim = dim3(grid, 1, 1);
    BNN_CUDA(cudaLaunchKernelEx(&cfg, kernel, mA[0], mA[1], mB[0], mB[1], mB[2], p));
    BNN_LAUNCH_CHECK("gemm_tc_kernel<pair>");
    return BNN_OK;
}

template <bool kI8, int BLOCK_N, int EPI, int STATS = BNN_STATS_NONE, bool kF4 = false>
int launch_tc(const bnn_gemm_desc& g, cudaStream_t stream) {
    using Cfg = TileCfg<BLOCK_N, false, kF4>;
    auto kernel = gemm_tc_kernel<kI8, BLOCK_N, EPI, STATS, false, kF4>;
    const int pack = kF4 ? 2 : 1;   // operand elements per byte of the tensor map's element type
    if constexpr (BLOCK_N == 256 && EPI != BNN_EPI_SCATTER_F32 && !kF4) {
        if (g.M > kBlockM && pair_form_requested()) return launch_tc_pair<kI8, EPI, STATS>(g, stream);
    }
    // the opt-in shared-memory size is a per-device attribute of the function
    static std::atomic<bool> attr_set[64];  // racing first calls set the same attribute twice: harmless
    int dev = 0;
    BNN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        BNN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    // cluster-multicast form: wide tiles, more than one row tile, not the NVLink scatter epilogue (its row-tile
    // rotation assumes single tiles).  Opt-in (BNN_GEMM_MC=1) until it has earned the default.
    const int mc_env = multicast_mode();
    const bool mc = BLOCK_N >= 128 && EPI != BNN_EPI_SCATTER_F32 && g.M > kBlockM && mc_env == 1;
    CUtensorMap mA[3], mB[3];
    const int nA = (g.A[1] != nullptr) ? 2 : 1;
    const void* Bp[3] = {g.B[0], g.B[1] ? g.B[1] : g.B[0], g.B2 ? g.B2 : g.B[0]};
    for (int i = 0; i < 3; ++i) {
        const void* a = g.A[i < nA ? i : 0];
        const void* b = Bp[i];
        int rc = make_operand_map(&mA[i], a, kI8, g.M, g.K / pack, g.lda / pack, g.stride_a ? g.batch : 1,
                                  g.stride_a / pack, kBlockM);
        if (rc) return rc;
        rc = make_operand_map(&mB[i], b, kI8, g.N, g.K / pack, g.ldb / pack, g.stride_b ? g.batch : 1,
                              g.stride_b / pack, mc ? BLOCK_N / 2 : BLOCK_N);   // multicast: half tiles
        if (rc) return rc;
    }
    TcParams p;
    p.M = g.M;
    p.N = g.N;
    p.batch = g.batch;
    const int kelems = kF4 ? 2 * kKBytes : (kI8 ? kKBytes : kKBytes / 2);
    p.kblocks = cdiv(g.K, kelems);
    p.n_pass = g.n_pass;
    // integer accumulation is exact: one chunk; fp32 accumulation is re-rounded every chunk
    p.chunk_iters = kI8 ? (p.kblocks * g.n_pass) : gemm_chunk_iters();
    for (int i = 0; i < BNN_GEMM_MAX_PASS; ++i) {
        p.pa[i] = g.pa[i];
        p.pb[i] = g.pb[i];
    }
    p.m_tiles = mc ? cdiv(cdiv(g.M, kBlockM), 2) : cdiv(g.M, kBlockM);   // multicast: pairs of row tiles
    p.n_tiles = cdiv(g.N, BLOCK_N);
    p.mc = mc ? 1 : 0;
    pick_tile_walk(g, kI8 ? 1 : 2, BLOCK_N / pack, p.m_tiles, p.n_tiles, p.group_m, p.group_n);
    if (mc && p.group_m > 4) p.group_m = 4;   // 4 pairs = the 8 row tiles of the single-CTA walk
    p.debug = gemm_debug_flags();
    p.epilogue = g.epilogue;
    const int esz = (g.epilogue == BNN_EPI_STORE_S16) ? 2 : (g.epilogue == BNN_EPI_SIGN_I8 ? 1 : 4);
    p.vec_ok = aligned16(g.D) && ((g.ldd * esz) % 16 == 0) && ((g.stride_d * esz) % 16 == 0);
    if (EPI == BNN_EPI_SIGN_F4) {
        // nibble output: whole 32-bit words of 8 columns are stored, so the pitch must cover the last word
        if (!(aligned16(g.D) && g.ldd % 32 == 0 && g.stride_d % 32 == 0 && g.ldd >= ((g.N + 7) / 8) * 8)) {
            set_error("bnn_gemm: BNN_EPI_SIGN_F4 needs a 16-byte aligned D and ldd, stride_d multiples of 32 "
                      "elements (ldd=%lld)", (long long)g.ldd);
            return BNN_ERR_ARG;
        }
        p.vec_ok = 1;
    }
    p.a_batched = g.stride_a != 0;
    p.b_batched = g.stride_b != 0;
    p.D = g.D;
    p.ldd = g.ldd;
    p.stride_d = g.stride_d;
    p.sa = g.sa;
    p.sb = g.sb;
    p.host_scale = g.host_scale;
    p.thr = g.epi_thr;
    p.sgn = g.epi_sgn;
    p.thr_stride = g.epi_thr_stride;
    p.a_unsigned = g.a_unsigned;
    p.rows_per_owner = 0;
    p.rank = 0;
    p.m_rot = 0;
    for (int r = 0; r < BNN_MAX_RANKS; ++r) p.peer[r] = 0;
    if (EPI == BNN_EPI_SCATTER_F32) {
        const bnn_peer_scatter& sc = *g.scatter;
        p.rows_per_owner = sc.rows_per_owner;
        p.rank = sc.rank;
        p.m_rot = (int)(((long long)sc.rank * (sc.rows_per_owner / kBlockM)) % p.m_tiles);
        p.vec_ok = (g.ldd * 4) % 16 == 0;
        for (int r = 0; r < sc.world; ++r) {
            p.peer[r] = sc.stage[r];
            p.vec_ok = p.vec_ok && (sc.stage[r] % 16 == 0);
        }
        p.D = nullptr;
        p.stride_d = 0;
    }
    const long long tiles = (long long)p.batch * p.m_tiles * p.n_tiles;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    p.hi_from_pass = g.hi_from_pass;
    p.col_scale = g.col_scale;
    if (g.hi_from_pass != 0 || g.col_scale != nullptr) {
        // handled by the staged full-width epilogue only
        const bool f32out = EPI == BNN_EPI_STORE_F32 || EPI == BNN_EPI_SGD_F32 || EPI == BNN_EPI_SCATTER_F32;
        if (!(f32out && p.vec_ok && g.N % BLOCK_N == 0 && g.batch == 1 &&
              (g.hi_from_pass == 0 || (kI8 && g.hi_from_pass > 0 && g.hi_from_pass < g.n_pass)) &&
              (!g.col_scale || aligned16(g.col_scale)))) {
            set_error("bnn_gemm: digit-split / column-scaled products need int8 operands (split), an fp32 "
                      "epilogue, N %% %d == 0, 16-byte aligned output and col_scale, batch == 1 (N=%d)",
                      BLOCK_N, g.N);
            return BNN_ERR_UNSUPPORTED;
        }
    }
    p.st_sums = g.stats_sums;
    p.st_max = g.stats_max;
    p.st_Z = g.stats_z;
    p.st_ldz = g.stats_ldz;
    p.st_zdtype = g.stats_z_dtype;
    p.st_mu = g.stats_mu;
    if (STATS != BNN_STATS_NONE) {
        // every tile must take the staged full-width path.  (The FP4 form only has integer outputs: their column
        // sums are a separate staged pass with per-column guards, so ragged widths are fine there.)
        if (!((kF4 || (p.vec_ok && g.N % BLOCK_N == 0)) && g.batch == 1)) {
            set_error("bnn_gemm: fused column statistics need N %% %d == 0, a 16-byte aligned output "
                      "and batch == 1 (N=%d)", BLOCK_N, g.N);
            return BNN_ERR_UNSUPPORTED;
        }
    }
    if (mc) {
        // clusters of two CTAs; the tile walk is static, so the grid is what the device holds at once (§3.5)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)sm_count() & ~1u, 1, 1);
        cfg.blockDim = dim3(kThreads, 1, 1);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        static std::atomic<int> max_clusters_dev[64];
        int max_clusters = (dev >= 0 && dev < 64) ? max_clusters_dev[dev].load() : 0;
        if (max_clusters <= 0) {
            BNN_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg));
            if (max_clusters <= 0) {
                set_error("bnn_gemm: the device cannot hold a single cluster of this kernel");
                return BNN_ERR_UNSUPPORTED;
            }
            if (dev >= 0 && dev < 64) max_clusters_dev[dev] = max_clusters;
        }
        const long long pairs = (long long)p.batch * p.m_tiles * p.n_tiles;
        cfg.gridDim = dim3(2u * (unsigned)(pairs < max_clusters ? pairs : max_clusters), 1, 1);
        BNN_CUDA(cudaLaunchKernelEx(&cfg, kernel, mA[0], mA[1], mB[0], mB[1], mB[2], p));
        BNN_LAUNCH_CHECK("gemm_tc_kernel<multicast>");
        return BNN_OK;
    }
    kernel<<<grid, kThreads, Cfg::kSmemBytes, stream>>>(mA[0], mA[1], mB[0], mB[1], mB[2], p);
    BNN_LAUNCH_CHECK("gemm_tc_kernel");
    return BNN_OK;
}

template <int BLOCK_N>
int pick_epilogue(const bnn_gemm_desc& g, bool i8, cudaStream_t stream) {
    if (g.stats == BNN_STATS_COLSUM) {
        if (i8 && g.epilogue == BNN_EPI_STORE_S16)
            return launch_tc<true, BLOCK_N, BNN_EPI_STORE_S16, BNN_STATS_COLSUM>(g, stream);
        if (i8 && g.epilogue == BNN_EPI_STORE_S32)
            return launch_tc<true, BLOCK_N, BNN_EPI_STORE_S32, BNN_STATS_COLSUM>(g, stream);
        if (!i8 && g.epilogue == BNN_EPI_STORE_F32)
            return launch_tc<false, BLOCK_N, BNN_EPI_STORE_F32, BNN_STATS_COLSUM>(g, stream);
        set_error("bnn_gemm: BNN_STATS_COLSUM goes with int8 S16/S32 or fp16 F32 stores");
        return BNN_ERR_UNSUPPORTED;
    }
    if (g.stats == BNN_STATS_BN_BWD) {
        if (!i8 && g.epilogue == BNN_EPI_STORE_F32)
            return launch_tc<false, BLOCK_N, BNN_EPI_STORE_F32, BNN_STATS_BN_BWD>(g, stream);
        set_error("bnn_gemm: BNN_STATS_BN_BWD goes with fp16 operands and BNN_EPI_STORE_F32");
        return BNN_ERR_UNSUPPORTED;
    }
    if (i8) {
        switch (g.epilogue) {
            case BNN_EPI_STORE_S16: return launch_tc<true, BLOCK_N, BNN_EPI_STORE_S16>(g, stream);
            case BNN_EPI_STORE_S32: return launch_tc<true, BLOCK_N, BNN_EPI_STORE_S32>(g, stream);
            case BNN_EPI_STORE_F32: return launch_tc<true, BLOCK_N, BNN_EPI_STORE_F32>(g, stream);
            case BNN_EPI_SIGN_I8: return launch_tc<true, BLOCK_N, BNN_EPI_SIGN_I8>(g, stream);
            case BNN_EPI_SIGN_F4: return launch_tc<true, BLOCK_N, BNN_EPI_SIGN_F4>(g, stream);
            case BNN_EPI_SCATTER_F32: return launch_tc<true, BLOCK_N, BNN_EPI_SCATTER_F32>(g, stream);
            default: return launch_tc<true, BLOCK_N, BNN_EPI_SGD_F32>(g, stream);
        }
    }
    switch (g.epilogue) {
        case BNN_EPI_SGD_F32: return launch_tc<false, BLOCK_N, BNN_EPI_SGD_F32>(g, stream);
        case BNN_EPI_SIGN_I8: return launch_tc<false, BLOCK_N, BNN_EPI_SIGN_I8>(g, stream);
        case BNN_EPI_SIGN_F4: return launch_tc<false, BLOCK_N, BNN_EPI_SIGN_F4>(g, stream);
        case BNN_EPI_SCATTER_F32: return launch_tc<false, BLOCK_N, BNN_EPI_SCATTER_F32>(g, stream);
        default: return launch_tc<false, BLOCK_N, BNN_EPI_STORE_F32>(g, stream);
    }
}

// E2M1 operands (BNN_DT_F4): integer epilogues only, tiles of 32 / 128 / 224 columns.  The training forward asks
// for the fused column sums (BNN_STATS_COLSUM, int16 / int32 stores) on any width.
template <int BLOCK_N>
int pick_epilogue_f4(const bnn_gemm_desc& g, cudaStream_t stream) {
    if (g.stats == BNN_STATS_COLSUM) {
        if (g.epilogue == BNN_EPI_STORE_S16)
            return launch_tc<true, BLOCK_N, BNN_EPI_STORE_S16, BNN_STATS_COLSUM, true>(g, stream);
        if (g.epilogue == BNN_EPI_STORE_S32)
            return launch_tc<true, BLOCK_N, BNN_EPI_STORE_S32, BNN_STATS_COLSUM, true>(g, stream);
        set_error("bnn_gemm: BNN_STATS_COLSUM on E2M1 operands goes with the int16 / int32 stores");
        return BNN_ERR_UNSUPPORTED;
    }
    switch (g.epilogue) {
        case BNN_EPI_STORE_S16: return launch_tc<true, BLOCK_N, BNN_EPI_STORE_S16, BNN_STATS_NONE, true>(g, stream);
        case BNN_EPI_STORE_S32: return launch_tc<true, BLOCK_N, BNN_EPI_STORE_S32, BNN_STATS_NONE, true>(g, stream);
        case BNN_EPI_SIGN_I8: return launch_tc<true, BLOCK_N, BNN_EPI_SIGN_I8, BNN_STATS_NONE, true>(g, stream);
        case BNN_EPI_SIGN_F4: return launch_tc<true, BLOCK_N, BNN_EPI_SIGN_F4, BNN_STATS_NONE, true>(g, stream);
        default:
            set_error("bnn_gemm: BNN_DT_F4 products have the integer epilogues STORE_S16 / STORE_S32 / SIGN_I8 / SIGN_F4");
            return BNN_ERR_UNSUPPORTED;
    }
}

}  // namespace

int gemm_tc_dispatch(const bnn_gemm_desc& g, cudaStream_t stream) {
    if (g.in_dtype == BNN_DT_F4) {
        BNN_CHECK_ARG(g.A[0] && g.B[0] && aligned16(g.A[0]) && aligned16(g.B[0]), "bnn_gemm: F4 operands not 16-byte aligned");
        BNN_CHECK_ARG(g.n_pass == 1 && g.stats != BNN_STATS_BN_BWD && g.hi_from_pass == 0 && !g.col_scale &&
                          !g.a_unsigned,
                      "bnn_gemm: BNN_DT_F4 is a single-pass +-1 product (fused statistics: column sums only)");
        BNN_CHECK_ARG(g.K % 32 == 0 && g.lda % 32 == 0 && g.ldb % 32 == 0 && g.stride_a % 32 == 0 &&
                          g.stride_b % 32 == 0 && g.lda >= g.K && g.ldb >= g.K,
                      "bnn_gemm: F4 operands need K, pitches and batch strides in multiples of 32 elements (16 bytes)");
        BNN_CHECK_ARG(g.K < (1 << 24), "bnn_gemm: F4 sums are exact in fp32 only for K < 2^24");
        return g.N <= 32 ? pick_epilogue_f4<32>(g, stream)
                         : (g.N <= 128 ? pick_epilogue_f4<128>(g, stream) : pick_epilogue_f4<224>(g, stream));
    }
    const bool i8 = g.in_dtype == BNN_DT_I8;
    const int es = i8 ? 1 : 2;
    for (int i = 0; i < 2; ++i) {
        if (g.A[i]) BNN_CHECK_ARG(aligned16(g.A[i]), "bnn_gemm: A[%d] not 16-byte aligned", i);
        if (g.B[i]) BNN_CHECK_ARG(aligned16(g.B[i]), "bnn_gemm: B[%d] not 16-byte aligned", i);
    }
    if (g.B2) BNN_CHECK_ARG(aligned16(g.B2), "bnn_gemm: B2 not 16-byte aligned");
    BNN_CHECK_ARG((g.lda * es) % 16 == 0 && (g.ldb * es) % 16 == 0,
                  "bnn_gemm: operand pitches must be multiples of 16 bytes (lda=%lld ldb=%lld)",
                  (long long)g.lda, (long long)g.ldb);
    BNN_CHECK_ARG((g.stride_a * es) % 16 == 0 && (g.stride_b * es) % 16 == 0,
                  "bnn_gemm: batch strides must be multiples of 16 bytes");
    BNN_CHECK_ARG(g.lda >= g.K && g.ldb >= g.K, "bnn_gemm: pitch smaller than K");
    return g.N <= 32 ? pick_epilogue<32>(g, i8, stream)
                     : (g.N <= 128 ? pick_epilogue<128>(g, i8, stream) : pick_epilogue<256>(g, i8, stream));
}

}  // namespace bnn
