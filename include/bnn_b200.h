/*
 * libbnn_b200 — C ABI of the B200-native (sm_100a) binarized-dense-layer hot path.
 *
 * This is the drop-in boundary for ONE path of volf52/deep-neural-net: the binarized dense layer
 * (mlpcode/layers.py) and the bit-flip error-injection callbacks (mlpcode/callbacks.py).  The
 * reference has no FFI of its own — its seam is the Python class surface of mlpcode.layers selected
 * by `Layer.xp` (layers.py:10-15) — so every entry point below names the reference code it
 * replaces (file:line relative to the reference root).  The host side that calls these through
 * ctypes lives in deep-neural-net_b200/mlpcode/ and mirrors the reference classes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the library never allocates or
 *     frees caller-visible memory (small internal scratch excepted) and never takes ownership;
 *   - matrices are row-major; `ld*` are row pitches in ELEMENTS; operands read by TMA need a
 *     16-byte aligned base and a pitch that is a multiple of 16 bytes;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant per stream;
 *   - return value: 0 on success, negative BNN_ERR_* otherwise; bnn_last_error() gives the text.
 *     No exceptions cross the ABI.  There is no CPU fallback: without a usable sm_100 device every
 *     compute entry point fails with BNN_ERR_CUDA.
 *   - sign convention (layers.py:278-279, activation.py:110-111, callbacks.py:68-72):
 *     +1 <-> bit 1 <-> (x >= 0), -1 <-> bit 0; sign(-0.0) = +1; NaN packs as bit 0.
 */
#ifndef BNN_B200_H
#define BNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNN_B200_VERSION 100

typedef void* bnn_stream_t; /* cudaStream_t */

enum {
    BNN_OK = 0,
    BNN_ERR_ARG = -1,         /* bad shape / pitch / alignment / enum */
    BNN_ERR_CUDA = -2,        /* CUDA runtime or driver error (text in bnn_last_error) */
    BNN_ERR_UNSUPPORTED = -3  /* valid request this build does not implement */
};

/* ---------------------------------------------------------------------------------------------
 * Library / device
 * ------------------------------------------------------------------------------------------- */
const char* bnn_last_error(void);
int bnn_version(void);
/* Fills SM count and compute capability of the current device. */
int bnn_device_query(int* sm_count_host, int* cc_major_host, int* cc_minor_host);

/* ---------------------------------------------------------------------------------------------
 * A1  sign() binarization and bit packing       (BinaryLayer.binarize, layers.py:275-281;
 *                                                 unitstep, activation.py:106-113)
 * ------------------------------------------------------------------------------------------- */

/* bits[r, w] bit j = (x[r, 32w + j] >= 0); pad bits of the last word are 0.
 * x: [R, C] fp32 pitch ldx; bits: [R, ldw] uint32, ldw >= ceil(C/32). */
int bnn_pack_sign_rows(const float* x, int R, int C, int64_t ldx, uint32_t* bits, int64_t ldw,
                       bnn_stream_t stream);

/* Same packing for an int8 +-1 matrix (bit = value > 0). */
int bnn_pack_i8_rows(const int8_t* x, int R, int C, int64_t ldx, uint32_t* bits, int64_t ldw,
                     bnn_stream_t stream);

/* Transposing binarization of a latent weight matrix W[K, N] (fp32, pitch N) into the K-major
 * operand layouts the forward GEMMs read.  Any output may be NULL.
 *   wt_i8  [N, ld_i8 ]  int8  +-1                 (tcgen05 kind::i8 operand)
 *   wt_f16 [N, ld_f16]  fp16  +-1                 (tcgen05 kind::f16 operand, real-input layer 0)
 *   wt_bits[N, ld_bits] uint32, bit j of word w = (W[32w+j, n] >= 0), pad bits 0 (XOR+popc operand,
 *                       and the "packed weights" the fault kernels mutate)
 *   wt_f4  [N, ld_f4 / 2 bytes] E2M1 nibbles +-1.0 (tcgen05 kind::mxf4 operand, BNN_DT_F4; layout of
 *                       bnn_bits_to_f4: ld_f4 counts elements, a multiple of 32, nibbles past K are zero)
 *   absmax_bits         device uint32 <- bits(max|W|): the bound the backward fp16 split of W needs,
 *                       obtained from the same read of W (see bnn_absmax_f32) */
int bnn_binarize_weights_t(const float* W, int K, int N, int8_t* wt_i8, int64_t ld_i8,
                           void* wt_f16, int64_t ld_f16, uint32_t* wt_bits, int64_t ld_bits,
                           void* wt_f4, int64_t ld_f4, uint32_t* absmax_bits, bnn_stream_t stream);

/* Expand packed sign bits [R, ldw] back to int8 +-1 [R, C] (pitch ldo) and/or fp16 +-1. */
int bnn_unpack_bits(const uint32_t* bits, int R, int C, int64_t ldw, int8_t* out_i8, int64_t ld_i8,
                    void* out_f16, int64_t ld_f16, bnn_stream_t stream);

/* Mini-batch assembly of the epoch loop (trainX[p, :] then X[k:k+bs], network.py:312-315, 480) without a shuffled
 * copy of the data set: dst[i, :] = src[idx[i], :] for i < n_rows, on raw bytes (row_bytes per row; pitches in
 * bytes).  idx: device int64 — a slice of the composed permutation.  Rows with idx outside [0, n_src_rows) are
 * zero-filled and counted in *bad_count (device int32, may be NULL; not cleared here). */
int bnn_gather_rows(const void* src, int64_t src_pitch_bytes, int64_t n_src_rows, const int64_t* idx, int n_rows,
                    int64_t row_bytes, void* dst, int64_t dst_pitch_bytes, int32_t* bad_count,
                    bnn_stream_t stream);

/* Expand packed sign bits [R, ldw] to E2M1 nibbles [R, ld_f4 / 2 bytes] — the BNN_DT_F4 operand form of a +-1
 * matrix: bit j of word w -> nibble (32w + j): 1 -> 0x2 (+1.0), 0 -> 0xA (-1.0); nibbles in [C, 32*ceil(C/32))
 * -> 0x0 (+0.0, the K padding of the product).  ld_f4 counts ELEMENTS (a multiple of 32).  This is how the
 * fault sweeps reach the FP4 tensor pipe: the flip kernels XOR their masks into the packed words
 * (callbacks.py:79-83 on 1 bit per weight) and this pass expands the result. */
int bnn_bits_to_f4(const uint32_t* bits, int64_t R, int C, int64_t ldw, void* out_f4, int64_t ld_f4,
                   bnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * fp32 -> 2 x fp16 split operands for the real-valued GEMMs (layer-0 forward, dW, dX).
 * x*s = hi + lo (+ O(2^-22)), s = 2^(14 - ceil(log2(bound))) a power of two, so three fp16
 * tensor-core products hi*hi + hi*lo + lo*hi with fp32 accumulation reproduce an fp32 SGEMM
 * (X.dot(W), layers.py:220,260,266) to better than SGEMM's own rounding error.
 * ------------------------------------------------------------------------------------------- */

/* absmax_bits <- max(absmax_bits, bits(max|x|)) (uint32 view of a non-negative float; the caller
 * zeroes it, or passes zero_first=1). */
int bnn_absmax_f32(const float* x, int64_t n, uint32_t* absmax_bits, int zero_first,
                   bnn_stream_t stream);

/* Split x[R, C] (pitch ldx).  bound: device float >= max|x| (e.g. the absmax above).
 * Row-major pieces hi/lo [R, ld_rm] and/or transposed pieces hi_t/lo_t [C, ld_t] (fp16), any pair may
 * be NULL.  scale_out (device float, may be NULL) receives s.  Pad columns between the logical
 * width and the pitch are never written and never read by the GEMMs. */
int bnn_split_f16(const float* x, int R, int C, int64_t ldx, const float* bound, void* hi, void* lo,
                  int64_t ld_rm, void* hi_t, void* lo_t, int64_t ld_t, float* scale_out,
                  bnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * A2/A9  GEMMs                                    (X.dot(weight) layers.py:220;
 *                                                  input.T.dot(delta) :260; delta.dot(W.T) :266;
 *                                                  weights -= lr*dw :268)
 * D[b][M, N] = epilogue( sum_p  A[pa[p]][b][M, K] * B[pb[p]][b][N, K]^T )
 * Both operands K-major (row-major with the contraction index contiguous).
 * ------------------------------------------------------------------------------------------- */
enum {
    BNN_DT_I8 = 0,
    BNN_DT_F16 = 1,
    BNN_DT_F4 = 2 /* E2M1 nibbles, two per byte, low nibble = even k (+1 = 0x2, -1 = 0xA, 0 = 0x0): the +-1 x +-1
                     products of the evaluation paths on tcgen05 kind::mxf4.block_scale with unit scale factors —
                     exact (fp32 accumulators hold integers, K < 2^24), twice the kind::i8 rate, half the operand
                     bytes.  K, lda, ldb, stride_a, stride_b count ELEMENTS and are multiples of 32; one pass,
                     integer epilogues (STORE_S32 / STORE_S16 / SIGN_I8 / SIGN_F4); of the fused statistics only
                     BNN_STATS_COLSUM with the integer stores (any N: the sums are guarded per column). */
};
enum {
    BNN_EPI_STORE_F32 = 0, /* D fp32  = acc * host_scale / (sa*sb)                       */
    BNN_EPI_STORE_S32 = 1, /* D int32 = acc                      (I8 only)               */
    BNN_EPI_STORE_S16 = 2, /* D int16 = acc                      (I8 only, |acc| < 2^15) */
    BNN_EPI_SGD_F32 = 3,   /* D fp32 -= acc * host_scale / (sa*sb)   (fused SGD update)  */
    BNN_EPI_SIGN_I8 = 4,   /* D int8 = +1 if acc*host_scale/(sa*sb) * epi_sgn[n] >= epi_thr[n] else -1 (epi_sgn holds
                              +-1; integer-accumulating kinds use its sign and decide acc * (+-1) >= ceil(epi_thr),
                              the same decision for every integer acc):
                              batch-norm + sign fused into the GEMM (thresholds from
                              bnn_bn_sign_thresholds); Z never reaches HBM (eval mode)           */
    BNN_EPI_SIGN_F4 = 6,   /* like SIGN_I8, stored as E2M1 nibbles (the A operand of a BNN_DT_F4 product): D uint8
                              [M, ldd / 2] with ldd, stride_d in ELEMENTS, multiples of 32; columns in [N, 16*ceil(N/16))
                              are written as 0 (they are the consumer's K padding)                     */
    BNN_EPI_SCATTER_F32 = 5 /* like STORE_F32, but every 128-row tile is stored over NVLink into the
                              staging buffer of the rank that owns those rows (desc.scatter): the
                              reduce-scatter half of the data-parallel dW all-reduce, fused into the
                              GEMM epilogue.  D / stride_d are ignored, ldd is the staging pitch.   */
};
#define BNN_MAX_RANKS 16

/* Column reductions fused into the GEMM epilogue (tcgen05 path; N must be a multiple of the tile
 * width — 256 for N > 128, 128 for N > 32, else 32 — the output 16-byte aligned and batch == 1,
 * otherwise BNN_ERR_UNSUPPORTED and the caller uses the stand-alone kernels):
 *   BNN_STATS_COLSUM  stats_sums[2,N] += (sum_m D, sum_m D^2)        == bnn_colstats on the result
 *   BNN_STATS_BN_BWD  stats_sums[2,N] += (sum_m D, sum_m D*(Z-mu)),  stats_max[2,N] = max(.., |D|, |Z-mu|)
 *                                                                    == bnn_bn_bwd_reduce(delta = D)
 * The accumulators are NOT cleared by the GEMM. */
enum { BNN_STATS_NONE = 0, BNN_STATS_COLSUM = 1, BNN_STATS_BN_BWD = 2 };

/* Row ownership and peer mappings for BNN_EPI_SCATTER_F32.  Rank r owns output rows
 * [r*rows_per_owner, (r+1)*rows_per_owner) (rows_per_owner a multiple of 128).  Every rank holds a
 * staging buffer float[world][rows_per_owner][ldd]; stage[o] is the address of rank o's buffer as
 * mapped into THIS process (symmetric memory).  Rank `rank` stores row m of its partial product at
 * stage[o][rank][m - o*rows_per_owner][.] with o = m / rows_per_owner, and visits the row tiles
 * starting from its own, so that at any moment the ranks write to different owners. */
typedef struct bnn_peer_scatter {
    int32_t world, rank;
    int32_t rows_per_owner;
    int32_t reserved;
    uint64_t stage[BNN_MAX_RANKS];
} bnn_peer_scatter;
enum {
    BNN_GEMM_TCGEN05 = 0, /* TMA + tcgen05.mma + TMEM accumulators (product path)          */
    BNN_GEMM_SIMT = 1     /* plain fp32/int32 CUDA-core kernel: cross-check + tiny shapes */
};
#define BNN_GEMM_MAX_PASS 4

typedef struct bnn_gemm_desc {
    int32_t M, N, K;
    int32_t batch;    /* >= 1 */
    int32_t in_dtype; /* BNN_DT_* */
    int32_t epilogue; /* BNN_EPI_* */
    int32_t n_pass;   /* 1..BNN_GEMM_MAX_PASS */
    int32_t pa[BNN_GEMM_MAX_PASS];
    int32_t pb[BNN_GEMM_MAX_PASS];
    const void* A[2]; /* operand pieces, each [batch][M, lda] */
    const void* B[2]; /* operand pieces, each [batch][N, ldb] */
    int64_t lda, ldb; /* row pitch, elements */
    int64_t stride_a, stride_b; /* batch stride in elements; 0 = operand shared by all batches */
    void* D;
    int64_t ldd, stride_d;
    const float* sa; /* device scalars (may be NULL = 1): the split scales of A and B */
    const float* sb;
    float host_scale; /* 1 for plain stores, lr for BNN_EPI_SGD_F32 */
    int32_t a_unsigned; /* BNN_DT_I8 only: 1 = the A pieces hold uint8 (0..255) instead of int8 — the
                           exact u8 x s8 product of a uint8 image batch with +-1 weights (layer 0 fed
                           from ImageErrorCallback output, imageBitflipTrainModelSel.py:26-28) */
    const float* epi_thr; /* BNN_EPI_SIGN_I8: per-column threshold and orientation (device, [N]);  */
    const float* epi_sgn; /* batch b reads epi_thr + b*epi_thr_stride (epi_sgn is shared)          */
    const bnn_peer_scatter* scatter; /* BNN_EPI_SCATTER_F32 (host pointer), else NULL */
    int32_t stats;         /* BNN_STATS_* */
    int32_t stats_z_dtype; /* BNN_Z_* of stats_z (BNN_STATS_BN_BWD) */
    double* stats_sums;    /* device [2, N] */
    uint32_t* stats_max;   /* device [2, N] (BNN_STATS_BN_BWD) */
    const void* stats_z;   /* device [M, stats_ldz] (BNN_STATS_BN_BWD) */
    int64_t stats_ldz;
    const float* stats_mu; /* device [N] (BNN_STATS_BN_BWD) */
    /* Exact fixed-point products on the int8 tensor pipe (dW = X^T . delta' with X = +-1 and delta'
     * cut into signed digits, layers.py:260).  B2 is a third B piece (pb[p] == 2).  With
     * hi_from_pass = k > 0 (int8 operands only) the passes [0, k) accumulate a LOW int32 accumulator
     * and the passes [k, n_pass) a HIGH one; the epilogue value is low + 256 * high, formed in fp32.
     * col_scale: optional device [N] per-column factor (1 / the fixed-point scale of that column)
     * multiplied in before host_scale.  tcgen05 path, fp32 epilogues, N a multiple of the tile
     * width (see BNN_STATS_*), batch == 1; otherwise BNN_ERR_UNSUPPORTED. */
    const void* B2;
    int32_t hi_from_pass;
    int32_t epi_thr_stride; /* BNN_EPI_SIGN_I8: elements between the threshold rows of consecutive
                               batches (a multiple of 4); 0 = one row shared by all batches */
    const float* col_scale;
} bnn_gemm_desc;

int bnn_gemm(const bnn_gemm_desc* desc_host, int impl, bnn_stream_t stream);

/* cudaMemsetAsync(ptr, 0, bytes) on the stream: clears the accumulators that the fused epilogue
 * reductions (BNN_STATS_*) add into — a memset node under graph capture instead of a fill kernel. */
int bnn_zero_async(void* ptr, int64_t bytes, bnn_stream_t stream);

/* Variant A of the binary GEMM: XOR + popc on CUDA cores over packed sign bits.
 * D[m, n] = K - 2*popc(a_bits[m,:] ^ b_bits[n,:]); pad bits must be equal in A and B (the pack
 * kernels write zeros).  out_dtype: 1 = int32 (BNN_EPI_STORE_S32), 2 = int16. */
int bnn_gemm_popc(const uint32_t* a_bits, int64_t lda_w, const uint32_t* b_bits, int64_t ldb_w,
                  int M, int N, int K, void* D, int64_t ldd, int out_dtype, bnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * A4/A5/A6  batch-norm forward + sign / softmax      (BatchNormLayer.forward layers.py:74-99;
 *                                                      unitstep activation.py:106-113)
 * ------------------------------------------------------------------------------------------- */
enum { BNN_Z_F32 = 0, BNN_Z_S16 = 1, BNN_Z_S32 = 2 };

/* Per-column sum and sum of squares over the B rows of Z[B, N] (pitch ldz), accumulated in fp64
 * (exact and order independent for the integer-valued Z of +-1 x +-1 layers).  sums: [2, N]
 * doubles (sum, sumsq); accumulate=0 overwrites, 1 adds (used when a batch arrives in chunks). */
int bnn_colstats(const void* Z, int z_dtype, int B, int N, int64_t ldz, double* sums,
                 int accumulate, bnn_stream_t stream);

/* mu = sum/count, var = sumsq/count - mu^2 (population variance, Z.var(axis=0) layers.py:82).
 * count is the GLOBAL batch size when sums were all-reduced across ranks. */
int bnn_bn_stats_finalize(const double* sums, int N, double count, float* mu, float* var,
                          bnn_stream_t stream);

/* out = gamma * ((Z - mu) / sqrt(var + eps)) + beta          (train, layers.py:84-86)  or
 * out = (((Z - mu) / sqrt(sigma + eps)) * gamma) + beta      (eval,  layers.py:94-97; pass the
 * running mean / running variance `sigma` as mu / var).  Both are the same fp32 operation sequence
 * (sub, sqrt, div, mul, add — multiplication commutes bit-exactly); it is evaluated with explicit
 * round-to-nearest operations, no FMA contraction, so that for identical Z and parameters the sign
 * decisions equal NumPy's.  Outputs (any may be NULL):
 *   out_f32 [B, ld_f32]   BN output (pre-activation, cache["z"])
 *   act_i8  [B, ld_i8]    sign(out) as int8 +-1            (next layer's forward operand)
 *   act_t16 [N, ld_t16]   sign(out) transposed, fp16 +-1   (next layer's dW operand)
 *   act_bits[B, ld_bits]  sign(out) packed, 32 columns per word */
int bnn_bn_apply(const void* Z, int z_dtype, int B, int N, int64_t ldz, const float* mu,
                 const float* var, const float* gamma, const float* beta, float eps,
                 float* out_f32, int64_t ld_f32, int8_t* act_i8, int64_t ld_i8, void* act_t16,
                 int64_t ld_t16, uint32_t* act_bits, int64_t ld_bits, bnn_stream_t stream);

/* Per-column sign thresholds of the batch-norm + sign pair: out(z) is monotone in z, so
 * sign(out(z)) = +1  <=>  z * sgn[n] >= thr[n]; thr is found by bisection over float bit patterns
 * with exactly the operation sequence above, so decisions are bit-identical to bnn_bn_apply for
 * |z| <= 2^40 (pre-activations of a K-input +-1 layer satisfy |z| <= K). */
int bnn_bn_sign_thresholds(const float* mu, const float* var, const float* gamma, const float* beta,
                           float eps, int N, float* thr, float* sgn, bnn_stream_t stream);

/* The sign-only form of bnn_bn_apply on those thresholds (same outputs, no out_f32), plus optionally
 * the transposed int8 copies act_t8 [N, ld_t8] = +-1 and act_t8x127 = +-127: the A pieces of the
 * digit-split int8 dW product (bnn_gemm hi_from_pass), and act_f4 [B, ld_f4 / 2 bytes] = E2M1 nibbles (the A
 * operand of a BNN_DT_F4 product; layout of bnn_bits_to_f4, the buffer's padding nibbles are left untouched). */
int bnn_bn_sign(const void* Z, int z_dtype, int B, int N, int64_t ldz, const float* thr,
                const float* sgn, int8_t* act_i8, int64_t ld_i8, void* act_t16, int64_t ld_t16,
                uint32_t* act_bits, int64_t ld_bits, int8_t* act_t8, int8_t* act_t8x127, int64_t ld_t8,
                void* act_f4, int64_t ld_f4, bnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * A6/A10  output activation, loss and its gradient   (softmax activation.py:44-53;
 *         cross_entropy_loss loss.py:10-36; cross_entropy_derivative(with_softmax) loss.py:58-66)
 * ------------------------------------------------------------------------------------------- */
/* logits [B, C] (pitch ldl) -> probs [B, C] (softmax, then clipped to [1e-15, 1-1e-15] exactly as
 * loss.py:32 does in place), loss_rows[b] = -sum_c y*log(p), grad [B, C] = (p - y) / count,
 * loss_sum (device double) = sum_b loss_rows[b].  onehot: int8 [B, C] (may be NULL when only probs
 * are wanted); probs / loss_rows / grad / loss_sum may be NULL.  C <= 32. */
int bnn_softmax_xent(const float* logits, int B, int C, int64_t ldl, const int8_t* onehot,
                     float* probs, float* loss_rows, float* grad, double count, double* loss_sum,
                     bnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * 8f row 4  the other output-layer activations and losses      (activation.py:8-131, loss.py:10-131)
 * A binarized network may end in any of them (network.py:236-262, 377-387); all are element-wise or
 * row-wise maps over [B, C] that follow the reference's operation order.
 * ------------------------------------------------------------------------------------------- */
enum {
    BNN_ACT_IDENTITY = 0,     /* activation.py:8-14                                              */
    BNN_ACT_SIGMOID = 1,      /* 1 / (1 + exp(-x));          derivative dA * (a * (1 - a))  :17-28 */
    BNN_ACT_TANH = 2,         /* tanh(x);                    dA * (1 - a^2)                 :31-41 */
    BNN_ACT_RELU = 3,         /* x < 0 ? 0 : x;              a <= 0 ? 0 : dA                :65-77 */
    BNN_ACT_LEAKY_RELU = 4,   /* x < 0 ? 0.01 x : x;         a <= 0 ? 0.01 dA : dA          :80-92 */
    BNN_ACT_HARD_SIGMOID = 5, /* clip((x + 1) / 2, 0, 1);    (no derivative in the reference) :95-99 */
    BNN_ACT_HARD_TANH = 6     /* clip(x, -1, 1);             (a < -1 or a > 1) ? 0 * dA : dA :115-131 */
};
enum { BNN_LOSS_CROSS_ENTROPY = 0, BNN_LOSS_MSE = 1, BNN_LOSS_HINGE = 2 };

/* out[i] = f(x[i]), i < n (out may alias x). */
int bnn_activation(const float* x, int64_t n, int kind, float* out, bnn_stream_t stream);
/* out[i] = dA[i] * f'(.) evaluated on the activation a[i], as ACTIVATION_DERIVATIVES does (activation.py:160-168). */
int bnn_activation_grad(const float* dA, const float* a, int64_t n, int kind, float* out, bnn_stream_t stream);
/* softmax_derivative (activation.py:56-68), including its second softmax of the activations:
 * s = softmax(a[r, :]); out[r, j] = s_j * (dA[r, j] - sum_k s_k dA[r, k]).  Dense [R, C], any C. */
int bnn_softmax_grad(const float* dA, const float* a, int R, int C, float* out, bnn_stream_t stream);
/* rows[r] = loss of row r (loss.py:10-36 cross-entropy on arbitrary probabilities — ypred is clipped IN PLACE to
 * [1e-15, 1] like np.clip(..., out=ypred) — :71-89 mse, :112-119 hinge with labels in {-1, +1}); y int8 [R, C];
 * loss_sum (device double, may be NULL) = sum of the rows; rows may be NULL when only the sum is wanted. */
int bnn_loss_rows(float* ypred, const int8_t* y, int R, int C, int kind, float* rows, double* loss_sum,
                  bnn_stream_t stream);
/* out = dL/dA (loss.py:39-68 cross-entropy WITHOUT the softmax fusion, clipping ypred in place; :92-107 mse;
 * :122-131 hinge), divided by `count` (the global batch in data-parallel runs). */
int bnn_loss_grad(float* ypred, const int8_t* y, int R, int C, int kind, double count, float* out,
                  bnn_stream_t stream);


/* ---------------------------------------------------------------------------------------------
 * A7/A8  backward through sign (identity STE) and batch-norm, in-place SGD on gamma/beta,
 *        running statistics                            (layers.py:101-127, :252-258;
 *                                                       hard_tanh_derivative activation.py:122-131)
 * ------------------------------------------------------------------------------------------- */
/* red[0,n] = sum_b delta[b,n]; red[1,n] = sum_b delta[b,n]*(Z[b,n]-mu[n])      (doubles [2, N]);
 * red_max[0,n] = bits(max_b |delta[b,n]|); red_max[1,n] = bits(max_b |Z[b,n]-mu[n]|)  (uint32 [2, N],
 * float bit patterns).  accumulate=0 zeroes both first. */
int bnn_bn_bwd_reduce(const float* delta, int64_t ldd, const void* Z, int z_dtype, int64_t ldz,
                      int B, int N, const float* mu, double* red, uint32_t* red_max, int accumulate,
                      bnn_stream_t stream);

/* From the (all-reduced) reductions build the per-column coefficients of
 *   delta' = c0*delta + c1*(Z-mu) + c2         (== layers.py:107-113)
 * then apply `gamma -= lr*dGamma; beta -= lr*dBeta` (:114-118) and the running-statistic update
 * mu_run = 0.9*mu_run + 0.1*mu, sigma_run = 0.9*sigma_run + 0.1*var (:120-123).
 * coef: [3, N] floats.  bound_bits: device uint32 receiving bits(max_n bound |delta'|) via atomicMax
 * (zeroed by the call).  count = global batch size.  sums (may be NULL): the forward column sums,
 * used for the rounding residue mean(Z) - mu of layers.py:109-110.  dgamma_out / dbeta_out (may be
 * NULL) receive the parameter gradients. */
int bnn_bn_bwd_finalize(const double* red, const uint32_t* red_max, int N, double count,
                        const float* mu, const float* var, float eps, float* gamma, float* beta,
                        float* mu_run, float* sigma_run, float lr, double momentum,
                        const double* sums, float* coef, uint32_t* bound_bits, float* dgamma_out,
                        float* dbeta_out, float* col_quant,
                        bnn_stream_t stream);

/* delta' (see above) written as fp32 (out_f32, may be NULL) and/or as scaled fp16 hi/lo pieces in
 * both operand layouts: d_hi/d_lo [B, ld_rm] (dX operand) and dt_hi/dt_lo [N, ld_t] (dW operand).
 * bound: device float (the value behind bound_bits); scale_out receives the split scale. */
int bnn_bn_bwd_apply(const float* delta, int64_t ldd, const void* Z, int z_dtype, int64_t ldz,
                     int B, int N, const float* mu, const float* coef,
                     const float* bound, float* out_f32, int64_t ld_f32, void* d_hi, void* d_lo,
                     int64_t ld_rm, void* dt_hi, void* dt_lo, int64_t ld_t, float* scale_out,
                     bnn_stream_t stream);

/* Fixed-point form of the same pass for the int8 dW product.  col_quant: device [2, N] written by
 * bnn_bn_bwd_finalize (row 0: s_n = 4.1e6 / bound_n, so that |delta'_n * s_n| <= 4.1e6 — 22.97 bits of
 * the column's bound; row 1: 1 / s_n = the GEMM's col_scale).  Outputs: d_hi / d_lo [B, ld_rm] as above
 * (may be NULL), and the TRANSPOSED signed digits of
 *   q = rn(delta' * s_n) = q1 + 256 * (q2 + 127 * q3):  q1t, q3t, q2t int8 [N, ld_q] with
 * q1, q3 in [-128, 127] and q2 in [-63, 63] — the three B pieces of
 *   dW = (1 / s_n) * ( X^T.q1 + 256 * ( (127 X)^T.q3 + X^T.q2 ) )       (exact integer sums;
 * |high sum| <= B * 16319 < 2^31 for B <= 131 072 rows).  Norm-wise error of a dW column:
 * 7e-8 * bound_n / rms(delta'_n)  (uniform rounding of q, summed in quadrature over the batch). */
int bnn_bn_bwd_apply_q(const float* delta, int64_t ldd, const void* Z, int z_dtype, int64_t ldz,
                       int B, int N, const float* coef, const float* bound, const float* col_quant,
                       void* d_hi, void* d_lo, int64_t ld_rm, int8_t* q1t, int8_t* q3t, int8_t* q2t,
                       int64_t ld_q, float* scale_out, bnn_stream_t stream);

/* p[i] -= lr * sum_s g[s*part_stride + i]     (weights -= lr*dw, layers.py:268): the data-parallel path
 * (dW all-reduced before the update, parts = 1) and split-K dW products of narrow layers (parts > 1). */
int bnn_sgd_update(float* p, const float* g, int64_t n, float lr, int parts, int64_t part_stride,
                   bnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel exchange fused into the finalize kernels (new in this build; the reference is
 * single-process).  The producers (bnn_colstats / bnn_bn_bwd_reduce) write this rank's [2, N] fp64
 * partial sums into its slot of a symmetric buffer; these kernels signal every peer, wait for every
 * peer (bounded spin on a flag row, st.release.sys / ld.acquire.sys), read all ranks' slots over
 * NVLink peer mappings in rank order and finalize on the totals — one launch instead of an NCCL
 * all-reduce plus a finalize.  Every call site (layer x direction) owns one slot and one flag row.
 *   peer_data_host[r]  (host array of `world` device addresses) rank r's slot of this call site
 *   peer_flags_host[r] rank r's flag row of this call site, uint32[BNN_MAX_RANKS], zeroed once
 *   epoch / epoch_ptr  the training-step counter, advanced identically on every rank: either an
 *                      immediate (epoch >= 1, epoch_ptr NULL) or a device word read by the kernel
 *                      (bnn_step_tick increments it), which lets a whole step be replayed as a CUDA
 *                      graph.  The same convention holds for every *_p2p / bnn_peer_* call below.
 * ------------------------------------------------------------------------------------------- */
int bnn_step_tick(uint32_t* step_dev, bnn_stream_t stream);

int bnn_bn_stats_finalize_p2p(const uint64_t* peer_data_host, const uint64_t* peer_flags_host, int rank,
                              int world, uint32_t epoch, const uint32_t* epoch_ptr, int N, double count,
                              float* mu, float* var, double* sums_total, bnn_stream_t stream);

/* The same exchange without the finalize arithmetic: out[i] = sum_r slot_r[i], i < n (used for the
 * batch-norm backward reductions, whose finalize is bnn_bn_bwd_finalize). */
int bnn_allreduce_f64_p2p(const uint64_t* peer_data_host, const uint64_t* peer_flags_host, int rank,
                          int world, uint32_t epoch, const uint32_t* epoch_ptr, int n, double* out,
                          bnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel dW: reduce-scatter fused into the dW GEMM (BNN_EPI_SCATTER_F32 above), then the owner
 * of each row block sums the `world` partial products in rank order, applies `W -= lr * dW`
 * (layers.py:268) and writes the updated rows into every rank's copy of W over NVLink — the
 * all-gather half.  Synchronisation is by monotonically increasing epochs in flag pads
 * uint32[n_slots][BNN_MAX_RANKS] that live in symmetric memory (zero-initialised once):
 *   bnn_peer_signal   after this rank's GEMM: flags_of_peer[p][slot][rank] = epoch for every p
 *   bnn_dw_reduce_sgd_p2p   waits until local flags[slot_pushed][*] >= epoch, reduces + updates + broadcasts
 *                     its rows, then (last block) sets flags_of_peer[p][slot_landed][rank] = epoch
 *   bnn_peer_wait     spins until local flags[slot][r] >= epoch for all r (before W is next read)
 *   bnn_peer_broadcast_rows   the copy-engine form of the all-gather
 * All waits are bounded (10 s -> trap).  peer_*_host are host arrays of `world` device addresses.
 * ------------------------------------------------------------------------------------------- */
int bnn_peer_signal(const uint64_t* peer_flags_host, int slot, int rank, int world, uint32_t epoch,
                    const uint32_t* epoch_ptr, bnn_stream_t stream);
/* waits on the n_slots flag rows slot, slot + slot_stride, ... in one launch */
int bnn_peer_wait(const uint32_t* flags_local, int slot, int slot_stride, int n_slots, int rank,
                  int world, uint32_t epoch, const uint32_t* epoch_ptr, bnn_stream_t stream);
/* stage_local: this rank's staging buffer float[world][rows_per_owner][ld]; W: [K, N] fp32 row-major
 * with pitch ld == N on every rank (peer_w_host[r] = rank r's W as mapped here; may alias W for
 * r == rank).  counter: device uint32, zero-initialised once (self-resetting block counter).
 * broadcast bit 1 (value 2): the kernel also performs this rank's bnn_peer_signal(slot_pushed) itself
 * (valid when it directly follows the dW GEMM on the same stream).
 * broadcast bit 0 clear: updates only the local rows and writes `epoch` to *epoch_word (device uint32): the
 * caller then runs bnn_peer_broadcast_rows on another stream, which moves the rows with the copy
 * engines while the compute stream carries on with the backward pass. */
int bnn_dw_reduce_sgd_p2p(const float* stage_local, const uint64_t* peer_w_host,
                          const uint64_t* peer_flags_host, int slot_pushed, int slot_landed, int rank,
                          int world, uint32_t epoch, const uint32_t* epoch_ptr, int K, int N,
                          int rows_per_owner, float lr, int broadcast, uint32_t* counter,
                          uint32_t* epoch_word, bnn_stream_t stream);

/* All-gather by copy engine: for every rank p, cudaMemcpyAsync of this rank's owned rows of W into
 * rank p's W, followed by a 4-byte copy of *epoch_word into flags_of_peer[p][slot_landed][rank] (also
 * for p == rank).  Stream order makes the flag land after the rows.  No kernel is launched. */
int bnn_peer_broadcast_rows(const uint64_t* peer_w_host, const uint64_t* peer_flags_host,
                            int slot_landed, int rank, int world, int K, int N, int rows_per_owner,
                            const uint32_t* epoch_word, bnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * A11  weight fault injection                          (ErrorCallback._flipBNNMask
 *                                                       callbacks.py:79-83, dispatch :94-98)
 * Flips act on the K-major binarized copy WbT[N, K]; element (n, k) has linear index e = n*K + k.
 * In-kernel RNG: Philox4x32-10, key = (seed_lo, seed_hi), counter = (e/4 lo, e/4 hi, stream_id,
 * trial); word e%4 of the output decides element e: flip iff word < threshold.
 * ------------------------------------------------------------------------------------------- */
/* Bernoulli(p) flips for trials [trial0, trial0+n_trials) in ONE launch.  threshold =
 * floor(p * 2^32) as uint64 (p = 1 -> 2^32 flips everything).  Inputs: clean wt_i8 [N, ld_i8] and/or
 * wt_bits [N, ld_bits].  Outputs per trial t (any may be NULL), trial stride given in elements:
 *   out_i8 [t][N, ld_i8], out_f16 [t][N, ld_f16], out_bits [t][N, ld_bits],
 *   mask_bits [t][N, ld_bits] (the flip mask itself, packed like wt_bits). */
int bnn_flip_weights_bernoulli(const int8_t* wt_i8, int64_t ld_i8, const uint32_t* wt_bits,
                               int64_t ld_bits, int N, int K, uint64_t threshold, uint64_t seed,
                               uint32_t stream_id, uint32_t trial0, int n_trials, int8_t* out_i8,
                               int64_t out_i8_stride, void* out_f16, int64_t ld_f16,
                               int64_t out_f16_stride, uint32_t* out_bits, int64_t out_bits_stride,
                               uint32_t* mask_bits, int64_t mask_stride, bnn_stream_t stream);

/* Exact-k flips without replacement over a weight space of `total` elements (all layers of the
 * network concatenated, this layer owning [offset, offset + N*K)).  Position i of trial t is
 * P_t(i), i < k_t, where P_t is a keyed cycle-walking Feistel bijection on [0, total) (key from
 * Philox(seed, trial)); positions are distinct by construction.  k_per_trial: device int32
 * [n_trials].  The outputs must already hold a copy of the clean weights (in/out). */
int bnn_flip_weights_exactk(int N, int K, uint64_t total, uint64_t offset, uint64_t seed,
                            uint32_t trial0, int n_trials, const int32_t* k_per_trial, int k_max,
                            int8_t* io_i8, int64_t ld_i8, int64_t io_i8_stride, void* io_f16,
                            int64_t ld_f16, int64_t io_f16_stride, uint32_t* io_bits,
                            int64_t ld_bits, int64_t io_bits_stride, bnn_stream_t stream);

/* External-mask mode: apply a caller-supplied boolean mask[K, N] (uint8, the reference's own
 * `uniform(size=W.shape) < p`, callbacks.py:81) to the K-major copies. */
int bnn_flip_weights_mask(const uint8_t* mask_kn, int K, int N, int8_t* io_i8, int64_t ld_i8,
                          void* io_f16, int64_t ld_f16, uint32_t* io_bits, int64_t ld_bits,
                          bnn_stream_t stream);

/* IEEE-754 bit flips of real-valued fp32 weights — the non-BNN branch of ErrorCallback
 * (_flipFunc callbacks.py:35-59, _flip :85-92, __call__ :100-123), in place on w[n]:
 *   n_sel = round(n * p) distinct elements (the caller rounds like Python: half to even) are chosen —
 *   element i of the draw is P(i), P a keyed Feistel bijection on [0, n) (key from Philox(seed, trial));
 *   in each, the candidate bits follow the reference: with the exponent MSB (bit 30) set the other
 *   exponent bits (29..23) are protected, otherwise bit 30 itself; mode 0 keeps only 0-bits, mode 1
 *   only 1-bits, mode 2 all; rint(cnt * p) of the cnt candidates are inverted, chosen by a second
 *   bijection keyed by (seed, trial, element) over the candidates ranked in the reference's bit order
 *   (np.unpackbits of the little-endian bytes: rank j <-> word bit 8*(j/8) + 7 - j%8).
 * n < 2^32. */
int bnn_flip_f32_bits(float* w, int64_t n, int64_t n_sel, double p, int mode, uint64_t seed,
                      uint32_t trial, bnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * A12  image fault injection                           (ImageErrorCallback.intError
 *                                                       callbacks.py:130-147, __call__ :149-172)
 * Per image row of F bytes: flip exactly min(k, candidates) distinct bits, MSB-first bit order
 * (np.unpackbits); mode 2 = any bit, 0 = among 0-bits, 1 = among 1-bits.  The positions are the
 * first k outputs of a keyed Feistel bijection over the candidate ranks, key from
 * Philox(seed, trial, row).  in: [R, F] uint8 (shared by all trials); out: [n_trials][R, F].
 * F <= 4096. */
int bnn_flip_images(const uint8_t* in, int R, int F, int k, int mode, uint64_t seed, uint32_t trial0,
                    int n_trials, uint8_t* out, bnn_stream_t stream);

/* The same flips with two by-products that let layer 0 consume the corrupted uint8 images directly
 * (no fp32 copy, no separate min/max pass):
 *   xor_out  every output byte is XOR-ed with this value (0, or 0x80: the byte then reads as the int8
 *            u - 128, the operand form of a plain s8 x s8 product);
 *   minmax   device uint32 [n_trials][2] (may be NULL) <- {min, max} over all R*F corrupted bytes of the
 *            trial (of the un-XOR-ed values): the xmin / xmax of utils.normalize (utils.py:181-182).
 *            Initialised by the call. */
int bnn_flip_images_ex(const uint8_t* in, int R, int F, int k, int mode, uint64_t seed, uint32_t trial0,
                       int n_trials, uint8_t* out, int xor_out, uint32_t* minmax, bnn_stream_t stream);

/* utils.normalize (utils.py:178-187) + layer-0 batch-norm + sign folded into per-trial integer
 * thresholds, so that layer 0 runs as an exact integer GEMM on the uint8 images:
 *   x = (u - mn) * (new_max - new_min) / (mx - mn) + new_min,   z_n = sum_k x_k * Wb[k, n]
 *       = r * (acc_n - mn * S_n) + new_min * S_n,   r = (new_max - new_min) / (mx - mn),
 *   acc_n = sum_k u_k * Wb[k, n] (what the int8 tensor pipe computes), S_n = sum_k Wb[k, n]
 * and sign(BN(z_n)) = +1  <=>  z_n * sgn[n] >= thr[n]  (bnn_bn_sign_thresholds)
 *                         <=>  g_n * sgn[n] >= thr_out[t, n]
 * where g_n = acc_n - bias * S_n is the GEMM result (bias = 0 for a_unsigned operands, 128 for
 * xor_out = 0x80 int8 operands).  thr_out[t, n] is the ceiling of the real threshold (g_n*sgn[n] is an
 * integer), evaluated in fp64; +-inf thresholds stay +-inf.  A trial with mx == mn (the reference
 * divides by zero there) gets +inf (every unit -1).
 *   wt_bits [N, ld_bits]  packed K-major binarized weights (bnn_binarize_weights_t; pad bits 0);
 *                         trial t reads wt_bits + t*bits_stride (0 = one matrix shared by all trials:
 *                         image faults; N*ld_bits or more = the per-trial copies bnn_flip_weights_*
 *                         writes: weight faults, whose S_n differs per trial)
 *   minmax  trial t reads {min, max} at minmax + t*minmax_stride (2 = per trial, from
 *                         bnn_flip_images_ex; 0 = one image set shared by all trials, bnn_minmax_u8)
 *   thr_out [n_trials, ld_thr]. */
int bnn_u8_input_thresholds(const uint32_t* wt_bits, int64_t ld_bits, int64_t bits_stride, int N, int K,
                            const float* thr, const float* sgn, const uint32_t* minmax,
                            int minmax_stride, int n_trials, float new_min, float new_max, int bias,
                            float* thr_out, int64_t ld_thr, bnn_stream_t stream);

/* uint8 -> fp32 affine map used by utils.normalize (utils.py:178-187) with caller-supplied
 * min/max (device floats): y = ((x - mn) * (new_max - new_min)) / (mx - mn) + new_min, in the
 * reference's operation order. */
int bnn_normalize_u8(const uint8_t* in, int64_t n, const float* mn, const float* mx, float new_min,
                     float new_max, float* out, bnn_stream_t stream);

/* Global min/max of a uint8 buffer: minmax (device uint32[2]) <- {min, max}; minmax_f32 (device
 * float[2], may be NULL) receives the same as floats for bnn_normalize_u8. */
int bnn_minmax_u8(const uint8_t* in, int64_t n, uint32_t* minmax, float* minmax_f32,
                  bnn_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * A13  evaluation helper                                (predict_classes network.py:401-404,
 *                                                        evaluate :406-421)
 * correct[g] += #{ b : argmax_c scores[g*B + b, c] == labels[b] } for g < groups (first maximum
 * wins, as ndarray.argmax).  groups > 1 evaluates several fault trials stacked along the rows
 * against one set of labels; it needs B >= 32.
 * ------------------------------------------------------------------------------------------- */
int bnn_count_correct(const float* scores, int B, int C, int64_t lds, const int32_t* labels,
                      int32_t* correct, int groups, bnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BNN_B200_H */
