"""Typed Python front-ends of the libbnn_b200 entry points (one function per C symbol).

Arguments are CUDA torch tensors used purely as device buffers; shapes/pitches are derived from
them and passed to the C ABI together with raw pointers and torch's current stream.  Nothing here
computes on the host and nothing falls back to another backend.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _native as nat
from .device import ptr, stream_ptr

launch_count = 0  # kernels launched through this module (bench.py reports it as gpu_launches)
gemm_events = None  # bench.py sets this to a list to collect (start, end, kind, 2MNK, passes) per GEMM

# launches (kernels + memsets are not counted) issued per C call
_LAUNCHES = {"bnn_minmax_u8": 2, "bnn_peer_broadcast_rows": 0, "bnn_zero_async": 0,
             "bnn_flip_images_ex": 2}


def _call(name, *args):
    global launch_count
    nat.call(name, *args, stream_ptr())
    launch_count += _LAUNCHES.get(name, 1)


def gemm_impl() -> int:
    """Product path is tcgen05; BNN_GEMM_IMPL=simt selects the CUDA-core cross-check kernel."""
    v = os.environ.get("BNN_GEMM_IMPL", "tcgen05").lower()
    if v in ("tcgen05", "tc"):
        return nat.GEMM_TCGEN05
    if v == "simt":
        return nat.GEMM_SIMT
    raise ValueError(f"BNN_GEMM_IMPL={v!r}: expected 'tcgen05' or 'simt'")


def _ld(t: torch.Tensor) -> int:
    assert t.stride(-1) == 1, "innermost dimension must be contiguous"
    return t.stride(-2) if t.ndim >= 2 else t.shape[0]


# ------------------------------------------------------------------------------- packing
def pack_sign_rows(x: torch.Tensor, bits: torch.Tensor) -> None:
    R, Cc = x.shape
    _call("bnn_pack_sign_rows", ptr(x), R, Cc, _ld(x), ptr(bits), _ld(bits))


def pack_i8_rows(x: torch.Tensor, C_: int, bits: torch.Tensor) -> None:
    _call("bnn_pack_i8_rows", ptr(x), x.shape[0], C_, _ld(x), ptr(bits), _ld(bits))


def binarize_weights_t(W: torch.Tensor, wt_i8=None, wt_f16=None, wt_bits=None, absmax=None, wt_f4=None) -> None:
    K, N = W.shape
    assert W.is_contiguous()
    _call("bnn_binarize_weights_t", ptr(W), K, N,
          ptr(wt_i8), _ld(wt_i8) if wt_i8 is not None else 0,
          ptr(wt_f16), _ld(wt_f16) if wt_f16 is not None else 0,
          ptr(wt_bits), _ld(wt_bits) if wt_bits is not None else 0,
          ptr(wt_f4), 2 * _ld(wt_f4) if wt_f4 is not None else 0, ptr(absmax))


def unpack_bits(bits: torch.Tensor, C_: int, out_i8=None, out_f16=None) -> None:
    _call("bnn_unpack_bits", ptr(bits), bits.shape[0], C_, _ld(bits),
          ptr(out_i8), _ld(out_i8) if out_i8 is not None else 0,
          ptr(out_f16), _ld(out_f16) if out_f16 is not None else 0)


def gather_rows(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor, bad=None) -> None:
    """dst[i] = src[idx[i]] for 2-D row-major tensors of one dtype (rows may be pitched); idx: device int64."""
    assert src.ndim == 2 and dst.ndim == 2 and src.dtype == dst.dtype and src.shape[1] == dst.shape[1]
    assert idx.dtype == torch.int64 and idx.is_contiguous() and idx.numel() == dst.shape[0]
    assert src.stride(1) == 1 and dst.stride(1) == 1
    es = src.element_size()
    _call("bnn_gather_rows", ptr(src), src.stride(0) * es, src.shape[0], ptr(idx), dst.shape[0],
          src.shape[1] * es, ptr(dst), dst.stride(0) * es, ptr(bad))


def bits_to_f4(bits: torch.Tensor, C_: int, out_f4: torch.Tensor) -> None:
    """Packed sign words [..., R, W] -> E2M1 nibbles [..., R, ld/2 bytes] (uint8 storage; leading dims contiguous)."""
    assert bits.is_contiguous() and out_f4.is_contiguous() and out_f4.dtype in (torch.uint8, torch.int8)
    rows = bits.numel() // bits.shape[-1]
    assert out_f4.numel() // out_f4.shape[-1] == rows
    _call("bnn_bits_to_f4", ptr(bits), rows, C_, bits.shape[-1], ptr(out_f4), 2 * out_f4.shape[-1])


def f4_ok(Kd: int, N: int) -> bool:
    """A +-1 x +-1 product [*, Kd] x [N, Kd] goes to the FP4 tensor pipe: tcgen05 path and K in whole
    16-byte words.  N is free: a 10-wide output layer reads its input as nibbles too — half the bytes of the one
    operand that makes that product HBM-bound.  BNN_FP4=0: int8."""
    return (Kd % 32 == 0 and N >= 1 and gemm_impl() == nat.GEMM_TCGEN05
            and os.environ.get("BNN_FP4", "1") != "0")


def f4_train_ok(Kd: int, N: int) -> bool:
    """The TRAINING forward of a +-1 x +-1 layer on the FP4 pipe (shapes as f4_ok; the batch-norm column sums are
    fused into that kind's epilogue on any width).  BNN_FP4_TRAIN=0: int8."""
    return f4_ok(Kd, N) and os.environ.get("BNN_FP4_TRAIN", "1") != "0"


def absmax_f32(x: torch.Tensor, out_bits: torch.Tensor, zero_first=True) -> None:
    assert x.is_contiguous()
    _call("bnn_absmax_f32", ptr(x), x.numel(), ptr(out_bits), int(zero_first))


def split_f16(x: torch.Tensor, bound: torch.Tensor, hi=None, lo=None, hi_t=None, lo_t=None,
              scale_out=None) -> None:
    R, Cc = x.shape
    _call("bnn_split_f16", ptr(x), R, Cc, _ld(x), ptr(bound), ptr(hi), ptr(lo),
          _ld(hi) if hi is not None else 0, ptr(hi_t), ptr(lo_t),
          _ld(hi_t) if hi_t is not None else 0, ptr(scale_out))


# ------------------------------------------------------------------------------- GEMMs
def gemm(*, M, N, K, A, B, lda, ldb, D, ldd, in_dtype, epilogue, passes=((0, 0),), sa=None,
         sb=None, host_scale=1.0, batch=1, stride_a=0, stride_b=0, stride_d=0, impl=None,
         epi_thr=None, epi_sgn=None, scatter=None, stats=None, hi_from_pass=0, col_scale=None,
         a_unsigned=False, epi_thr_stride=0) -> None:
    """D = epi(sum_p A[pa] @ B[pb]^T); A/B are sequences of 1..2 operand-piece tensors.
    scatter (EPI_SCATTER_F32): dict(world, rank, rows_per_owner, stage=[address per rank]).
    stats: dict(mode=STATS_COLSUM, sums) or dict(mode=STATS_BN_BWD, sums, max, Z, mu) — column
    reductions of the result accumulated by the epilogue (see fused_stats_ok)."""
    d = nat.GemmDesc()
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.in_dtype, d.epilogue, d.n_pass = in_dtype, epilogue, len(passes)
    for i, (a, b) in enumerate(passes):
        d.pa[i], d.pb[i] = a, b
    for i in range(2):
        d.A[i] = ptr(A[i]) if i < len(A) else None
        d.B[i] = ptr(B[i]) if i < len(B) else None
    d.B2 = ptr(B[2]) if len(B) > 2 else None
    d.hi_from_pass, d.col_scale = hi_from_pass, ptr(col_scale)
    d.lda, d.ldb, d.stride_a, d.stride_b = lda, ldb, stride_a, stride_b
    d.D, d.ldd, d.stride_d = ptr(D), ldd, stride_d
    d.sa, d.sb = ptr(sa), ptr(sb)
    d.host_scale = host_scale
    d.epi_thr, d.epi_sgn = ptr(epi_thr), ptr(epi_sgn)
    d.a_unsigned, d.epi_thr_stride = int(a_unsigned), epi_thr_stride
    if stats is not None:
        d.stats, d.stats_sums = stats["mode"], ptr(stats["sums"])
        if stats["mode"] == nat.STATS_BN_BWD:
            Zs = stats["Z"]
            d.stats_max, d.stats_z, d.stats_mu = ptr(stats["max"]), ptr(Zs), ptr(stats["mu"])
            d.stats_ldz, d.stats_z_dtype = _ld(Zs), z_dtype_of(Zs)
    if scatter is not None:
        sc = nat.PeerScatter()
        sc.world, sc.rank, sc.rows_per_owner = scatter["world"], scatter["rank"], scatter["rows_per_owner"]
        for r, a in enumerate(scatter["stage"]):
            sc.stage[r] = a
        d.scatter = C.pointer(sc)  # keeps `sc` alive for the duration of the call
    if gemm_events is None:
        _call("bnn_gemm", C.byref(d), gemm_impl() if impl is None else impl)
        return
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    _call("bnn_gemm", C.byref(d), gemm_impl() if impl is None else impl)
    end.record()
    kind = {nat.DT_F16: "f16", nat.DT_F4: "f4"}.get(in_dtype, "i8dw" if hi_from_pass else "i8")
    gemm_events.append((start, end, kind, 2.0 * M * N * K * batch, len(passes)))


def full_tiles(N: int, impl=None) -> bool:
    """tcgen05 path and N a multiple of its tile width (256 for N > 128, 128 for N > 32, else 32)."""
    if (gemm_impl() if impl is None else impl) != nat.GEMM_TCGEN05:
        return False
    tile = 256 if N > 128 else (128 if N > 32 else 32)
    return N % tile == 0


def int8_dw_ok(N: int, B: int = 0) -> bool:
    """dW = X^T . delta' as an exact fixed-point product on the int8 tensor pipe (X = +-1 activations,
    delta' in three signed digits) instead of two fp16 passes.  B: rows contracted over — the HIGH int32
    accumulator holds sums of B terms of magnitude <= 16319, so B <= 131072.  BNN_DW_INT8=0 keeps the
    fp16 form."""
    return (N >= 64 and B <= 131072 and full_tiles(N)
            and os.environ.get("BNN_DW_INT8", "1") != "0")


def zero(t: torch.Tensor) -> None:
    """Clear a contiguous device buffer with a memset on the current stream."""
    assert t.is_contiguous()
    _call("bnn_zero_async", ptr(t), t.numel() * t.element_size())


def fused_stats_ok(N: int, impl=None, f4: bool = False) -> bool:
    """Shapes on which the tcgen05 epilogue can carry the column reductions: every tile full-width
    (tile = 256 columns for N > 128, 128 for N > 32, else 32).  The FP4 kind has integer outputs only, whose
    sums are guarded per column: any N."""
    if (gemm_impl() if impl is None else impl) != nat.GEMM_TCGEN05:
        return False
    if os.environ.get("BNN_FUSE_STATS", "1") == "0":
        return False
    tile = 256 if N > 128 else (128 if N > 32 else 32)
    return f4 or N % tile == 0


def gemm_popc(a_bits, b_bits, M, N, K, D, out_s16=True) -> None:
    _call("bnn_gemm_popc", ptr(a_bits), _ld(a_bits), ptr(b_bits), _ld(b_bits), M, N, K, ptr(D),
          _ld(D), nat.EPI_STORE_S16 if out_s16 else nat.EPI_STORE_S32)


# ------------------------------------------------------------------------------- batch norm
def z_dtype_of(Z: torch.Tensor) -> int:
    return {torch.float32: nat.Z_F32, torch.int16: nat.Z_S16, torch.int32: nat.Z_S32}[Z.dtype]


def colstats(Z, B, N, sums, accumulate=False) -> None:
    _call("bnn_colstats", ptr(Z), z_dtype_of(Z), B, N, _ld(Z), ptr(sums), int(accumulate))


def bn_stats_finalize(sums, N, count, mu, var) -> None:
    _call("bnn_bn_stats_finalize", ptr(sums), N, float(count), ptr(mu), ptr(var))


def _peer_arrays(call):
    n = call["world"]
    return (C.c_uint64 * n)(*call["data"]), (C.c_uint64 * n)(*call["flags"])


def step_tick(step_dev) -> None:
    _call("bnn_step_tick", ptr(step_dev))


def bn_stats_finalize_p2p(call, N, count, mu, var, sums_total) -> None:
    data, flags = _peer_arrays(call)
    _call("bnn_bn_stats_finalize_p2p", data, flags, call["rank"], call["world"], 0, ptr(call["step"]), N,
          float(count), ptr(mu), ptr(var), ptr(sums_total))


def allreduce_f64_p2p(call, n, out) -> None:
    data, flags = _peer_arrays(call)
    _call("bnn_allreduce_f64_p2p", data, flags, call["rank"], call["world"], 0, ptr(call["step"]), n,
          ptr(out))


def _u64_array(values):
    return (C.c_uint64 * len(values))(*values)


def peer_signal(peer_flags, slot, rank, world, step_dev) -> None:
    _call("bnn_peer_signal", _u64_array(peer_flags), slot, rank, world, 0, ptr(step_dev))


def peer_wait(flags_local_ptr, slot, slot_stride, n_slots, rank, world, step_dev) -> None:
    _call("bnn_peer_wait", C.c_void_p(flags_local_ptr), slot, slot_stride, n_slots, rank, world, 0,
          ptr(step_dev))


def dw_reduce_sgd_p2p(stage_local_ptr, peer_w, peer_flags, slot_pushed, slot_landed, rank, world,
                      step_dev, K, N, rows_per_owner, lr, broadcast, counter, signal=False) -> None:
    _call("bnn_dw_reduce_sgd_p2p", C.c_void_p(stage_local_ptr), _u64_array(peer_w),
          _u64_array(peer_flags), slot_pushed, slot_landed, rank, world, 0, ptr(step_dev), K, N,
          rows_per_owner, lr, int(bool(broadcast)) | (2 if signal else 0), ptr(counter), None)


def peer_broadcast_rows(peer_w, peer_flags, slot_landed, rank, world, K, N, rows_per_owner,
                        epoch_word) -> None:
    """Copy-engine all-gather of this rank's rows of W (no kernel); runs on torch's current stream."""
    _call("bnn_peer_broadcast_rows", _u64_array(peer_w), _u64_array(peer_flags), slot_landed, rank,
          world, K, N, rows_per_owner, ptr(epoch_word))


def bn_apply(Z, B, N, mu, var, gamma, beta, eps, out_f32=None, act_i8=None, act_t16=None,
             act_bits=None) -> None:
    _call("bnn_bn_apply", ptr(Z), z_dtype_of(Z), B, N, _ld(Z), ptr(mu), ptr(var), ptr(gamma),
          ptr(beta), eps,
          ptr(out_f32), _ld(out_f32) if out_f32 is not None else 0,
          ptr(act_i8), _ld(act_i8) if act_i8 is not None else 0,
          ptr(act_t16), _ld(act_t16) if act_t16 is not None else 0,
          ptr(act_bits), _ld(act_bits) if act_bits is not None else 0)


def bn_sign_thresholds(mu, var, gamma, beta, eps, N, thr, sgn) -> None:
    _call("bnn_bn_sign_thresholds", ptr(mu), ptr(var), ptr(gamma), ptr(beta), eps, N, ptr(thr),
          ptr(sgn))


def bn_sign(Z, B, N, thr, sgn, act_i8=None, act_t16=None, act_bits=None, act_t8=None,
            act_t8x127=None, act_f4=None) -> None:
    _call("bnn_bn_sign", ptr(Z), z_dtype_of(Z), B, N, _ld(Z), ptr(thr), ptr(sgn),
          ptr(act_i8), _ld(act_i8) if act_i8 is not None else 0,
          ptr(act_t16), _ld(act_t16) if act_t16 is not None else 0,
          ptr(act_bits), _ld(act_bits) if act_bits is not None else 0,
          ptr(act_t8), ptr(act_t8x127), _ld(act_t8) if act_t8 is not None else 0,
          ptr(act_f4), 2 * _ld(act_f4) if act_f4 is not None else 0)


def activation(x: torch.Tensor, kind: int, out: torch.Tensor) -> None:
    assert x.is_contiguous() and out.is_contiguous() and x.dtype == out.dtype == torch.float32
    _call("bnn_activation", ptr(x), x.numel(), kind, ptr(out))


def activation_grad(dA: torch.Tensor, a: torch.Tensor, kind: int, out: torch.Tensor) -> None:
    assert dA.is_contiguous() and a.is_contiguous() and out.is_contiguous() and dA.shape == a.shape
    _call("bnn_activation_grad", ptr(dA), ptr(a), a.numel(), kind, ptr(out))


def softmax_grad(dA: torch.Tensor, a: torch.Tensor, out: torch.Tensor) -> None:
    R, Cc = a.shape
    assert dA.is_contiguous() and a.is_contiguous() and out.is_contiguous()
    _call("bnn_softmax_grad", ptr(dA), ptr(a), R, Cc, ptr(out))


def loss_rows(ypred: torch.Tensor, y: torch.Tensor, kind: int, rows=None, loss_sum=None) -> None:
    R, Cc = ypred.shape
    assert ypred.is_contiguous() and y.is_contiguous() and y.dtype == torch.int8 and tuple(y.shape) == (R, Cc)
    _call("bnn_loss_rows", ptr(ypred), ptr(y), R, Cc, kind, ptr(rows), ptr(loss_sum))


def loss_grad(ypred: torch.Tensor, y: torch.Tensor, kind: int, count: float, out: torch.Tensor) -> None:
    R, Cc = ypred.shape
    assert ypred.is_contiguous() and y.is_contiguous() and y.dtype == torch.int8 and out.is_contiguous()
    _call("bnn_loss_grad", ptr(ypred), ptr(y), R, Cc, kind, float(count), ptr(out))


def softmax_xent(logits, B, Cc, onehot=None, probs=None, loss_rows=None, grad=None, count=None,
                 loss_sum=None) -> None:
    _call("bnn_softmax_xent", ptr(logits), B, Cc, _ld(logits), ptr(onehot), ptr(probs),
          ptr(loss_rows), ptr(grad), float(count if count is not None else B), ptr(loss_sum))


def bn_bwd_reduce(delta, Z, B, N, mu, red, red_max, accumulate=False) -> None:
    _call("bnn_bn_bwd_reduce", ptr(delta), _ld(delta), ptr(Z), z_dtype_of(Z), _ld(Z), B, N,
          ptr(mu), ptr(red), ptr(red_max), int(accumulate))


def bn_bwd_finalize(red, red_max, N, count, mu, var, eps, gamma, beta, mu_run, sigma_run, lr,
                    momentum, sums, coef, bound_bits, dgamma_out=None, dbeta_out=None,
                    col_quant=None) -> None:
    _call("bnn_bn_bwd_finalize", ptr(red), ptr(red_max), N, float(count), ptr(mu), ptr(var), eps,
          ptr(gamma), ptr(beta), ptr(mu_run), ptr(sigma_run), lr, momentum, ptr(sums), ptr(coef),
          ptr(bound_bits), ptr(dgamma_out), ptr(dbeta_out), ptr(col_quant))


def bn_bwd_apply_q(delta, Z, B, N, coef, bound, col_quant, q1t, q3t, q2t, d_hi=None, d_lo=None,
                   scale_out=None) -> None:
    _call("bnn_bn_bwd_apply_q", ptr(delta), _ld(delta), ptr(Z), z_dtype_of(Z), _ld(Z), B, N, ptr(coef),
          ptr(bound), ptr(col_quant), ptr(d_hi), ptr(d_lo), _ld(d_hi) if d_hi is not None else 0,
          ptr(q1t), ptr(q3t), ptr(q2t), _ld(q1t), ptr(scale_out))


def bn_bwd_apply(delta, Z, B, N, mu, coef, bound, out_f32=None, d_hi=None, d_lo=None, dt_hi=None,
                 dt_lo=None, scale_out=None) -> None:
    _call("bnn_bn_bwd_apply", ptr(delta), _ld(delta), ptr(Z), z_dtype_of(Z), _ld(Z), B, N, ptr(mu),
          ptr(coef), ptr(bound),
          ptr(out_f32), _ld(out_f32) if out_f32 is not None else 0,
          ptr(d_hi), ptr(d_lo), _ld(d_hi) if d_hi is not None else 0,
          ptr(dt_hi), ptr(dt_lo), _ld(dt_hi) if dt_hi is not None else 0, ptr(scale_out))


def sgd_update(p, g, lr, parts=1) -> None:
    """p -= lr * sum over the `parts` leading slices of g (g: [parts, *p.shape] or p.shape)."""
    assert p.is_contiguous() and g.is_contiguous() and p.numel() * parts == g.numel()
    _call("bnn_sgd_update", ptr(p), ptr(g), p.numel(), lr, parts, p.numel())


def count_correct(scores, B, Cc, labels_i32, correct_i32, groups=1) -> None:
    _call("bnn_count_correct", ptr(scores), B, Cc, _ld(scores), ptr(labels_i32), ptr(correct_i32),
          groups)


# ------------------------------------------------------------------------------- fault injection
def bernoulli_threshold(p: float) -> int:
    """floor(p * 2^32) as the kernels' 64-bit compare threshold (p >= 1 flips everything)."""
    if p <= 0:
        return 0
    return min(1 << 32, int(p * 4294967296.0))


def flip_weights_bernoulli(*, N, K, p, seed, stream_id, trial0, n_trials, wt_i8=None, wt_bits=None,
                           out_i8=None, out_f16=None, out_bits=None, mask_bits=None) -> None:
    """Outputs are [n_trials, N, ld] tensors (or None)."""
    def st(t):
        return t.stride(0) if t is not None else 0

    ld_bits = 0
    for t in (wt_bits, out_bits, mask_bits):
        if t is not None:
            ld_bits = t.stride(-2)
    ld_i8 = 0
    for t in (wt_i8, out_i8):
        if t is not None:
            ld_i8 = t.stride(-2)
    _call("bnn_flip_weights_bernoulli", ptr(wt_i8), ld_i8, ptr(wt_bits), ld_bits, N, K,
          bernoulli_threshold(p), seed & (2**64 - 1), stream_id, trial0, n_trials,
          ptr(out_i8), st(out_i8), ptr(out_f16), out_f16.stride(-2) if out_f16 is not None else 0,
          st(out_f16), ptr(out_bits), st(out_bits), ptr(mask_bits), st(mask_bits))


def flip_weights_exactk(*, N, K, total, offset, seed, trial0, k_per_trial, k_max, io_i8=None,
                        io_f16=None, io_bits=None) -> None:
    n_trials = k_per_trial.numel()

    def st(t):
        return t.stride(0) if (t is not None and t.ndim == 3) else 0

    def ld(t):
        return t.stride(-2) if t is not None else 0

    _call("bnn_flip_weights_exactk", N, K, total, offset, seed & (2**64 - 1), trial0, n_trials,
          ptr(k_per_trial), k_max, ptr(io_i8), ld(io_i8), st(io_i8), ptr(io_f16), ld(io_f16),
          st(io_f16), ptr(io_bits), ld(io_bits), st(io_bits))


def flip_weights_mask(mask_kn_u8, K, N, io_i8=None, io_f16=None, io_bits=None) -> None:
    def ld(t):
        return t.stride(-2) if t is not None else 0

    _call("bnn_flip_weights_mask", ptr(mask_kn_u8), K, N, ptr(io_i8), ld(io_i8), ptr(io_f16),
          ld(io_f16), ptr(io_bits), ld(io_bits))


def flip_f32_bits(w_f32, p, mode, seed, trial) -> None:
    """In place: IEEE-754 bit flips in round(size*p) elements of a contiguous fp32 tensor."""
    assert w_f32.is_contiguous() and w_f32.dtype == torch.float32
    n = w_f32.numel()
    _call("bnn_flip_f32_bits", ptr(w_f32), n, round(n * p), float(p), mode, seed & (2**64 - 1), trial)


def flip_images(inp_u8, k, mode, seed, trial0, n_trials, out_u8, xor_out=0, minmax=None) -> None:
    """xor_out / minmax: by-products for the uint8-input layer 0 (bnn_flip_images_ex): the output bytes
    XOR-ed into int8 operand form, and the per-trial {min, max} ([n_trials, 2] int32) of the corrupted set."""
    R, F = inp_u8.shape
    assert inp_u8.is_contiguous() and out_u8.is_contiguous()
    if xor_out == 0 and minmax is None:
        _call("bnn_flip_images", ptr(inp_u8), R, F, k, mode, seed & (2**64 - 1), trial0, n_trials,
              ptr(out_u8))
        return
    assert minmax is None or (minmax.is_contiguous() and minmax.numel() >= 2 * n_trials)
    _call("bnn_flip_images_ex", ptr(inp_u8), R, F, k, mode, seed & (2**64 - 1), trial0, n_trials,
          ptr(out_u8), xor_out, ptr(minmax))


def u8_input_thresholds(wt_bits, N, K, thr, sgn, minmax, n_trials, new_min, new_max, bias, thr_out,
                        per_trial_weights=False, shared_minmax=False) -> None:
    """Per-trial integer thresholds of normalize + batch-norm + sign for a uint8-fed layer 0.
    wt_bits: [N, words] (image faults) or, with per_trial_weights, [n_trials, N, words] (weight faults);
    minmax: [n_trials, 2], or with shared_minmax one {min, max} pair for every trial."""
    bits_stride = wt_bits.stride(0) if per_trial_weights else 0
    _call("bnn_u8_input_thresholds", ptr(wt_bits), _ld(wt_bits), bits_stride, N, K, ptr(thr), ptr(sgn),
          ptr(minmax), 0 if shared_minmax else 2, n_trials, new_min, new_max, bias, ptr(thr_out), _ld(thr_out))


def minmax_u8(inp_u8, minmax_u32, minmax_f32=None) -> None:
    _call("bnn_minmax_u8", ptr(inp_u8), inp_u8.numel(), ptr(minmax_u32), ptr(minmax_f32))


def normalize_u8(inp_u8, mn, mx, new_min, new_max, out_f32) -> None:
    _call("bnn_normalize_u8", ptr(inp_u8), inp_u8.numel(), ptr(mn), ptr(mx), new_min, new_max,
          ptr(out_f32))
