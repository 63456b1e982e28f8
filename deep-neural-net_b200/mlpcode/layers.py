"""B200-native drop-in for the reference's layer classes (mlpcode/layers.py).

Same class names, constructor arguments, methods and attributes as the reference
(`Layer`, `BatchNormLayer`, `LinearLayer`, `BinaryLayer`; `build`, `forward`, `backwards`,
`load_parameters`, `loadBatchNormParams`, `batchNormParams`, `addCallback`, `clearCallbacks`,
`weights`, `bias`, `cache`, `batchNormLayer`), but every array lives in HBM and every arithmetic
step is a libbnn_b200 kernel:

  forward  (layers.py:283-292, 209-238, 74-99)
    binarize W -> K-major +-1 operands        bnn_binarize_weights_t
    z = X . sign(W)                           bnn_gemm   (+-1 x +-1: tcgen05 kind::i8, exact int32;
                                                          real x +-1: fp16 hi/lo split, kind::f16)
    batch statistics / running statistics     bnn_colstats + bnn_bn_stats_finalize
    BN + sign (or softmax / identity)         bnn_bn_apply (+ bnn_softmax_xent)
  backward (layers.py:247-271, 101-127)
    BN backward, gamma/beta SGD, running stats   bnn_bn_bwd_reduce / _finalize / _apply
    dX = delta' . W^T  (latent, pre-update W)    bnn_gemm (3 fp16 split products, fp32 accumulate)
    W -= lr * X^T . delta'                       bnn_gemm with the fused SGD epilogue

Hidden activations travel between layers as `SignActivation` (int8 +-1 operand + transposed fp16
copy for the next layer's dW) instead of fp32 arrays; `.t` materialises the fp32 view on demand.
Only the configuration the BNN scripts use is implemented on the device: batch-norm on, no bias
(network.py:37-41 disables bias whenever batch-norm is on).
"""
from __future__ import annotations

import contextlib
import os

import numpy as np
import torch

from . import _native as nat
from . import kernels as K_
from .activation import ACTIVATION_DERIVATIVES, ACTIVATION_FUNCTIONS
from .activation import ActivationFuncs as af
from .callbacks import Callback
from .device import DeviceArray, asarray, empty, pad_to, require_cuda, to_tensor, zeros


def binary_gemm_variant() -> str:
    """'i8' (tcgen05 kind::i8, default — chosen from the ncu comparison in profiles/) or 'popc'."""
    v = os.environ.get("BNN_BINARY_GEMM", "i8").lower()
    if v not in ("i8", "popc"):
        raise ValueError(f"BNN_BINARY_GEMM={v!r}: expected 'i8' or 'popc'")
    return v


class SignActivation(DeviceArray):
    """Output of a sign-activated layer: values in {+1,-1} kept in GEMM-operand form."""

    def __init__(self, i8, n, t16=None, bits=None, t8=None, t8x127=None, f4=None):
        self.i8, self.n, self.t16, self.bits = i8, n, t16, bits
        self.t8, self.t8x127 = t8, t8x127   # transposed int8 +-1 / +-127 (int8 dW operand pieces)
        self.f4 = f4                        # E2M1 nibbles [B, pad32(n) / 2] (the consumer's FP4 forward operand)
        self.rows = (i8 if i8 is not None else f4).shape[0]
        self._dense = None

    @property
    def t(self):  # fp32 [B, n] view for generic consumers (not used on the hot path)
        if self._dense is None:
            if self.i8 is not None:
                self._dense = self.i8[:, : self.n].to(torch.float32)
            else:   # nibbles 0x2 / 0xA -> +1 / -1
                b = self.f4
                nib = torch.stack((b & 0xF, b >> 4), dim=-1).reshape(b.shape[0], -1)[:, : self.n]
                self._dense = torch.where(nib == 0x2, 1.0, -1.0).to(torch.float32)
        return self._dense

    @t.setter
    def t(self, v):
        self._dense = v

    @property
    def shape(self):
        return (self.rows, self.n)

    @property
    def ndim(self):
        return 2

    @property
    def size(self):
        return self.rows * self.n

    @property
    def dtype(self):
        return np.dtype(np.float32)

    def __len__(self):
        return self.rows


class SoftmaxOutput(DeviceArray):
    """Softmax probabilities that remember their logits, so loss + gradient fuse into one kernel."""

    def __init__(self, probs, logits):
        super().__init__(probs)
        self.logits = logits
        self.fused = None  # (loss_rows, grad, loss_sum) once mlpcode.loss has run


class _Workspace:
    """Named device buffers, allocated once per (name, shape, dtype) and reused across steps."""

    def __init__(self):
        self._bufs = {}

    def get(self, name, shape, dtype, zero=False):
        key = (name, tuple(shape), dtype)
        b = self._bufs.get(key)
        if b is None:
            b = zeros(shape, dtype) if zero else empty(shape, dtype)
            self._bufs[key] = b
        return b

    def clear(self):
        self._bufs.clear()


class DwOverlap:
    """Side stream for the weight-gradient GEMMs of a single-GPU training step.

    dW_l (layers.py:260,268) is needed by nothing until the next step's forward, while the backward
    pass of the layer below starts with HBM-bound kernels (batch-norm backward, operand splits).  The
    persistent dW GEMM's last wave leaves most SMs idle (512 tiles on 148 SMs), so it is launched on
    this side stream and the layer below's memory-bound kernels fill the hole; Network joins the
    stream at the end of the step.  Works identically under CUDA-graph capture (fork / join edges)."""

    def __init__(self):
        self.side = torch.cuda.Stream()
        self.events = []

    @contextlib.contextmanager
    def branch(self):
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        self.side.wait_event(fork)
        with torch.cuda.stream(self.side):
            yield
            done = torch.cuda.Event()
            done.record(self.side)
        self.events.append(done)

    def join(self):
        main = torch.cuda.current_stream()
        for e in self.events:
            main.wait_event(e)
        self.events.clear()


class _NoComm:
    world = 1

    def begin_step(self):
        pass

    def finish_step(self):
        pass

    def allreduce_sum(self, t):
        return None

    def global_rows(self, local_rows):
        return local_rows

    def allreduce_sum_async(self, t):
        return None

    peer_exchange = None


class Layer:
    def __init__(self, layerUnits: int, gpu=False):
        self.layerUnits = layerUnits
        self.gpu = gpu
        self.cache = {}
        self.isBuilt = False

    def build(self):
        self.isBuilt = True


class BatchNormLayer(Layer):
    EPSILON = 1e-4
    MA_BETA = 0.9  # moving-average coefficient of the running statistics

    def __init__(self, layerUnits: int, gpu=False):
        super().__init__(layerUnits, gpu=gpu)
        self.beta = None
        self.gamma = None
        self.mu = None
        self.sigma = None  # running VARIANCE (layers.py:121-123), despite the name
        self._ws = _Workspace()
        self.comm = _NoComm()
        self.exchange_site = 0  # call-site base of this layer's peer exchanges (2 per layer)

    @property
    def parameter_shape(self):
        return (self.layerUnits,)

    @property
    def batchNormParams(self):
        assert self.isBuilt
        return (self.gamma, self.beta, self.mu, self.sigma)

    def loadBatchNormParams(self, gamma, beta, mu, sigma):
        self.version = getattr(self, "version", 0) + 1
        shp = self.parameter_shape
        for a in (gamma, beta, mu, sigma):
            assert tuple(a.shape) == shp
        # like the reference, the layer adopts the arrays it is given (device copies are made only
        # for host inputs)
        self.gamma = asarray(gamma, np.float32)
        self.beta = asarray(beta, np.float32)
        self.mu = asarray(mu, np.float32)
        self.sigma = asarray(sigma, np.float32)

    def build(self):
        n = self.layerUnits
        if self.beta is None:
            self.beta = DeviceArray(zeros((n,)))
        if self.gamma is None:
            self.gamma = DeviceArray(torch.ones((n,), dtype=torch.float32, device=require_cuda()))
        if self.mu is None:
            self.mu = DeviceArray(zeros((n,)))
        if self.sigma is None:
            self.sigma = DeviceArray(torch.ones((n,), dtype=torch.float32, device=require_cuda()))
        super().build()

    # -- device primitives shared with BinaryLayer ----------------------------------------
    def stats_target(self):
        """Zeroed [2, n] fp64 accumulator for (sum z, sum z^2) that a GEMM epilogue may fill
        (BNN_STATS_COLSUM) instead of a bnn_colstats pass; hand it back to batch_stats()."""
        n = self.layerUnits
        px = self.comm.peer_exchange
        if px is not None and n <= px.max_cols:
            slot, call = px.begin(self.exchange_site, 2 * n)
            acc = slot.view(2, n)
            self._stats_call = call
        else:
            acc = self._ws.get("sums", (2, n), torch.float64)
            self._stats_call = None
        K_.zero(acc)
        return acc

    def batch_stats(self, Z, B, filled=None):
        """Column mean / population variance of Z over the GLOBAL batch (layers.py:81-82).
        filled: the accumulator from stats_target() if the producing GEMM already summed the columns."""
        n = self.layerUnits
        sums = self._ws.get("sums", (2, n), torch.float64)
        mu = self._ws.get("mu_b", (n,), torch.float32)
        var = self._ws.get("var_b", (n,), torch.float32)
        px = self.comm.peer_exchange
        if px is not None and n <= px.max_cols:
            # sync-BN over NVLink peer memory: partial sums land in this rank's symmetric slot and the
            # finalize kernel itself gathers every rank's slot (one launch, no NCCL call)
            if filled is None:
                slot, call = px.begin(self.exchange_site, 2 * n)
                K_.colstats(Z, B, n, slot.view(2, n))
            else:
                call = self._stats_call
            K_.bn_stats_finalize_p2p(call, n, self.comm.global_rows(B), mu, var, sums)
            return sums, mu, var
        if filled is None:
            K_.colstats(Z, B, n, sums)
        self.comm.allreduce_sum(sums)  # sync-BN: exact (integer-valued fp64 sums) for +-1 layers
        K_.bn_stats_finalize(sums, n, self.comm.global_rows(B), mu, var)
        return sums, mu, var

    def bwd_reduce_target(self, Z, mu):
        """Zeroed accumulators for the batch-norm backward reductions of THIS layer, to be filled by
        the epilogue of the dX GEMM that produces its incoming delta (BNN_STATS_BN_BWD)."""
        n = self.layerUnits
        red_max = self._ws.get("red_max", (2, n), torch.int32)
        px = self.comm.peer_exchange
        if px is not None and n <= px.max_cols:
            slot, call = px.begin(self.exchange_site + 1, 2 * n)
            red = slot.view(2, n)
            self._red_call = call
        else:
            red = self._ws.get("red", (2, n), torch.float64)
            self._red_call = None
        K_.zero(red)
        K_.zero(red_max)
        return dict(mode=nat.STATS_BN_BWD, sums=red, max=red_max, Z=Z, mu=mu)

    def backward_kernels(self, delta, Z, B, sums, mu, var, lr, want_f32, d_rm, d_t, scale,
                         reduced=False, digits=None):
        """reduced: the reductions were already accumulated into bwd_reduce_target() by the GEMM that
        produced `delta`.  digits: (q1t, q3t, q2t) int8 [N, pad16(B)] — emit delta' transposed as
        signed fixed-point digits for the int8 dW product instead of the fp16 pieces `d_t`; returns
        the [2, N] per-column scales (row 1 is the GEMM's col_scale)."""
        n = self.layerUnits
        red = self._ws.get("red", (2, n), torch.float64)
        red_max = self._ws.get("red_max", (2, n), torch.int32)
        coef = self._ws.get("coef", (3, n), torch.float32)
        bound = self._ws.get("bound", (1,), torch.float32)  # written as uint32 bits of a float
        px = self.comm.peer_exchange
        if px is not None and n <= px.max_cols:
            if reduced:
                call = self._red_call
            else:
                slot, call = px.begin(self.exchange_site + 1, 2 * n)
                K_.bn_bwd_reduce(delta, Z, B, n, mu, slot.view(2, n), red_max)
            K_.allreduce_f64_p2p(call, 2 * n, red)
            if getattr(self.comm, "dw_exchange", None) is not None:
                self.comm.dw_exchange.start_broadcasts()
        else:
            if not reduced:
                K_.bn_bwd_reduce(delta, Z, B, n, mu, red, red_max)
            self.comm.allreduce_sum(red)
        colq = self._ws.get("col_quant", (2, n), torch.float32) if digits is not None else None
        K_.bn_bwd_finalize(red, red_max, n, self.comm.global_rows(B), mu, var, self.EPSILON,
                           self.gamma.t, self.beta.t, self.mu.t, self.sigma.t, lr, self.MA_BETA,
                           sums, coef, bound, col_quant=colq)
        if digits is not None:
            assert not want_f32
            K_.bn_bwd_apply_q(delta, Z, B, n, coef, bound, colq, digits[0], digits[1], digits[2],
                              d_hi=d_rm[0] if d_rm else None, d_lo=d_rm[1] if d_rm else None,
                              scale_out=scale)
            return colq
        out = empty((B, n)) if want_f32 else None
        K_.bn_bwd_apply(delta, Z, B, n, mu, coef, bound, out_f32=out,
                        d_hi=d_rm[0] if d_rm else None, d_lo=d_rm[1] if d_rm else None,
                        dt_hi=d_t[0] if d_t else None, dt_lo=d_t[1] if d_t else None,
                        scale_out=scale)
        return out

    # -- reference API on fp32 arrays -------------------------------------------------------
    def forward(self, Z, isTrain=True):
        assert self.isBuilt
        self.cache.clear()
        Zt = to_tensor(Z, torch.float32)
        B, n = Zt.shape
        out = empty((B, n))
        if isTrain:
            sums, mu, var = self.batch_stats(Zt, B)
            K_.bn_apply(Zt, B, n, mu, var, self.gamma.t, self.beta.t, self.EPSILON, out_f32=out)
            self.cache.update(input=Zt, mu=mu, var=var, sums=sums, output=out)
        else:
            K_.bn_apply(Zt, B, n, self.mu.t, self.sigma.t, self.gamma.t, self.beta.t, self.EPSILON,
                        out_f32=out)
        return DeviceArray(out)

    def backwards(self, delta, lr: float):
        Zt = self.cache["input"]
        B = Zt.shape[0]
        d = to_tensor(delta, torch.float32)
        out = self.backward_kernels(d, Zt, B, self.cache["sums"], self.cache["mu"],
                                    self.cache["var"], float(lr), True, None, None, None)
        self.cache.clear()
        return DeviceArray(out)


class LinearLayer(Layer):
    def __init__(self, layerUnits: int, inputUnits: int, useBias=False, gpu=False, batchNorm=False):
        super().__init__(layerUnits, gpu=gpu)
        self.inputUnits = inputUnits
        self.useBias = useBias
        self._pending_update = None   # weight update of a data-parallel step still in flight
        self._dwx = None              # (parallel.DwExchange, layer index): fused dW exchange over NVLink
        self._dw_overlap = None       # DwOverlap shared by the layers of a Network (single-GPU runs)
        self.weights = None
        self.bias = None
        self.activation = None
        self.batchNorm = batchNorm
        self.callbacks = []
        self.comm = _NoComm()
        self.skip_input_grad = False  # set on the first layer: its dX is discarded (network.py:389-390)
        # diagnostics: a [K, N] fp32 device tensor placed here receives dW (layers.py:260) of the next
        # backwards() call, formed from the very operands and passes of the update GEMM
        self.capture_dw = None
        self._ws = _Workspace()
        if batchNorm:
            self.batchNormLayer = BatchNormLayer(layerUnits, gpu=gpu)

    @property
    def weights(self):
        """Latent fp32 [K, N] weights (layers.py:191-194).  A data-parallel update still in flight
        (all-reduce / peer rows on their way) is completed before the array is handed out."""
        if self._pending_update is not None:
            self.finish_update()
        if self._dwx is not None and self._dwx[1] in self._dwx[0].pending:
            self._dwx[0].finish_step()  # a step driven layer by layer, outside Network.train_step
        return self._weights

    @weights.setter
    def weights(self, value):
        self._weights = value
        self.version = getattr(self, "version", 0) + 1  # invalidates captured step graphs

    @property
    def weights_shape(self):
        return self.inputUnits, self.layerUnits

    @property
    def bias_shape(self):
        return (self.layerUnits,)

    @property
    def batchNormParams(self):
        assert self.batchNorm
        return self.batchNormLayer.batchNormParams

    def load_parameters(self, weights, bias=None):
        assert tuple(weights.shape) == self.weights_shape
        self.weights = asarray(weights, np.float32)
        if self.useBias:
            raise NotImplementedError("bias is outside the B200 hot path (batch-norm nets drop it, "
                                      "network.py:37-41)")
        self._home_weights()

    def attach_dw_exchange(self, dwx, index: int) -> None:
        """Data-parallel runs keep W in the exchange's symmetric arena so that the owner of a row block
        can store its updated rows into every rank's copy (parallel.DwExchange)."""
        self._dwx = (dwx, index) if dwx is not None else None
        self._home_weights()

    def _home_weights(self) -> None:
        if self._dwx is None or self._weights is None:
            return
        dwx, li = self._dwx
        view = dwx.weight_view(li)
        if self._weights.t.data_ptr() != view.data_ptr():
            assert tuple(self._weights.shape) == tuple(view.shape)
            view.copy_(self._weights.t)
            self.weights = DeviceArray(view)

    def loadBatchNormParams(self, gamma, beta, mu, sigma):
        assert self.batchNorm
        self.batchNormLayer.loadBatchNormParams(gamma, beta, mu, sigma)

    def addCallback(self, callback: Callback):
        self.callbacks.append(callback)

    def clearCallbacks(self):
        self.callbacks.clear()

    def set_comm(self, comm):
        self.version = getattr(self, "version", 0) + 1
        self.comm = comm
        if self.batchNorm:
            self.batchNormLayer.comm = comm

    def build(self, activation: af = None):
        if not self.gpu:
            raise NotImplementedError("this build is the device path only (useGpu=True); run the "
                                      "reference NumPy classes for CPU execution")
        if self.useBias or not self.batchNorm:
            raise NotImplementedError("B200 hot path implements batch-norm layers without bias (the "
                                      "configuration of every BNN script in the reference)")
        dev = require_cuda()
        if self.weights is None:
            # same distribution as layers.py:191-194 (randn * sqrt(1/K)); the stream differs
            w = torch.randn(self.weights_shape, dtype=torch.float32, device=dev)
            self.weights = DeviceArray(w * float(np.sqrt(1.0 / self.inputUnits)))
        self.activation = activation
        if activation is not None:
            assert activation in ACTIVATION_FUNCTIONS
        if self.batchNorm:
            self.batchNormLayer.build()
        super().build()

    # ------------------------------------------------------------------ operand preparation (shared)
    def _real_input_operands(self, Xt, train: bool):
        B, Kd = Xt.shape
        ws = self._ws
        amax = ws.get("x_amax", (1,), torch.float32)
        scale = ws.get("x_scale", (1,), torch.float32)
        hi = ws.get(f"x_hi{B}", (B, pad_to(Kd, 8)), torch.float16, zero=True)
        lo = ws.get(f"x_lo{B}", (B, pad_to(Kd, 8)), torch.float16, zero=True)
        hi_t = lo_t = None
        if train:
            hi_t = ws.get(f"xt_hi{B}", (Kd, pad_to(B, 8)), torch.float16, zero=True)
            lo_t = ws.get(f"xt_lo{B}", (Kd, pad_to(B, 8)), torch.float16, zero=True)
        K_.absmax_f32(Xt, amax)
        K_.split_f16(Xt, amax, hi=hi, lo=lo, hi_t=hi_t, lo_t=lo_t, scale_out=scale)
        return dict(hi=hi, lo=lo, hi_t=hi_t, lo_t=lo_t, scale=scale)

    def finish_update(self) -> None:
        """Apply `W -= lr * dW` of a data-parallel step once its all-reduce has landed."""
        if self._pending_update is None:
            return
        work, dW, lr = self._pending_update
        self._pending_update = None
        if work is not None:
            work.wait()  # stream-side wait: the update kernel queues behind the collective
        K_.sgd_update(self.weights.t, dW, lr)

    def _dw_reduce(self, lr: float, broadcast: bool) -> None:
        """Owner update of a data-parallel step: signal "my partial rows are pushed", wait for the peers', sum,
        update, (store to the peers).  BNN_DW_WAIT=split moves signal + wait into one-block kernels of their own so
        that the update kernel's blocks do not hold the SMs' thread slots while they wait; measured, it makes no
        difference at N = 2 (3.37 vs 3.36 ms) and costs 1 % at N = 8 (3.47 vs 3.43 ms): ranks run in lock step and the
        two extra launches outweigh the shorter occupancy.  Default: one fused kernel."""
        dwx, li = self._dwx
        L = dwx.layout[li]
        if os.environ.get("BNN_DW_WAIT", "fused") == "split":
            K_.peer_signal(dwx.flag_ptrs(), 2 * li, dwx.rank, dwx.world, dwx.comm.step_dev)
            K_.peer_wait(dwx.flag_ptrs()[dwx.rank], 2 * li, 1, 1, dwx.rank, dwx.world, dwx.comm.step_dev)
            signal = False
        else:
            signal = True
        K_.dw_reduce_sgd_p2p(dwx.stage_ptrs(li)[dwx.rank], dwx.w_ptrs(li), dwx.flag_ptrs(), 2 * li,
                             2 * li + 1, dwx.rank, dwx.world, dwx.comm.step_dev, L["K"], L["N"],
                             L["rows_per_owner"], lr, broadcast, dwx.counters[li:li + 1], signal=signal)

    # ------------------------------------------------------------------ forward, real-valued weights
    def forward(self, X, isTrain=True):
        """z = X . W with the LATENT real-valued weights (layers.py:209-238 without BinaryLayer's binarize):
        three fp16-split tensor-core passes (hi.hi + hi.lo + lo.hi, fp32 accumulation re-rounded per chunk) at
        SGEMM accuracy, batch-norm, activation.  Eval-mode weight hooks (layers.py:215-218) receive and return the
        dense fp32 weight — this is where ErrorCallback(p, mode, bnn=False) injects IEEE-754 bit flips
        (callbacks.py:85-123)."""
        assert self.isBuilt
        Kd, N = self.weights_shape
        ws = self._ws
        bn = self.batchNormLayer
        self.cache.clear()
        Xt = X.t if isinstance(X, DeviceArray) else to_tensor(X, torch.float32)   # +-1 activations arrive dense
        Xt = Xt.to(torch.float32)
        assert Xt.ndim == 2 and Xt.shape[1] == Kd, f"expected [B, {Kd}] input, got {tuple(Xt.shape)}"
        B = Xt.shape[0]
        xin = self._real_input_operands(Xt.contiguous(), isTrain)
        W = self.weights
        if not isTrain:
            for cb in self.callbacks:
                W = cb(W, gpu=self.gpu)
        Wt = to_tensor(W, torch.float32).contiguous()
        assert tuple(Wt.shape) == (Kd, N)
        w_amax = ws.get("w_amax", (1,), torch.float32)
        w_scale_f = ws.get("w_scale_fwd", (1,), torch.float32)
        wt_hi = ws.get("wt_hi", (N, pad_to(Kd, 8)), torch.float16, zero=True)
        wt_lo = ws.get("wt_lo", (N, pad_to(Kd, 8)), torch.float16, zero=True)
        K_.absmax_f32(Wt, w_amax)      # also the bound the backward pass's split of W reads
        K_.split_f16(Wt, w_amax, hi_t=wt_hi, lo_t=wt_lo, scale_out=w_scale_f)
        acc = bn.stats_target() if (isTrain and K_.fused_stats_ok(N)) else None
        stats = dict(mode=nat.STATS_COLSUM, sums=acc) if acc is not None else None
        Z = ws.get(f"z_f32_{B}", (B, pad_to(N, 4)), torch.float32)
        K_.gemm(M=B, N=N, K=Kd, A=(xin["hi"], xin["lo"]), B=(wt_hi, wt_lo), lda=xin["hi"].stride(0),
                ldb=wt_hi.stride(0), D=Z, ldd=Z.stride(0), in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32,
                passes=((0, 0), (0, 1), (1, 0)), sa=xin["scale"], sb=w_scale_f, stats=stats)
        if isTrain:
            sums, mu, var = bn.batch_stats(Z, B, filled=acc)
        else:
            sums, mu, var = None, bn.mu.t, bn.sigma.t
        logits = empty((B, N))
        K_.bn_apply(Z, B, N, mu, var, bn.gamma.t, bn.beta.t, bn.EPSILON, out_f32=logits)
        if self.activation == af.softmax:
            probs = empty((B, N))
            K_.softmax_xent(logits, B, N, probs=probs)
            out = SoftmaxOutput(probs, logits)
        elif self.activation in (None, af.identity):
            out = DeviceArray(logits)
        else:
            out = ACTIVATION_FUNCTIONS[self.activation](DeviceArray(logits))
        if isTrain:
            self.cache.update(input=xin, z=Z, a=out, sums=sums, mu=mu, var=var, B=B, real_input=True)
        return out

    # ------------------------------------------------------------------ backward
    def backwards(self, dA, lr: float, activDeriv=True):
        assert "input" in self.cache
        assert "z" in self.cache
        Kd, N = self.weights_shape
        ws = self._ws
        c = self.cache
        B, Z, xin = c["B"], c["z"], c["input"]
        lr = float(lr)
        # STE: hard_tanh_derivative evaluated on the post-sign activations is the identity
        # (activation.py:122-131 via layers.py:252-254), and softmax+CE arrives pre-differentiated.
        # Any other activation: delta = ACTIVATION_DERIVATIVES[act](dA, a) on the cached activations (:251-254).
        if activDeriv and self.activation not in (None, af.sign, af.identity):
            dA = ACTIVATION_DERIVATIVES[self.activation](dA, c["a"])
        delta = to_tensor(dA, torch.float32)
        assert tuple(delta.shape) == (B, N)
        # the layer above may already have summed this layer's batch-norm reductions in the epilogue
        # of the GEMM that produced dA
        reduced = getattr(dA, "_bn_reduced_for", None) is self

        Np, Bp = pad_to(N, 8), pad_to(B, 8)
        d_hi = ws.get(f"d_hi{B}", (B, Np), torch.float16, zero=True)
        d_lo = ws.get(f"d_lo{B}", (B, Np), torch.float16, zero=True)
        d_scale = ws.get("d_scale", (1,), torch.float32)
        need_dx = not self.skip_input_grad
        # dW on the int8 tensor pipe (exact fixed-point sums, 2x the fp16 rate) when the input is a
        # +-1 activation whose producer emitted the transposed int8 pieces; else two fp16 passes
        use_q = ((not c["real_input"]) and getattr(xin, "t8", None) is not None
                 and (xin.t16 is None or K_.int8_dw_ok(N, B)))  # follow what the producer emitted
        digits = dt_hi = dt_lo = colq = None
        if use_q:
            Bq = pad_to(B, 16)
            digits = tuple(ws.get(f"{nm}{B}", (N, Bq), torch.int8, zero=True) for nm in ("q1t", "q3t", "q2t"))
        else:
            dt_hi = ws.get(f"dt_hi{B}", (N, Bp), torch.float16, zero=True)
            dt_lo = ws.get(f"dt_lo{B}", (N, Bp), torch.float16, zero=True)
        colq = self.batchNormLayer.backward_kernels(
            delta, Z, B, c["sums"], c["mu"], c["var"], lr, False, (d_hi, d_lo) if need_dx else None,
            (dt_hi, dt_lo) if not use_q else None, d_scale, reduced=reduced, digits=digits)

        # ---- dX = delta' . W^T with the latent, pre-update weights (layers.py:266)
        dX = None
        bstats = below = None
        if need_dx:
            w_amax = ws.get("w_amax", (1,), torch.float32)
            w_scale = ws.get("w_scale", (1,), torch.float32)
            w_hi = ws.get("w_hi", (Kd, Np), torch.float16, zero=True)
            w_lo = ws.get("w_lo", (Kd, Np), torch.float16, zero=True)
            # W has not changed since this step's forward, whose binarize pass left max|W| in w_amax
            K_.split_f16(self._weights.t, w_amax, hi=w_hi, lo=w_lo, scale_out=w_scale)
            dX = empty((B, pad_to(Kd, 4)))
            # dX is the delta of the layer below: its batch-norm backward needs sum(delta) and
            # sum(delta*(z-mu)) per column, which this GEMM's epilogue accumulates on the way out
            below = getattr(self, "input_layer", None)
            bstats = None
            # (only when the layer below passes dX on unchanged: sign's STE or identity — any other
            # activation multiplies dX by its derivative before its batch-norm backward)
            if (below is not None and isinstance(below, LinearLayer) and "z" in below.cache
                    and below.activation in (None, af.sign, af.identity)
                    and below.cache.get("B") == B and K_.fused_stats_ok(Kd)):
                bstats = below.batchNormLayer.bwd_reduce_target(below.cache["z"], below.cache["mu"])
            K_.gemm(M=B, N=Kd, K=N, A=(d_hi, d_lo), B=(w_hi, w_lo), lda=Np, ldb=Np, D=dX,
                    ldd=dX.stride(0), in_dtype=nat.DT_F16, epilogue=nat.EPI_STORE_F32,
                    passes=((0, 0), (0, 1), (1, 0)), sa=d_scale, sb=w_scale, stats=bstats)
            dX = dX[:, :Kd]

        # ---- dW = X^T . delta'  (layers.py:260) and W -= lr * dW (:268)
        if use_q:
            # dW = 1/s_n * (X^T.q1 + 256 * ((127 X)^T.q3 + X^T.q2)): low / high int32 accumulators
            dw = dict(in_dtype=nat.DT_I8, A=(xin.t8, xin.t8x127), B=digits, lda=xin.t8.stride(0),
                      ldb=digits[0].stride(0), passes=((0, 0), (1, 1), (0, 2)), hi_from_pass=1,
                      col_scale=colq[1], sa=None, sb=None)
        elif c["real_input"]:
            dw = dict(in_dtype=nat.DT_F16, A=(xin["hi_t"], xin["lo_t"]), B=(dt_hi, dt_lo),
                      lda=xin["hi_t"].stride(0), ldb=Bp, passes=((0, 0), (0, 1), (1, 0)),
                      sa=xin["scale"], sb=d_scale)
        else:
            dw = dict(in_dtype=nat.DT_F16, A=(xin.t16,), B=(dt_hi, dt_lo), lda=xin.t16.stride(0), ldb=Bp,
                      passes=((0, 0), (0, 1)), sa=None, sb=d_scale)
        Wt = self._weights.t
        if self.capture_dw is not None:
            cap = self.capture_dw
            assert tuple(cap.shape) == (Kd, N) and cap.dtype == torch.float32 and cap.is_contiguous()
            K_.gemm(M=Kd, N=N, K=B, D=cap, ldd=N, epilogue=nat.EPI_STORE_F32, **dw)
        split = 4 if (N <= 32 and Kd >= 1024 and B % 256 == 0 and self.comm.world == 1) else 1
        # single GPU: the dW GEMM goes to the side stream (see DwOverlap) unless it is the last kernel
        # of the step (first layer), where there is nothing left to overlap it with
        ov = self._dw_overlap if (self.comm.world == 1 and not self.skip_input_grad) else None
        with (ov.branch() if ov is not None else contextlib.nullcontext()):
            if split > 1:
                # narrow layer (the 10-unit output): a 128 x 32 tile grid has only Kd/128 CTAs, so the
                # contraction over the batch is cut into `split` slices run as GEMM batches (operand
                # slices are column ranges: batch stride = slice length), summed by the update kernel
                Ks = B // split
                parts = ws.get("dW_parts", (split, Kd, N), torch.float32)
                K_.gemm(M=Kd, N=N, K=Ks, D=parts, ldd=N, epilogue=nat.EPI_STORE_F32, batch=split,
                        stride_a=Ks, stride_b=Ks, stride_d=Kd * N, **dw)
                K_.sgd_update(Wt, parts, lr, parts=split)
            elif self.comm.world == 1:
                K_.gemm(M=Kd, N=N, K=B, D=Wt, ldd=N, epilogue=nat.EPI_SGD_F32, host_scale=lr, **dw)
        if split > 1 or self.comm.world == 1:
            pass  # done above (possibly on the side stream)
        elif self._dwx is not None:
            # reduce-scatter fused into the GEMM epilogue: each 128-row tile of this rank's partial dW
            # is stored over NVLink into the staging buffer of the rank that owns those rows
            self._home_weights()
            dwx, li = self._dwx
            dwx.comm.collectives += 1
            # Like the single-GPU step, the whole weight-gradient chain — dW GEMM, the owner's sum + update (which
            # first waits for the peers' partial rows) — runs on the side stream: nothing on this stream needs
            # W_l before the next step, and the 40 us owner update of a 4096 x 4096 layer plus the wait for the
            # slowest peer's GEMM no longer sit between this layer's dX and the next batch-norm exchange.
            ovd = self._dw_overlap if not self.skip_input_grad else None
            with (ovd.branch() if ovd is not None else contextlib.nullcontext()):
                K_.gemm(M=Kd, N=N, K=B, D=None, ldd=N, epilogue=nat.EPI_SCATTER_F32,
                        scatter=dict(world=dwx.world, rank=dwx.rank,
                                     rows_per_owner=dwx.layout[li]["rows_per_owner"],
                                     stage=dwx.stage_ptrs(li)), **dw)
                if dwx.copy_engine_broadcast and not self.skip_input_grad:
                    # owner update now (local rows only; the kernel also signals "pushed" to the peers);
                    # the rows then travel to the peers on the copy engines, on their own stream, while this
                    # stream carries on with the backward pass.  The copies are started only after the
                    # next batch-norm exchange (dwx.start_broadcasts): queued right here, 10s of MB of
                    # copy traffic would sit in front of that exchange's flags on the NVLink ports.
                    self._dw_reduce(lr, broadcast=False)
                    dwx.reduced[li].record(torch.cuda.current_stream())
                    dwx.deferred.append(li)
                else:
                    # first layer (its dW GEMM ends the step and its weights open the next one, so there
                    # is nothing left to overlap copies with) or BNN_DW_BCAST=kernel: the owner-update
                    # kernel stores the row blocks into the peers itself
                    self._dw_reduce(lr, broadcast=True)
            dwx.pending.append(li)  # awaited by comm.finish_step() at the end of the training step
        else:
            dW = ws.get("dW", (Kd, N), torch.float32)
            K_.gemm(M=Kd, N=N, K=B, D=dW, ldd=N, epilogue=nat.EPI_STORE_F32, **dw)
            # batch-sharded dW -> global dW over NCCL/NVLink, started now and awaited only when the
            # weights are next needed, so it overlaps the lower layers' backward kernels
            self._pending_update = (self.comm.allreduce_sum_async(dW), dW, lr)

        self.cache.clear()
        if dX is None:
            return None
        out = DeviceArray(dX)
        if bstats is not None:
            out._bn_reduced_for = below
        return out


class BinaryLayer(LinearLayer):
    # ------------------------------------------------------------------ reference-style helper
    @staticmethod
    def binarize(x):
        """+1 where x >= 0 else -1, as a new fp32 device array (layers.py:275-281)."""
        xt = to_tensor(x, torch.float32)
        two_d = xt if xt.ndim == 2 else xt.reshape(1, -1)
        R, Cn = two_d.shape
        bits = zeros((R, pad_to(Cn, 32) // 32), torch.int32)
        K_.pack_sign_rows(two_d, bits)
        i8 = empty((R, pad_to(Cn, 16)), torch.int8)
        K_.unpack_bits(bits, Cn, out_i8=i8)
        return DeviceArray(i8[:, :Cn].to(torch.float32).reshape(xt.shape))

    # ------------------------------------------------------------------ operand preparation
    def _weight_operands(self, real_input: bool, f4: bool = False):
        Kd, N = self.weights_shape
        ws = self._ws
        ops = {}
        ops["bits"] = ws.get("wt_bits", (N, pad_to(Kd, 32) // 32), torch.int32, zero=True)
        if real_input:
            ops["f16"] = ws.get("wt_f16", (N, pad_to(Kd, 8)), torch.float16, zero=True)
        elif f4:   # the input arrives as E2M1 nibbles: so do the weights
            ops["f4"] = ws.get("wt_f4", (N, pad_to(Kd, 32) // 2), torch.uint8, zero=True)
        else:
            ops["i8"] = ws.get("wt_i8", (N, pad_to(Kd, 16)), torch.int8, zero=True)
        # the same pass over W yields max|W|, the bound of the backward pass's fp16 split of W
        ops["amax"] = ws.get("w_amax", (1,), torch.float32)
        K_.binarize_weights_t(self.weights.t, wt_i8=ops.get("i8"), wt_f16=ops.get("f16"),
                              wt_bits=ops["bits"], absmax=ops["amax"], wt_f4=ops.get("f4"))
        return ops

    def _apply_foreign_callback(self, cb, wops):
        """Any callable with the reference signature `cb(weight, gpu=)` (layers.py:218): it is
        handed the binarized weight as an fp32 [K, N] device array and its result is re-packed."""
        Kd, N = self.weights_shape
        dense = empty((N, pad_to(Kd, 16)), torch.int8)
        K_.unpack_bits(wops["bits"], Kd, out_i8=dense)
        w_kn = DeviceArray(dense[:, :Kd].t().contiguous().to(torch.float32))
        new_w = to_tensor(cb(w_kn, gpu=self.gpu), torch.float32)
        assert tuple(new_w.shape) == (Kd, N)
        K_.binarize_weights_t(new_w, wt_i8=wops.get("i8"), wt_f16=wops.get("f16"),
                              wt_bits=wops["bits"])

    # ------------------------------------------------------------------ uint8-fed layer 0 (eval)
    def _u8_input_ok(self, X) -> bool:
        """`utils.normalize(uint8 images)` arriving at a sign-activated layer: consumed as bytes."""
        from .utils import NormalizedU8
        Kd, _N = self.weights_shape
        return (isinstance(X, NormalizedU8) and self.activation == af.sign and X.u8.ndim == 2
                and X.u8.shape[1] == Kd and Kd % 16 == 0 and X.new_max > X.new_min
                and X.u8.data_ptr() % 16 == 0)

    def _forward_u8(self, X):
        """Evaluation-mode forward on raw uint8 input (imageBitflipTrainModelSel.py:26-28): normalize
        (utils.py:178-187), X.dot(sign(W)) (layers.py:220), running-statistic batch-norm (:93-97) and sign
        as ONE exact u8 x s8 GEMM whose epilogue compares the integer sums with per-unit thresholds
        (bnn_u8_input_thresholds).  Weight-fault callbacks act on the packed copies as usual."""
        Kd, N = self.weights_shape
        ws, bn = self._ws, self.batchNormLayer
        B = X.u8.shape[0]
        wops = self._weight_operands(False)
        for cb in self.callbacks:
            if hasattr(cb, "apply_to_layer"):
                cb.apply_to_layer(self, wops)
            else:
                self._apply_foreign_callback(cb, wops)
        thr = ws.get("thr", (N,), torch.float32)
        sgn = ws.get("sgn", (N,), torch.float32)
        thr_u8 = ws.get("thr_u8", (1, pad_to(N, 4)), torch.float32)
        K_.bn_sign_thresholds(bn.mu.t, bn.sigma.t, bn.gamma.t, bn.beta.t, bn.EPSILON, N, thr, sgn)
        K_.u8_input_thresholds(wops["bits"], N, Kd, thr, sgn, X.mm_u32, 1, X.new_min, X.new_max, 0, thr_u8,
                               shared_minmax=True)
        if self._consumer_takes_f4():
            act_f4 = ws.get(f"a_f4_{B}", (B, pad_to(N, 32) // 2), torch.uint8, zero=True)
            K_.gemm(M=B, N=N, K=Kd, A=(X.u8,), B=(wops["i8"],), lda=X.u8.stride(0), ldb=wops["i8"].stride(0),
                    D=act_f4, ldd=2 * act_f4.stride(0), in_dtype=nat.DT_I8, epilogue=nat.EPI_SIGN_F4,
                    epi_thr=thr_u8, epi_sgn=sgn, a_unsigned=True)
            return SignActivation(None, N, f4=act_f4)
        act_i8 = ws.get(f"a_i8_{B}", (B, pad_to(N, 16)), torch.int8, zero=True)
        K_.gemm(M=B, N=N, K=Kd, A=(X.u8,), B=(wops["i8"],), lda=X.u8.stride(0), ldb=wops["i8"].stride(0),
                D=act_i8, ldd=act_i8.stride(0), in_dtype=nat.DT_I8, epilogue=nat.EPI_SIGN_I8,
                epi_thr=thr_u8, epi_sgn=sgn, a_unsigned=True)
        act_bits = None
        if binary_gemm_variant() == "popc":
            act_bits = ws.get(f"a_bits{B}", (B, pad_to(N, 32) // 32), torch.int32, zero=True)
            K_.pack_i8_rows(act_i8, N, act_bits)
        return SignActivation(act_i8, N, bits=act_bits)

    def _consumer_takes_f4(self) -> bool:
        """Eval mode: the layer above multiplies this layer's +-1 output on the FP4 tensor pipe (kernels.f4_ok), so
        the sign epilogue writes E2M1 nibbles instead of int8 — half the bytes written here and read there."""
        consumer = getattr(self, "output_layer", None)
        return (isinstance(consumer, BinaryLayer) and binary_gemm_variant() != "popc"
                and K_.f4_ok(self.layerUnits, consumer.layerUnits))

    # ------------------------------------------------------------------ forward
    def forward(self, X, isTrain=True):
        assert self.isBuilt
        Kd, N = self.weights_shape
        ws = self._ws
        bn = self.batchNormLayer
        self.cache.clear()

        if not isTrain and self._u8_input_ok(X):
            return self._forward_u8(X)
        real_input = not isinstance(X, SignActivation)
        if real_input:
            Xt = to_tensor(X, torch.float32)
            assert Xt.ndim == 2 and Xt.shape[1] == Kd, f"expected [B, {Kd}] input, got {tuple(Xt.shape)}"
            B = Xt.shape[0]
        else:
            assert X.n == Kd
            B = X.rows
        f4_in = (not real_input) and X.f4 is not None and X.i8 is None   # the producer emitted nibbles

        # operand preparation of a real-valued input does not depend on W: queue it before the weight
        # read, which in data-parallel runs first waits for the peers' updated rows to land
        xin = self._real_input_operands(Xt, isTrain) if real_input else X
        wops = self._weight_operands(real_input, f4=f4_in)
        # fault-injection hooks act on the already-binarized weight, eval mode only (layers.py:215-218)
        if not isTrain:
            for cb in self.callbacks:
                if hasattr(cb, "apply_to_layer"):
                    cb.apply_to_layer(self, wops)
                else:
                    self._apply_foreign_callback(cb, wops)
            if f4_in and self.callbacks:   # the packed words are what the hooks mutate (callbacks.py:79-83)
                K_.bits_to_f4(wops["bits"], Kd, wops["f4"])

        sign_out = self.activation == af.sign
        popc = binary_gemm_variant() == "popc"
        # eval mode: running statistics are known before the GEMM, so batch-norm + sign collapse into
        # per-column thresholds applied in the GEMM epilogue and Z never reaches HBM
        fuse_sign = sign_out and not isTrain and not (popc and not real_input)
        thr = ws.get("thr", (N,), torch.float32)
        sgn = ws.get("sgn", (N,), torch.float32)
        act_i8 = act_f4 = None
        emit_f4 = fuse_sign and self._consumer_takes_f4()
        consumer = getattr(self, "output_layer", None)
        # training: a wide consumer multiplies on the FP4 pipe too — this layer then hands over packed sign
        # bits (4 MB instead of 33 MB of int8 for 8192 x 4096) which bnn_bits_to_f4 expands to nibbles
        train_f4 = (sign_out and isTrain and not popc and isinstance(consumer, BinaryLayer)
                    and K_.f4_train_ok(N, consumer.layerUnits))
        if emit_f4 or train_f4:
            act_f4 = ws.get(f"a_f4_{B}", (B, pad_to(N, 32) // 2), torch.uint8, zero=True)
        elif sign_out:
            act_i8 = ws.get(f"a_i8_{B}", (B, pad_to(N, 16)), torch.int8, zero=True)
        if fuse_sign:
            K_.bn_sign_thresholds(bn.mu.t, bn.sigma.t, bn.gamma.t, bn.beta.t, bn.EPSILON, N, thr, sgn)

        # training: the GEMM epilogue also accumulates the batch-norm column sums (sum z, sum z^2), so
        # Z is not re-read for them (shapes the epilogue cannot cover use the bnn_colstats pass)
        acc = None
        if isTrain and not (popc and not real_input) and K_.fused_stats_ok(N, f4=f4_in):
            acc = bn.stats_target()
        stats = dict(mode=nat.STATS_COLSUM, sums=acc) if acc is not None else None

        # ---- z = X . sign(W)
        Z = None
        if real_input:
            if emit_f4:
                D, ldd, epi = act_f4, 2 * act_f4.stride(0), nat.EPI_SIGN_F4
            elif fuse_sign:
                D, ldd, epi = act_i8, act_i8.stride(0), nat.EPI_SIGN_I8
            else:
                Z = ws.get(f"z_f32_{B}", (B, pad_to(N, 4)), torch.float32)
                D, ldd, epi = Z, Z.stride(0), nat.EPI_STORE_F32
            K_.gemm(M=B, N=N, K=Kd, A=(xin["hi"], xin["lo"]), B=(wops["f16"],),
                    lda=xin["hi"].stride(0), ldb=wops["f16"].stride(0), D=D, ldd=ldd,
                    in_dtype=nat.DT_F16, epilogue=epi, passes=((0, 0), (1, 0)), sa=xin["scale"],
                    epi_thr=thr, epi_sgn=sgn, stats=stats)
        else:
            s16 = Kd < 32768
            if f4_in:   # E2M1 operands: K and pitches in elements (two per byte)
                pm1 = dict(A=(X.f4,), B=(wops["f4"],), lda=2 * X.f4.stride(0), ldb=2 * wops["f4"].stride(0),
                           in_dtype=nat.DT_F4, K=pad_to(Kd, 32))
            else:
                pm1 = dict(A=(X.i8,), B=(wops["i8"],), lda=X.i8.stride(0), ldb=wops["i8"].stride(0),
                           in_dtype=nat.DT_I8, K=Kd)
            if emit_f4:
                K_.gemm(M=B, N=N, D=act_f4, ldd=2 * act_f4.stride(0), epilogue=nat.EPI_SIGN_F4, epi_thr=thr,
                        epi_sgn=sgn, **pm1)
            elif fuse_sign:
                K_.gemm(M=B, N=N, D=act_i8, ldd=act_i8.stride(0), epilogue=nat.EPI_SIGN_I8, epi_thr=thr,
                        epi_sgn=sgn, **pm1)
            elif f4_in:   # also the training forward of wide layers: column sums fused on 128-wide tiles
                Z = ws.get(f"z_int_{B}", (B, pad_to(N, 8)), torch.int16 if s16 else torch.int32)
                K_.gemm(M=B, N=N, D=Z, ldd=Z.stride(0), epilogue=nat.EPI_STORE_S16 if s16 else nat.EPI_STORE_S32,
                        stats=stats, **pm1)
            else:
                Z = ws.get(f"z_int_{B}", (B, pad_to(N, 8)), torch.int16 if s16 else torch.int32)
                if popc:
                    if X.bits is None:
                        X.bits = ws.get(f"x_bits{B}", (B, pad_to(Kd, 32) // 32), torch.int32, zero=True)
                        K_.pack_i8_rows(X.i8, Kd, X.bits)
                    K_.gemm_popc(X.bits, wops["bits"], B, N, Kd, Z, out_s16=s16)
                else:
                    K_.gemm(M=B, N=N, K=Kd, A=(X.i8,), B=(wops["i8"],), lda=X.i8.stride(0),
                            ldb=wops["i8"].stride(0), D=Z, ldd=Z.stride(0), in_dtype=nat.DT_I8,
                            epilogue=nat.EPI_STORE_S16 if s16 else nat.EPI_STORE_S32, stats=stats)

        # ---- batch-norm statistics
        if isTrain:
            sums, mu, var = bn.batch_stats(Z, B, filled=acc)
        else:
            sums, mu, var = None, bn.mu.t, bn.sigma.t

        # ---- BN + activation
        if sign_out:
            act_t16 = act_t8 = act_t8x127 = None
            if isTrain and consumer is not None and K_.int8_dw_ok(consumer.layerUnits, B):
                # the layer above forms its dW on the int8 tensor pipe: transposed +-1 / +-127 bytes
                act_t8 = ws.get(f"a_t8_{B}", (N, pad_to(B, 16)), torch.int8, zero=True)
                act_t8x127 = ws.get(f"a_t8x127_{B}", (N, pad_to(B, 16)), torch.int8, zero=True)
            elif isTrain:
                act_t16 = ws.get(f"a_t16_{B}", (N, pad_to(B, 8)), torch.float16, zero=True)
            act_bits = ws.get(f"a_bits{B}", (B, pad_to(N, 32) // 32), torch.int32, zero=True) if popc else None
            if not fuse_sign:
                # sign(BN(z)) through the per-column thresholds: one compare per element, bit-identical
                # to evaluating gamma*((z-mu)/sqrt(var+eps))+beta and taking the sign
                K_.bn_sign_thresholds(mu, var, bn.gamma.t, bn.beta.t, bn.EPSILON, N, thr, sgn)
                K_.bn_sign(Z, B, N, thr, sgn, act_i8=act_i8, act_t16=act_t16, act_bits=act_bits,
                           act_t8=act_t8, act_t8x127=act_t8x127, act_f4=act_f4 if train_f4 else None)
            elif act_bits is not None:
                K_.pack_i8_rows(act_i8, N, act_bits)
            out = SignActivation(act_i8, N, t16=act_t16, bits=act_bits, t8=act_t8, t8x127=act_t8x127, f4=act_f4)
        else:
            logits = empty((B, N))
            K_.bn_apply(Z, B, N, mu, var, bn.gamma.t, bn.beta.t, bn.EPSILON, out_f32=logits)
            if self.activation == af.softmax:
                probs = empty((B, N))
                K_.softmax_xent(logits, B, N, probs=probs)
                out = SoftmaxOutput(probs, logits)
            elif self.activation in (None, af.identity):
                out = DeviceArray(logits)
            else:  # sigmoid / tanh / relu / leaky_relu on the batch-norm output (layers.py:230-233)
                out = ACTIVATION_FUNCTIONS[self.activation](DeviceArray(logits))

        if isTrain:
            self.cache.update(input=xin, z=Z, a=out, sums=sums, mu=mu, var=var, B=B,
                              real_input=real_input)
        return out
