"""Fault-injection callbacks of the B200 path — interface of mlpcode/callbacks.py.

`ErrorCallback(p, mode=2, bnn=True)` and `ImageErrorCallback(p, mode=2)` keep the reference's
constructor and call signatures, but the flips are drawn and applied on the device:

  * weights (callbacks.py:79-83, dispatch :94-98): i.i.d. Bernoulli(p) sign flips, `mode` ignored for
    BNNs exactly as in the reference.  The mask comes from Philox4x32-10 keyed by (seed, call
    counter, layer id) and is XOR-ed into the packed K-major binarized weights by
    bnn_flip_weights_bernoulli — a fresh mask per layer per forward call, like the reference's
    fresh `np.random.uniform` draw.
  * images (callbacks.py:130-147): exactly round(p * 8F) distinct bit positions per uint8 row
    (mode 2: any bit; 0: among 0-bits; 1: among 1-bits), chosen by a keyed Feistel bijection in
    bnn_flip_images.

The reference draws from NumPy's global MT19937 stream, which no counter-based device RNG can
replay.  For bit-exact parity against the reference's own draw, pass `mask_source=` (weights): a
callable `shape -> bool ndarray` that is applied through bnn_flip_weights_mask.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K_
from .device import DeviceArray, empty, to_tensor, zeros

_SEED_COUNTER = [0x5EED0000]


def _next_seed() -> int:
    _SEED_COUNTER[0] += 1
    return _SEED_COUNTER[0]


class Callback:
    def __init__(self, p: float, mode: int):
        assert mode in (0, 1, 2)  # 0: 0 -> 1 only, 1: 1 -> 0 only, 2: both
        self.p = 0 if p is None else p
        self.mode = mode


class ErrorCallback(Callback):
    """Weight sign-flip faults for binarized layers."""

    def __init__(self, p: float, mode=2, bnn=False, seed: int | None = None, mask_source=None):
        super().__init__(p, mode)
        self.forBnn = bnn
        self.seed = _next_seed() if seed is None else int(seed)
        self.calls = 0  # trial counter: every application draws a fresh, reproducible mask
        self.mask_source = mask_source
        self.last_mask_bits = None

    # ---- fast path used by BinaryLayer.forward (eval mode) --------------------------------
    def apply_to_layer(self, layer, wops) -> None:
        if not self.forBnn:
            raise NotImplementedError(
                "IEEE-754 bit flips (callbacks.py:35-59, 85-123) turn +-1 weights into arbitrary reals; the "
                "binarized layer's integer forward cannot consume them.  The flips themselves are available "
                "on dense arrays: ErrorCallback(p, mode, bnn=False)(array)")
        Kd, N = layer.weights_shape
        trial, self.calls = self.calls, self.calls + 1
        i8, f16, bits = wops.get("i8"), wops.get("f16"), wops["bits"]
        if self.mask_source is not None:
            mask = np.ascontiguousarray(self.mask_source((Kd, N)), dtype=np.uint8)
            assert mask.shape == (Kd, N)
            K_.flip_weights_mask(to_tensor(mask), Kd, N, io_i8=i8, io_f16=f16, io_bits=bits)
            return
        if self.p <= 0:
            return
        # in place: each thread reads its 32-element word before writing it
        K_.flip_weights_bernoulli(N=N, K=Kd, p=float(self.p), seed=self.seed,
                                  stream_id=id_of_layer(layer), trial0=trial, n_trials=1,
                                  wt_bits=bits, out_i8=None if i8 is None else i8.unsqueeze(0),
                                  out_f16=None if f16 is None else f16.unsqueeze(0),
                                  out_bits=bits.unsqueeze(0))

    # ---- reference call signature on a dense +-1 weight array -----------------------------
    def __call__(self, inp, gpu=False):
        assert inp.ndim == 2
        if not self.forBnn:
            # IEEE-754 bit flips of a real-valued array (callbacks.py:100-123): round(size*p) elements,
            # round(candidates*p) bits in each.  The input is left untouched, like the reference's
            # flatten() copy; with no element selected the input itself is returned (callbacks.py:111-112).
            w = to_tensor(inp, torch.float32)
            if round(w.numel() * self.p) == 0:
                return inp
            out = w.clone()
            trial, self.calls = self.calls, self.calls + 1
            K_.flip_f32_bits(out, float(self.p), self.mode, self.seed, trial)
            return DeviceArray(out)
        w = to_tensor(inp, torch.float32)
        Kd, N = w.shape
        words = (Kd + 31) // 32
        bits = zeros((N, words), torch.int32)
        K_.binarize_weights_t(w, wt_bits=bits)
        trial, self.calls = self.calls, self.calls + 1
        if self.mask_source is not None:
            mask = np.ascontiguousarray(self.mask_source((Kd, N)), dtype=np.uint8)
            K_.flip_weights_mask(to_tensor(mask), Kd, N, io_bits=bits)
        elif self.p > 0:
            K_.flip_weights_bernoulli(N=N, K=Kd, p=float(self.p), seed=self.seed, stream_id=0,
                                      trial0=trial, n_trials=1, wt_bits=bits,
                                      out_bits=bits.unsqueeze(0))
        i8 = empty((N, (Kd + 15) // 16 * 16), torch.int8)
        K_.unpack_bits(bits, Kd, out_i8=i8)
        return DeviceArray(i8[:, :Kd].t().contiguous().to(torch.float32))


def id_of_layer(layer) -> int:
    """Philox stream id of a layer: distinct layers must not share a mask stream."""
    lid = getattr(layer, "layer_index", None)
    return int(lid) if lid is not None else (id(layer) >> 4) & 0x7FFFFFFF


class ImageErrorCallback(Callback):
    """Exact-count bit flips in uint8 images (one row = one image)."""

    def __init__(self, p: float, mode=2, seed: int | None = None):
        super().__init__(p, mode)
        self.nbits = self.p * 8
        self.seed = _next_seed() if seed is None else int(seed)
        self.calls = 0

    def flips_per_row(self, n_features: int) -> int:
        # round(p * unpacked.size) with Python's round-half-to-even (callbacks.py:135)
        return int(round(self.p * n_features * 8))

    def __call__(self, arr, gpu=False):
        if self.p is None or self.p == 0:
            return arr
        t = to_tensor(arr)
        if t.dtype != torch.uint8:
            raise ValueError("Only uint8 allowed")
        return DeviceArray(self.flip_trials(t, 1)[0])

    def flip_trials(self, images_u8: torch.Tensor, n_trials: int) -> torch.Tensor:
        """`n_trials` independently corrupted copies [n_trials, R, F] in one launch."""
        R, F = images_u8.shape
        out = empty((n_trials, R, F), torch.uint8)
        K_.flip_images(images_u8, self.flips_per_row(F), self.mode, self.seed, self.calls, n_trials,
                       out)
        self.calls += n_trials
        return out
