"""Loss registry of the B200 path — the interface of mlpcode/loss.py:134-156.

Cross-entropy on a softmax output is what the BNN scripts use (test.py:44).  When the prediction is the
`SoftmaxOutput` produced by the last BinaryLayer, loss rows, their sum and the softmax-fused gradient (p - y)/B
(loss.py:58-66) come from ONE bnn_softmax_xent launch on the logits; the two registry functions then just hand out
the cached pieces.  Every other combination the reference allows — cross-entropy on any other output (with the
fused `ypred - y` gradient for sigmoid, network.py:380-385, or the general -(y/p - (1-y)/(1-p)) form), mse, hinge —
runs through bnn_loss_rows / bnn_loss_grad, which follow loss.py's operation order including the in-place clip.
"""
from enum import Enum

import numpy as np
import torch

from . import _native as nat
from . import kernels as K_
from .device import DeviceArray, empty, to_tensor, zeros


class LossFuncs(Enum):
    cross_entropy = "cross_entropy"
    mse = "mse"
    hinge = "hinge"

    def __repr__(self):
        return self.value

    def __str__(self):
        return self.value


GLOBAL_BATCH_ROWS = None  # data-parallel runs set this: loss.py:66 divides by the GLOBAL batch
                          # when the batch is sharded over ranks (None: the local batch)


def _fused(ypred, y):
    """Run softmax + clipped CE + gradient once per SoftmaxOutput and cache the result on it."""
    if ypred.fused is None:
        logits = ypred.logits
        B, C = logits.shape
        yt = to_tensor(y, torch.int8)
        assert tuple(yt.shape) == (B, C), "labels must be one-hot [B, C]"
        loss_rows, grad = empty((B,)), empty((B, C))
        loss_sum = zeros((1,), torch.float64)
        # probabilities are rewritten clipped, which is the in-place np.clip of loss.py:32
        K_.softmax_xent(logits, B, C, onehot=yt, probs=ypred.t, loss_rows=loss_rows, grad=grad,
                        count=GLOBAL_BATCH_ROWS or B, loss_sum=loss_sum)
        ypred.fused = (loss_rows, grad, loss_sum)
    return ypred.fused


class _LossRows(DeviceArray):
    """Per-sample losses whose mean is already reduced on the device."""

    def __init__(self, rows, loss_sum):
        super().__init__(rows)
        self._sum = loss_sum

    def mean(self, axis=None):
        if axis is not None:
            return super().mean(axis)
        return DeviceArray((self._sum / self.t.shape[0]).to(torch.float32).reshape(()))


def _pred_and_labels(ypred, y):
    """(contiguous fp32 [B, C] view of the prediction that in-place clipping may write through, int8 labels)."""
    pt = ypred.t if isinstance(ypred, DeviceArray) else to_tensor(ypred, torch.float32)
    assert pt.dtype == torch.float32 and pt.is_contiguous(), "predictions must be a contiguous fp32 device array"
    yt = to_tensor(y).to(torch.int8).contiguous()
    assert tuple(yt.shape) == tuple(pt.shape), "labels must have the shape of the predictions [B, C]"
    return pt, yt


def _rows(ypred, y, kind):
    pt, yt = _pred_and_labels(ypred, y)
    rows, loss_sum = empty((pt.shape[0],)), zeros((1,), torch.float64)
    K_.loss_rows(pt, yt, kind, rows=rows, loss_sum=loss_sum)
    return _LossRows(rows, loss_sum)


def _grad(ypred, y, kind):
    pt, yt = _pred_and_labels(ypred, y)
    out = empty(tuple(pt.shape))
    K_.loss_grad(pt, yt, kind, GLOBAL_BATCH_ROWS or pt.shape[0], out)
    return DeviceArray(out)


def cross_entropy_loss(ypred, y, eps=1e-15):
    """loss.py:10-36.  Clips the prediction in place, like the reference (:32)."""
    from .layers import SoftmaxOutput
    if isinstance(ypred, SoftmaxOutput):
        rows, _g, loss_sum = _fused(ypred, y)
        return _LossRows(rows, loss_sum)
    return _rows(ypred, y, nat.LOSS_CROSS_ENTROPY)


def cross_entropy_derivative(ypred, y, eps=1e-15, with_softmax=False):
    """loss.py:39-68: `ypred - y` when the output activation's derivative is fused in (softmax or sigmoid,
    network.py:380-385), else -(y/p - (1-y)/(1-p)) on the in-place-clipped prediction; divided by the batch."""
    from .layers import SoftmaxOutput
    if with_softmax and isinstance(ypred, SoftmaxOutput):
        return DeviceArray(_fused(ypred, y)[1])
    if with_softmax:   # (p - y) / B is exactly mse_derivative's arithmetic (loss.py:58, 66 vs :107)
        return _grad(ypred, y, nat.LOSS_MSE)
    return _grad(ypred, y, nat.LOSS_CROSS_ENTROPY)


def mse(ypred, y):
    """0.5 * sum_c (ypred - y)^2 per row (loss.py:71-89)."""
    return _rows(ypred, y, nat.LOSS_MSE)


def mse_derivative(ypred, y, **kwargs):
    """(ypred - y) / B (loss.py:92-107)."""
    return _grad(ypred, y, nat.LOSS_MSE)


def hinge_loss(ypred, y):
    """sum_c max(0, 1 - ypred * y) per row, labels in {-1, +1} (loss.py:112-119; network.py:301-303)."""
    return _rows(ypred, y, nat.LOSS_HINGE)


def hinge_derivative(ypred, y, **kwargs):
    """where(ypred * y < 1, -y / B, 0) (loss.py:122-131; the reference's result is float64, this one fp32)."""
    return _grad(ypred, y, nat.LOSS_HINGE)


LOSS_FUNCS = {
    LossFuncs.cross_entropy: cross_entropy_loss,
    LossFuncs.mse: mse,
    LossFuncs.hinge: hinge_loss,
}

LOSS_DERIVATES = {
    LossFuncs.cross_entropy: cross_entropy_derivative,
    LossFuncs.mse: mse_derivative,
    LossFuncs.hinge: hinge_derivative,
}
