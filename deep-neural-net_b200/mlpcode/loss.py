"""Loss registry of the B200 path — the interface of mlpcode/loss.py:134-156.

Cross-entropy on a softmax output is the only loss the BNN scripts use (test.py:44).  When the
prediction is the `SoftmaxOutput` produced by the last BinaryLayer, loss rows, their sum and the
softmax-fused gradient (p - y)/B (loss.py:58-66) come from ONE bnn_softmax_xent launch on the
logits; the two registry functions then just hand out the cached pieces.
"""
from enum import Enum

import numpy as np
import torch

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


def cross_entropy_loss(ypred, y, eps=1e-15):
    from .layers import SoftmaxOutput
    if not isinstance(ypred, SoftmaxOutput):
        raise NotImplementedError("device cross-entropy expects the softmax output of the network")
    rows, _grad, loss_sum = _fused(ypred, y)
    return _LossRows(rows, loss_sum)


def cross_entropy_derivative(ypred, y, eps=1e-15, with_softmax=False):
    from .layers import SoftmaxOutput
    if not (with_softmax and isinstance(ypred, SoftmaxOutput)):
        raise NotImplementedError("only the softmax-fused derivative (loss.py:58-66) is on the hot path")
    return DeviceArray(_fused(ypred, y)[1])


def _outside(name):
    def f(*_a, **_k):
        raise NotImplementedError(f"{name} is not used by the binarized scripts and not built here")
    f.__name__ = name
    return f


LOSS_FUNCS = {
    LossFuncs.cross_entropy: cross_entropy_loss,
    LossFuncs.mse: _outside("mse"),
    LossFuncs.hinge: _outside("hinge_loss"),
}

LOSS_DERIVATES = {
    LossFuncs.cross_entropy: cross_entropy_derivative,
    LossFuncs.mse: _outside("mse_derivative"),
    LossFuncs.hinge: _outside("hinge_derivative"),
}
