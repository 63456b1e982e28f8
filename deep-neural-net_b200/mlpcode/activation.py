"""Activation registry of the B200 path — same enum and tables as mlpcode/activation.py:134-168.

Only the activations the binarized network can use are device-implemented: `sign` for hidden
layers (network.py:242-244 forces it for BNNs) and `softmax` / `identity` on the output layer.
They are fused into the layer kernels (bnn_bn_apply, bnn_softmax_xent); the functions below exist
so code written against ACTIVATION_FUNCTIONS / ACTIVATION_DERIVATIVES keeps working on device arrays.
"""
from enum import Enum

import torch

from . import kernels as K_
from .device import DeviceArray, empty, to_tensor


class ActivationFuncs(Enum):
    sigmoid = "sigmoid"
    softmax = "softmax"
    relu = "relu"
    tanh = "tanh"
    sign = "unitstep"
    leaky_relu = "leaky_relu"
    identity = "identity"

    def __repr__(self):
        return self.value

    def __str__(self):
        return self.value


def _outside(name):
    def f(*_a, **_k):
        raise NotImplementedError(f"{name} is not part of the binarized hot path of this build")
    f.__name__ = name
    return f


def identity(x):
    return x


def identity_derivative(dA, a):
    return DeviceArray(to_tensor(dA).clone())


def unitstep(x):
    """+1 where x >= 0 else -1 (activation.py:106-113)."""
    from .layers import BinaryLayer
    return BinaryLayer.binarize(x)


def hard_tanh_derivative(dA, a):
    """dA * 1[-1 <= a <= 1] (activation.py:122-131).  The layer calls it with the post-sign
    activations a in {+-1}, where the mask is identically one, so it is the identity."""
    return DeviceArray(to_tensor(dA).clone())


def softmax(x):
    """Row softmax with max subtraction (activation.py:44-53)."""
    xt = to_tensor(x, torch.float32)
    B, C = xt.shape
    probs = empty((B, C))
    # clipping to [1e-15, 1] cannot change a softmax output by more than 1e-15
    K_.softmax_xent(xt, B, C, probs=probs)
    return DeviceArray(probs)


ACTIVATION_FUNCTIONS = {
    ActivationFuncs.sigmoid: _outside("sigmoid"),
    ActivationFuncs.softmax: softmax,
    ActivationFuncs.relu: _outside("relu"),
    ActivationFuncs.tanh: _outside("tanh"),
    ActivationFuncs.sign: unitstep,
    ActivationFuncs.leaky_relu: _outside("leaky_relu"),
    ActivationFuncs.identity: identity,
}

ACTIVATION_DERIVATIVES = {
    ActivationFuncs.relu: _outside("relu_derivative"),
    ActivationFuncs.sigmoid: _outside("sigmoid_derivate"),
    ActivationFuncs.softmax: _outside("softmax_derivative"),
    ActivationFuncs.tanh: _outside("tanh_derivative"),
    ActivationFuncs.sign: hard_tanh_derivative,
    ActivationFuncs.leaky_relu: _outside("leaky_relu_derivative"),
    ActivationFuncs.identity: identity_derivative,
}
