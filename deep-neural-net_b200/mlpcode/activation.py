"""Activation registry of the B200 path — same enum and tables as mlpcode/activation.py:134-168.

Hidden layers of a binarized network use `sign` (network.py:242-244 forces it); it and the softmax / identity
outputs of the BNN scripts are fused into the layer kernels (bnn_bn_sign, bnn_bn_apply, bnn_softmax_xent).  The
output layer may end in any other activation of the reference (network.py:236-262): sigmoid, tanh, relu,
leaky_relu and their derivatives are element-wise kernels (bnn_activation / bnn_activation_grad), the softmax
derivative — including the reference's second softmax of the activations, activation.py:56-68 — is
bnn_softmax_grad.  Every function takes and returns device arrays, like the reference's take and return ndarrays.
"""
from enum import Enum

import torch

from . import _native as nat
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


def _elementwise(kind):
    def f(x):
        xt = to_tensor(x, torch.float32).contiguous()
        out = torch.empty_like(xt)
        K_.activation(xt, kind, out)
        return DeviceArray(out)
    return f


def _elementwise_grad(kind):
    def f(dA, a):
        dt, at = to_tensor(dA, torch.float32).contiguous(), to_tensor(a, torch.float32).contiguous()
        assert dt.shape == at.shape
        out = torch.empty_like(at)
        K_.activation_grad(dt, at, kind, out)
        return DeviceArray(out)
    return f


sigmoid = _elementwise(nat.ACT_SIGMOID)                    # activation.py:17-22
tanh = _elementwise(nat.ACT_TANH)                          # :31-35
relu = _elementwise(nat.ACT_RELU)                          # :65-69
leaky_relu = _elementwise(nat.ACT_LEAKY_RELU)              # :80-84
hard_sigmoid = _elementwise(nat.ACT_HARD_SIGMOID)          # :95-99
hard_tanh = _elementwise(nat.ACT_HARD_TANH)                # :115-119
sigmoid_derivate = _elementwise_grad(nat.ACT_SIGMOID)      # :25-28
tanh_derivative = _elementwise_grad(nat.ACT_TANH)          # :38-41
relu_derivative = _elementwise_grad(nat.ACT_RELU)          # :72-77
leaky_relu_derivative = _elementwise_grad(nat.ACT_LEAKY_RELU)   # :87-92


def softmax_derivative(dA, a):
    """activation.py:56-68: the Jacobian of softmax evaluated at softmax(a) — the reference soft-maxes the
    activations a second time — contracted with dA."""
    dt, at = to_tensor(dA, torch.float32).contiguous(), to_tensor(a, torch.float32).contiguous()
    out = torch.empty_like(at)
    K_.softmax_grad(dt, at, out)
    return DeviceArray(out)


def identity(x):
    return x


def identity_derivative(dA, a):
    return DeviceArray(to_tensor(dA).clone())


def unitstep(x):
    """+1 where x >= 0 else -1 (activation.py:106-113)."""
    from .layers import BinaryLayer
    return BinaryLayer.binarize(x)


def hard_tanh_derivative(dA, a):
    """dA * 1[-1 <= a <= 1] (activation.py:122-131).  The binarized layer calls it with the post-sign
    activations a in {+-1}, where the mask is identically one (the layer skips the launch)."""
    return _elementwise_grad(nat.ACT_HARD_TANH)(dA, a)


def softmax(x):
    """Row softmax with max subtraction (activation.py:44-53)."""
    xt = to_tensor(x, torch.float32)
    B, C = xt.shape
    probs = empty((B, C))
    # clipping to [1e-15, 1] cannot change a softmax output by more than 1e-15
    K_.softmax_xent(xt, B, C, probs=probs)
    return DeviceArray(probs)


ACTIVATION_FUNCTIONS = {
    ActivationFuncs.sigmoid: sigmoid,
    ActivationFuncs.softmax: softmax,
    ActivationFuncs.relu: relu,
    ActivationFuncs.tanh: tanh,
    ActivationFuncs.sign: unitstep,
    ActivationFuncs.leaky_relu: leaky_relu,
    ActivationFuncs.identity: identity,
}

ACTIVATION_DERIVATIVES = {
    ActivationFuncs.relu: relu_derivative,
    ActivationFuncs.sigmoid: sigmoid_derivate,
    ActivationFuncs.softmax: softmax_derivative,
    ActivationFuncs.tanh: tanh_derivative,
    ActivationFuncs.sign: hard_tanh_derivative,
    ActivationFuncs.leaky_relu: leaky_relu_derivative,
    ActivationFuncs.identity: identity_derivative,
}
