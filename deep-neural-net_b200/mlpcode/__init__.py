"""mlpcode — B200-native (sm_100a) drop-in for the binarized-layer hot path of
volf52/deep-neural-net.  Same module and class names as the reference package
(`mlpcode.layers`, `mlpcode.network.Network`, `mlpcode.optim`, `mlpcode.loss`,
`mlpcode.callbacks`, `mlpcode.activation`); the arithmetic runs in libbnn_b200.so
(deep-neural-net_b200/csrc, C ABI in include/bnn_b200.h) on raw device pointers.

Importing the package does not touch CUDA; building a layer does, and fails loudly when no
device or no built library is present — there is no CPU fallback.
"""
__all__ = ["layers", "network", "activation", "loss", "optim", "callbacks", "kernels", "device",
           "parallel", "sweep"]
