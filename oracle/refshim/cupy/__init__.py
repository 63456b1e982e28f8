"""Test-infrastructure only: a `cupy` stand-in that forwards to NumPy.

The reference (volf52/deep-neural-net) does `import cupy as cp` at the top of six
modules (mlpcode/layers.py:1, activation.py:3, loss.py:3, callbacks.py:2, utils.py:7,
network.py:6) although its CPU path only touches `cp.get_array_module`, `cp.asnumpy`
and `cp.cuda.Stream.null.synchronize`.  CuPy is not installed in this image, so the
golden-vector generator (oracle/gen_golden.py) puts this directory first on sys.path
to import the *unmodified* reference NumPy path.  Nothing in the product imports it.
"""
import numpy as _np
from numpy import *  # noqa: F401,F403  (reference uses cp.<ufunc> only on the GPU path)
from numpy import ndarray, random  # noqa: F401
from . import cuda  # noqa: F401


def get_array_module(*_args):
    return _np


def asnumpy(a, *_, **__):
    return _np.asarray(a)
