"""Data helpers of the B200 path — the part of mlpcode/utils.py the fault-sweep scripts import
(`from mlpcode.utils import DATASETS, MODELDIR, loadDataset, normalize`, imageBitflipTrainModelSel.py:2).

`normalize` (utils.py:178-187) is the one function here that sits on the hot path: the image sweep
calls it on the corrupted uint8 test set before every evaluation.  For a uint8 device array it does NOT
produce the fp32 copy: it measures min / max (bnn_minmax_u8) and returns a `NormalizedU8`, a lazy view that
layer 0 of a binarized network consumes directly as bytes (exact u8 x s8 GEMM with the affine map folded
into the sign thresholds, layers.BinaryLayer.forward); the fp32 values are materialised, bit-identical to
the reference's, only if generic code touches them.

Dataset files are not shipped (mlpcode/data/*/readme.txt in the reference says where to get them); the IDX /
NPY readers below are host-side file I/O that returns device arrays, kept minimal.
"""
from __future__ import annotations

import os
import struct
from enum import Enum
from pathlib import Path

import numpy as np
import torch

from . import kernels as K_
from .device import DeviceArray, asarray, empty, to_tensor, zeros

_PKG = Path(__file__).resolve().parent
DATADIR = Path(os.environ.get("BNN_DATADIR", _PKG / "data"))
MODELDIR = _PKG / "models"
MNIST_CLASSES = 10

_MNIST_C = ("brightness canny_edges dotted_line fog glass_blur identity impulse_noise motion_blur rotate scale "
            "shear shot_noise spatter stripe translate zigzag contrast defocus_blur elastic_transform frost "
            "gaussian_blur gaussian_noise inverse jpeg_compression line pessimal_noise pixelate quantize "
            "saturate snow speckle_noise zoom_blur").split()
DATASETS = Enum("DATASETS", {"mnist": "mnist", "fashion": "fashion-mnist",
                             **{f"mnistc_{c}": f"mnist_c-{c}" for c in _MNIST_C}})
DATASETS.__str__ = lambda self: self.value
DATASETS.__repr__ = lambda self: self.value


# ----------------------------------------------------------------------------------------------
# normalize                                                                     utils.py:178-187
# ----------------------------------------------------------------------------------------------
class NormalizedU8(DeviceArray):
    """`normalize(X_uint8, newMin, newMax)` kept as (bytes, global min / max, target range).

    Behaves like the fp32 [R, F] array the reference returns (`shape`, `dtype`, row slicing, `.get()`,
    `.t`), but the fp32 values only come into existence when one of those is used; a binarized layer 0 in
    evaluation mode reads `u8` / `mm_u32` instead.  Row slices share the parent's min / max — the
    reference normalizes the whole set before `evaluate` slices it (network.py:406-421)."""

    def __init__(self, u8: torch.Tensor, mm_u32: torch.Tensor, mm_f32: torch.Tensor, new_min: float,
                 new_max: float):
        self.u8, self.mm_u32, self.mm_f32 = u8, mm_u32, mm_f32
        self.new_min, self.new_max = float(new_min), float(new_max)
        self._dense = None

    @property
    def t(self) -> torch.Tensor:
        if self._dense is None:
            out = empty(tuple(self.u8.shape), torch.float32)
            K_.normalize_u8(self.u8, self.mm_f32[0:1], self.mm_f32[1:2], self.new_min, self.new_max, out)
            self._dense = out
        return self._dense

    @t.setter
    def t(self, v):
        self._dense = v

    @property
    def shape(self):
        return tuple(self.u8.shape)

    @property
    def ndim(self):
        return self.u8.ndim

    @property
    def size(self):
        return self.u8.numel()

    @property
    def dtype(self):
        return np.dtype(np.float32)

    def __len__(self):
        return self.u8.shape[0]

    def __getitem__(self, idx):
        rows_only = isinstance(idx, (slice, torch.Tensor, DeviceArray, np.ndarray, list))
        if self.u8.ndim == 2 and rows_only:
            if isinstance(idx, DeviceArray):
                idx = idx.t
            elif isinstance(idx, (np.ndarray, list)):
                idx = torch.as_tensor(np.asarray(idx), device=self.u8.device)
            return NormalizedU8(self.u8[idx].contiguous(), self.mm_u32, self.mm_f32, self.new_min, self.new_max)
        return super().__getitem__(idx)


def normalize(X, newMin=0., newMax=1.):
    """Min-max scaling to [newMin, newMax] with the GLOBAL min / max of X (utils.py:178-187)."""
    t = to_tensor(X)
    if t.dtype == torch.uint8:
        mm_u32, mm_f32 = zeros((2,), torch.int32), zeros((2,), torch.float32)
        K_.minmax_u8(t, mm_u32, mm_f32)
        return NormalizedU8(t, mm_u32, mm_f32, newMin, newMax)
    # real-valued input (not on the sweep path): the reference's operation order, on the device
    lo = t.min()
    span = (t.max() - lo).to(torch.float64)
    out = (t - lo).to(torch.float32) * (float(newMax) - float(newMin))
    out = (out.to(torch.float64) / span).to(torch.float32)
    out += float(newMin)
    return DeviceArray(out)


# ----------------------------------------------------------------------------------------------
# labels / splits                                                       utils.py:93-119, 141-175
# ----------------------------------------------------------------------------------------------
def oneHotEncode(classes: int, y):
    yt = to_tensor(y).reshape(-1).to(torch.int64)
    return DeviceArray(torch.eye(classes, dtype=torch.int8, device=yt.device)[yt])


def split_train_valid(X, y, valSize: int = 10000, shuffle=True):
    X, y = asarray(X), asarray(y)
    n = len(X)
    assert len(y) == n
    if shuffle:
        perm = torch.randperm(n, device=X.t.device)
        X, y = X[perm], y[perm]
    return X[:n - valSize], X[n - valSize:], y[:n - valSize], y[n - valSize:]


# ----------------------------------------------------------------------------------------------
# dataset files (host I/O)                                                     utils.py:191-661
# ----------------------------------------------------------------------------------------------
def _read_idx(path: Path, labels: bool) -> np.ndarray:
    if not path.exists():
        raise FileNotFoundError(f"{path} not found: datasets are not shipped; put the IDX / NPY files under "
                                f"{DATADIR} (or point BNN_DATADIR at them)")
    with path.open("rb") as f:
        header = f.read(8 if labels else 16)
        magic = struct.unpack(">I", header[:4])[0]
        if magic != (2049 if labels else 2051):
            raise ValueError(f"Magic number mismatch in {path.name}: {magic}")
        return np.frombuffer(f.read(), dtype=np.uint8).copy()   # writable: it becomes a tensor


def _xy(X: np.ndarray, y: np.ndarray, encoded: bool, test: bool):
    Xd = asarray(np.ascontiguousarray(X.reshape(len(y), -1)))         # uint8 [instances, features]
    if encoded and not test:
        return Xd, oneHotEncode(MNIST_CLASSES, y.astype(np.int64))
    return Xd, asarray(y.astype(np.uint8 if test else np.int8))


def _load_idx_dir(d: Path, encoded: bool):
    tr = _xy(_read_idx(d / "train-images-idx3-ubyte", False), _read_idx(d / "train-labels-idx1-ubyte", True),
             encoded, False)
    te = _xy(_read_idx(d / "t10k-images-idx3-ubyte", False), _read_idx(d / "t10k-labels-idx1-ubyte", True),
             encoded, True)
    return tr[0], tr[1], te[0], te[1]


def _load_npy_dir(d: Path, encoded: bool):
    def npy(name):
        p = d / name
        if not p.exists():
            raise FileNotFoundError(f"{p} not found: datasets are not shipped (BNN_DATADIR)")
        return np.load(p)
    tr = _xy(npy("train_images.npy"), npy("train_labels.npy"), encoded, False)
    te = _xy(npy("test_images.npy"), npy("test_labels.npy"), encoded, True)
    return tr[0], tr[1], te[0], te[1]


def loadDataset(dataset, useGpu=True, encoded=True, quant_precision=None, quantize_test=False):
    """(trainX, trainY, testX, testY) as device arrays; images stay uint8 as in the files (the scripts
    normalize them themselves, test.py:23-24)."""
    if not useGpu:
        raise NotImplementedError("this build is the device path only (useGpu=True)")
    if quant_precision is not None:
        raise NotImplementedError("utils.quantize is outside the binarized hot path")
    name = str(dataset)
    if name.startswith("mnist_c"):
        return _load_npy_dir(DATADIR / "mnist-c" / name.split("-", 1)[1], encoded)
    return _load_idx_dir(DATADIR / name, encoded)
