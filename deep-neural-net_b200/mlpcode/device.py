"""Device-array plumbing for the B200 path.

The reference switches array module with `xp = cp if gpu else np` (mlpcode/layers.py:10-15,
network.py:32-35).  CuPy is not part of this stack, so device memory, streams and (for the
data-parallel path) the NCCL process group come from PyTorch — as plumbing only: every
arithmetic kernel on the hot path is a libbnn_b200 call on raw device pointers.

`DeviceArray` is the ndarray-like the unchanged Network/driver code manipulates; it supports
exactly the operations SURVEY §8(b) lists (`shape`, `size`, `dtype`, `copy`, `astype`, slicing
and fancy row indexing, `argmax(axis=1)`, `==`, `sum`, `mean`, `T`-free) and exposes
`__cuda_array_interface__` / `__dlpack__` so CuPy or any other consumer can view it zero-copy.
"""
from __future__ import annotations

import numpy as np
import torch

_NP2T = {
    np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
    np.dtype(np.float16): torch.float16, np.dtype(np.int8): torch.int8,
    np.dtype(np.uint8): torch.uint8, np.dtype(np.int16): torch.int16,
    np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
    np.dtype(np.bool_): torch.bool,
}
_T2NP = {v: k for k, v in _NP2T.items()}


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("mlpcode (B200 build): no CUDA device visible — this path has no CPU "
                           "fallback; use the reference NumPy path for CPU runs")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream, handed to every libbnn_b200 call."""
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    if t is None:
        return None
    if isinstance(t, DeviceArray):
        t = t.t
    return t.data_ptr()


def pad_to(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def empty(shape, dtype=torch.float32) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device=require_cuda())


def zeros(shape, dtype=torch.float32) -> torch.Tensor:
    return torch.zeros(shape, dtype=dtype, device=require_cuda())


def to_tensor(a, dtype=None) -> torch.Tensor:
    """Accept DeviceArray / torch tensor / NumPy array / __cuda_array_interface__ object."""
    if isinstance(a, DeviceArray):
        t = a.t
    elif isinstance(a, torch.Tensor):
        t = a
    elif isinstance(a, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(a))
    elif hasattr(a, "__dlpack__"):
        t = torch.from_dlpack(a)
    elif hasattr(a, "__cuda_array_interface__"):
        t = torch.as_tensor(a, device=require_cuda())
    else:
        t = torch.as_tensor(np.asarray(a))
    if t.device.type != "cuda":
        t = t.to(require_cuda(), non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class DeviceArray:
    """Minimal ndarray-like view of a CUDA tensor."""

    __array_priority__ = 1000

    def __init__(self, t: torch.Tensor):
        self.t = t

    # ---- ndarray surface ---------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.ndim

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return _T2NP[self.t.dtype]

    def __len__(self):
        return self.t.shape[0]

    def copy(self):
        return DeviceArray(self.t.clone())

    def astype(self, dtype):
        return DeviceArray(self.t.to(_NP2T[np.dtype(dtype)]))

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            idx = tuple(i.t if isinstance(i, DeviceArray) else i for i in idx)
        elif isinstance(idx, DeviceArray):
            idx = idx.t
        return DeviceArray(self.t[idx])

    def argmax(self, axis=None):
        t = self.t
        if t.dtype in (torch.int8, torch.uint8, torch.bool):  # one-hot labels arrive as int8
            t = t.to(torch.int32)
        return DeviceArray(t.argmax(dim=axis) if axis is not None else t.argmax())

    def sum(self, axis=None):
        return DeviceArray(self.t.sum(dim=axis) if axis is not None else self.t.sum())

    def mean(self, axis=None):
        return DeviceArray(self.t.mean(dim=axis) if axis is not None else self.t.mean())

    def __eq__(self, other):  # noqa: D105 (elementwise, like ndarray)
        return DeviceArray(self.t == (other.t if isinstance(other, DeviceArray) else other))

    __hash__ = None

    def __float__(self):
        return float(self.t)

    def __int__(self):
        return int(self.t)

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype})"

    # ---- interop -----------------------------------------------------------------------
    def get(self) -> np.ndarray:
        """Device -> host copy (the analogue of cupy.ndarray.get / cp.asnumpy)."""
        return self.t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a.astype(dtype) if dtype is not None else a

    @property
    def __cuda_array_interface__(self):
        return self.t.__cuda_array_interface__

    def __dlpack__(self, stream=None):
        return self.t.__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return self.t.__dlpack_device__()


def asnumpy(a) -> np.ndarray:
    """cp.asnumpy equivalent (network.py:442)."""
    if isinstance(a, DeviceArray):
        return a.get()
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def asarray(a, dtype=None) -> DeviceArray:
    td = _NP2T[np.dtype(dtype)] if dtype is not None else None
    return DeviceArray(to_tensor(a, td))
