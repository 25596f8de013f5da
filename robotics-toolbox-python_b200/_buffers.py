"""Device-buffer plumbing: PyTorch tensors are used ONLY as HBM buffers and stream handles
for the C ABI (data_ptr / cuda_stream); no torch op is on the compute path."""
from __future__ import annotations

import numpy as np

from . import _lib

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as t

        _torch = t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError(
            "b2kin: no CUDA device is available and there is no CPU fallback "
            "(the CPU restatement under oracle/ is test infrastructure only)"
        )
    return t


def is_tensor(x) -> bool:
    return _torch is not None and isinstance(x, _torch.Tensor) or type(x).__module__.startswith("torch")


def pick_dtype(x, dtype=None):
    """fp64 unless the caller passes float32 data or dtype=float32 (the reference is fp64-only)."""
    if dtype is not None:
        if isinstance(dtype, np.dtype):
            d = dtype
        elif isinstance(dtype, type):  # np.float32 / np.float64 / float
            d = np.dtype(dtype)
        else:  # "float32", torch.float32
            d = np.dtype(str(dtype).replace("torch.", ""))
        if d not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("dtype must be float32 or float64")
        return d
    if is_tensor(x):
        return np.dtype(np.float32) if x.dtype == torch().float32 else np.dtype(np.float64)
    if isinstance(x, np.ndarray) and x.dtype == np.float32:
        return np.dtype(np.float32)
    return np.dtype(np.float64)


def code(dt: np.dtype) -> int:
    return _lib.F32 if dt == np.dtype(np.float32) else _lib.F64


def tdtype(dt: np.dtype):
    t = torch()
    return t.float32 if dt == np.dtype(np.float32) else t.float64


def to_device(x, dt: np.dtype, device=None):
    """numpy / list / tensor -> contiguous CUDA tensor of dtype dt (copying only if needed)."""
    t = require_cuda()
    if is_tensor(x):
        y = x
        if not y.is_cuda:
            y = y.cuda(device) if device is not None else y.cuda()
        if y.dtype != tdtype(dt):
            y = y.to(tdtype(dt))
        return y.contiguous()
    a = np.ascontiguousarray(np.asarray(x), dtype=dt)
    y = t.from_numpy(a)
    return y.cuda(device) if device is not None else y.cuda()


def empty(shape, dt: np.dtype, like=None):
    t = require_cuda()
    dev = like.device if like is not None else None
    return t.empty(shape, dtype=tdtype(dt), device=dev if dev is not None else "cuda")


def empty_i32(shape, like=None):
    t = require_cuda()
    return t.empty(shape, dtype=t.int32, device=like.device if like is not None else "cuda")


def stream_ptr(like=None) -> int:
    t = torch()
    dev = like.device if like is not None else None
    return int(t.cuda.current_stream(dev).cuda_stream)


def ptr(x) -> int:
    return 0 if x is None else int(x.data_ptr())


def to_host(x):
    return x.detach().cpu().numpy()


def check_numeric(x, name="q"):
    """The reference raises TypeError('Symbolic value') for non-numeric input (fknm.cpp:1304-1318)."""
    if is_tensor(x):
        return
    if isinstance(x, str):
        raise TypeError(f"{name} must be numeric")
    a = np.asarray(x)
    if a.dtype == object or not (np.issubdtype(a.dtype, np.number) or a.dtype == bool):
        raise TypeError("Symbolic value")
