"""Column plumbing shared by the array-level methods: SoA columns are 1-D contiguous
torch CUDA tensors (device memory + stream handling is all torch is used for)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi

_SUF = {torch.float64: "f64", torch.float32: "f32"}
_NP = {torch.float64: np.float64, torch.float32: np.float32}


def suffix_of(t: torch.Tensor) -> str:
    try:
        return _SUF[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported column dtype {t.dtype} (Float64 / Float32 only)") from None


def check_columns(cols, names):
    """All columns: same dtype, same CUDA device, 1-D, contiguous, same length."""
    first = cols[0]
    if not isinstance(first, torch.Tensor):
        raise TypeError(f"{names[0]} must be a torch tensor (device array)")
    if not first.is_cuda:
        raise _abi.CuMicroError(
            "cumicro array methods take CUDA device arrays; there is no CPU fallback "
            "(use the *_host entry points for host buffers)")
    n = first.shape[0] if first.dim() == 1 else -1
    for t, nm in zip(cols, names):
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{nm} must be a torch tensor")
        if t.dtype != first.dtype or t.device != first.device:
            raise TypeError(f"{nm}: dtype/device {t.dtype}/{t.device} differs from {names[0]} {first.dtype}/{first.device}")
        if t.dim() != 1 or t.shape[0] != n or not t.is_contiguous():
            raise ValueError(f"{nm} must be a contiguous 1-D column of length {n}")
    return suffix_of(first), n, first.device


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def ptr_table(tensors):
    """HOST array of device column pointers (NULL for None)."""
    arr = (C.c_void_p * len(tensors))(*[t.data_ptr() if t is not None else None for t in tensors])
    return arr


def stream_handle(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def zero_column(like: torch.Tensor):
    """A length-n all-zero column that occupies one element of HBM (stride 0): the
    reference returns literal zeros for these fields (BMT:840-853); materialising
    them would add 32 B/point of dead traffic."""
    key = (like.device, like.dtype)
    z = _ZERO.get(key)
    if z is None:   # one element per (device, dtype) for the life of the process: no fill kernel per call
        z = _ZERO[key] = torch.zeros(1, dtype=like.dtype, device=like.device)
    return z.expand(like.shape[0])


_ZERO = {}


class Tendencies(dict):
    """NamedTuple-like result: ``out.dq_lcl_dt`` or ``out['dq_lcl_dt']``."""

    def __getattr__(self, name):
        # AttributeError (not KeyError) for a missing name: hasattr(), getattr(x, n, default), copy and pickle rely on it
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None
