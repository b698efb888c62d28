"""Tensor plumbing shared by the op wrappers: PyTorch is used only as the device allocator and
stream provider; every computation goes through the C ABI."""
from __future__ import annotations

import torch

from . import _lib


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def req(t, dtype, ndim, name):
    """The reference wrappers' OP_REQUIRES shape/dtype checks -> ValueError (InvalidArgument)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor on a CUDA device, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: dispu_b200 ops are GPU-only (no CPU fallback); got a {t.device} tensor")
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dimensions, got shape {tuple(t.shape)}")
    return t.contiguous()


def ptr(t):
    return None if t is None else t.data_ptr()


def call(name, *args):
    fn = getattr(_lib.lib(), name)
    _lib.check(name, fn(*args))
