"""Tensor plumbing shared by the op wrappers: PyTorch is used only as the device allocator and
stream provider; every computation goes through the C ABI."""
from __future__ import annotations

import torch

from . import _lib


def stream_ptr() -> int:
    """The current stream of the CURRENT device.  req()/rows() reject tensors that live on any other device, so a
    kernel is never enqueued on one GPU's stream with another GPU's pointers."""
    return torch.cuda.current_stream().cuda_stream


def _same_device(t, name):
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"{name}: tensor lives on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                           "call torch.cuda.set_device() / use `with torch.cuda.device(t.device)` around dispu_b200 ops")


def req(t, dtype, ndim, name):
    """The reference wrappers' OP_REQUIRES shape/dtype checks -> ValueError (InvalidArgument)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor on a CUDA device, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: dispu_b200 ops are GPU-only (no CPU fallback); got a {t.device} tensor")
    _same_device(t, name)
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dimensions, got shape {tuple(t.shape)}")
    return t.contiguous()


def ptr(t):
    return None if t is None else t.data_ptr()


# Optional per-call device timing (bench.py's roofline leg): a list that receives (key, start_event, end_event) for
# every C-ABI call while it is set.  Never set on the product path.
_profile = None


def _profile_key(name, args):
    if name == "dispu_dense_tf32x3_ex":
        a = args[0]._obj
        tag = ((" gather" if a.gather_idx else "") + (" batched" if a.rows_per_batch else "") + (" +res" if a.R1 else "") +
               (f" x{a.n_blocks}blocks" if a.n_blocks > 1 else "") + (" +weightnet" if a.wn_xyz else ""))
        return f"{name} M{a.M} K{a.K} N{a.N}{tag}"
    ints = [str(x) for x in args if isinstance(x, int) and not isinstance(x, bool) and 0 <= x < (1 << 24)]
    return name + (" " + ",".join(ints[:6]) if ints else "")


def call(name, *args):
    fn = getattr(_lib.lib(), name)
    if _profile is None:
        _lib.check(name, fn(*args))
        return
    key = _profile_key(name, args)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    _profile.append((key, e0, e1))
    _lib.check(name, rc)


def rows(t, name="tensor"):
    """(data_ptr, n_rows, row_stride) of a float32 CUDA tensor whose last axis is contiguous and whose rows are
    uniformly strided — a contiguous tensor or a column slice of one (x[..., a:b])."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: dispu_b200 ops are GPU-only (no CPU fallback)")
    _same_device(t, name)
    if t.dtype != torch.float32:
        raise ValueError(f"{name}: expected float32, got {t.dtype}")
    if t.dim() < 1 or (t.shape[-1] > 1 and t.stride(-1) != 1):
        raise ValueError(f"{name}: last axis must be contiguous")
    if t.dim() == 1:
        return t.data_ptr(), 1, t.shape[0]
    ld = t.stride(-2) if t.shape[-2] > 1 else t.shape[-1]
    n = t.shape[-2]
    for i in range(t.dim() - 3, -1, -1):
        if t.shape[i] > 1 and t.stride(i) != ld * n:
            raise ValueError(f"{name}: rows are not uniformly strided (shape {tuple(t.shape)}, strides {t.stride()})")
        n *= t.shape[i]
    return t.data_ptr(), n, ld
