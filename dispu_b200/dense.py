"""Dense layer on the tcgen05 tensor cores with FP32-level accuracy (3xTF32): y = act(x.W + b).
The weight is packed once (TF32 hi/lo planes, UMMA layout); see dispu_dense_tf32x3 in include/dispu_b200.h."""
from __future__ import annotations

import torch

from . import _lib
from ._t import call, ptr, req, stream_ptr


class PackedLinear:
    """W (K,N) float32 with N <= 256, optional bias (N,).  __call__(x (..., K), relu) -> (..., N)."""

    def __init__(self, weight, bias=None):
        weight = req(weight, torch.float32, 2, "weight")
        self.K, self.N = weight.shape
        if self.N > 256:
            raise ValueError("PackedLinear supports N <= 256")
        self.bias = None if bias is None else req(bias, torch.float32, 1, "bias")
        nbytes = int(_lib.lib().dispu_dense_packed_bytes(self.K, self.N))
        self.packed = torch.empty((nbytes,), dtype=torch.uint8, device=weight.device)
        call("dispu_dense_pack", self.K, self.N, ptr(weight), ptr(self.packed), stream_ptr())

    def __call__(self, x, relu=False):
        x = x.contiguous()
        if x.dtype != torch.float32 or not x.is_cuda or x.shape[-1] != self.K:
            raise ValueError(f"PackedLinear expects a float32 CUDA tensor with last dim {self.K}")
        M = x.numel() // self.K
        y = torch.empty((*x.shape[:-1], self.N), dtype=torch.float32, device=x.device)
        call("dispu_dense_tf32x3", M, self.K, self.N, ptr(x), ptr(self.packed), ptr(self.bias), 1 if relu else 0, ptr(y),
             stream_ptr())
        return y
