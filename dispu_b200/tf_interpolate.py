"""Mirror of tf_ops/interpolation/tf_interpolate.py on torch CUDA tensors.  The reference only has
CPU kernels for these ops (tf_interpolate.cpp:187,222,262); these are GPU kernels."""
from __future__ import annotations

import torch

from ._t import call, ptr, req, stream_ptr


def three_nn(xyz1, xyz2):
    """tf_interpolate.py:8.  xyz1 (b,n,3) unknown, xyz2 (b,m,3) known -> dist (b,n,3) SQUARED, idx (b,n,3) i32."""
    xyz1 = req(xyz1, torch.float32, 3, "xyz1")
    xyz2 = req(xyz2, torch.float32, 3, "xyz2")
    b, n, c = xyz1.shape
    if c != 3:
        raise ValueError("ThreeNN expects (b,n,3) xyz1 shape")  # tf_interpolate.cpp:165
    if xyz2.shape[0] != b or xyz2.shape[2] != 3:
        raise ValueError("ThreeNN expects (b,m,3) xyz2 shape")  # tf_interpolate.cpp:170
    m = xyz2.shape[1]
    dist = torch.empty((b, n, 3), dtype=torch.float32, device=xyz1.device)
    idx = torch.empty((b, n, 3), dtype=torch.int32, device=xyz1.device)
    call("dispu_three_nn", b, n, m, ptr(xyz1), ptr(xyz2), ptr(dist), ptr(idx), stream_ptr())
    return dist, idx


class _ThreeInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx, weight):
        b, m, c = points.shape
        n = idx.shape[1]
        out = torch.empty((b, n, c), dtype=torch.float32, device=points.device)
        call("dispu_three_interpolate", b, m, c, n, ptr(points), ptr(idx), ptr(weight), ptr(out), stream_ptr())
        ctx.save_for_backward(idx, weight)
        ctx.shape = (b, m, c)
        return out

    @staticmethod
    def backward(ctx, grad_out):  # tf_interpolate.py:29-34
        idx, weight = ctx.saved_tensors
        b, m, c = ctx.shape
        n = idx.shape[1]
        grad_out = grad_out.contiguous()
        g = torch.empty((b, m, c), dtype=torch.float32, device=grad_out.device)
        call("dispu_three_interpolate_grad", b, n, c, m, ptr(grad_out), ptr(idx), ptr(weight), ptr(g), stream_ptr())
        return g, None, None


def three_interpolate(points, idx, weight):
    """tf_interpolate.py:19.  points (b,m,c), idx (b,n,3) i32, weight (b,n,3) -> (b,n,c)."""
    points = req(points, torch.float32, 3, "points")
    idx = req(idx, torch.int32, 3, "idx")
    weight = req(weight, torch.float32, 3, "weight")
    b = points.shape[0]
    if idx.shape[0] != b or idx.shape[2] != 3:
        raise ValueError("ThreeInterpolate expects (b,n,3) idx shape")  # tf_interpolate.cpp:202
    if tuple(weight.shape) != tuple(idx.shape):
        raise ValueError("ThreeInterpolate expects (b,n,3) weight shape")  # tf_interpolate.cpp:205
    return _ThreeInterpolate.apply(points, idx, weight)


def three_nn_weights(dist):
    """Inverse-distance weights of pointnet_fp_module (Common/pointnet_util.py:205-208)."""
    dist = req(dist, torch.float32, 3, "dist")
    b, n, _ = dist.shape
    w = torch.empty_like(dist)
    call("dispu_three_nn_weights", b, n, ptr(dist), ptr(w), stream_ptr())
    return w
