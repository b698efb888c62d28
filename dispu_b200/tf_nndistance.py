"""Mirror of tf_ops/nn_distance/tf_nndistance.py on torch CUDA tensors."""
from __future__ import annotations

import torch

from ._t import call, ptr, req, stream_ptr


class _NnDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        dev = xyz1.device
        d1 = torch.empty((b, n), dtype=torch.float32, device=dev)
        i1 = torch.empty((b, n), dtype=torch.int32, device=dev)
        d2 = torch.empty((b, m), dtype=torch.float32, device=dev)
        i2 = torch.empty((b, m), dtype=torch.int32, device=dev)
        call("dispu_nn_distance", b, n, ptr(xyz1), m, ptr(xyz2), ptr(d1), ptr(i1), ptr(d2), ptr(i2), stream_ptr())
        ctx.save_for_backward(xyz1, xyz2, i1, i2)
        ctx.mark_non_differentiable(i1, i2)
        return d1, i1, d2, i2

    @staticmethod
    def backward(ctx, g1, _gi1, g2, _gi2):  # tf_nndistance.py:31-37
        xyz1, xyz2, i1, i2 = ctx.saved_tensors
        b, n, _ = xyz1.shape
        m = xyz2.shape[1]
        g1 = g1.contiguous()
        g2 = g2.contiguous()
        o1 = torch.empty_like(xyz1)
        o2 = torch.empty_like(xyz2)
        call("dispu_nn_distance_grad", b, n, ptr(xyz1), m, ptr(xyz2), ptr(g1), ptr(i1), ptr(g2), ptr(i2), ptr(o1),
             ptr(o2), stream_ptr())
        return o1, o2


def nn_distance(xyz1, xyz2):
    """tf_nndistance.py:14.  xyz1 (b,n,3), xyz2 (b,m,3) -> dist1 (b,n), idx1 (b,n) i32, dist2 (b,m), idx2 (b,m) i32.
    Squared distances; differentiable w.r.t. both clouds."""
    xyz1 = req(xyz1, torch.float32, 3, "xyz1")
    xyz2 = req(xyz2, torch.float32, 3, "xyz2")
    if xyz1.shape[2] != 3:
        raise ValueError("NnDistance only accepts 3d point set xyz1")  # tf_nndistance.cpp:52
    if xyz2.shape[2] != 3:
        raise ValueError("NnDistance only accepts 3d point set xyz2")  # tf_nndistance.cpp:57
    if xyz2.shape[0] != xyz1.shape[0]:
        raise ValueError("NnDistance expects xyz1 and xyz2 have same batch size")  # tf_nndistance.cpp:58
    return _NnDistance.apply(xyz1, xyz2)


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    """The NnDistanceGrad op (tf_nndistance.cpp:10-18)."""
    xyz1 = req(xyz1, torch.float32, 3, "xyz1")
    xyz2 = req(xyz2, torch.float32, 3, "xyz2")
    g1 = req(grad_dist1, torch.float32, 2, "grad_dist1")
    g2 = req(grad_dist2, torch.float32, 2, "grad_dist2")
    i1 = req(idx1, torch.int32, 2, "idx1")
    i2 = req(idx2, torch.int32, 2, "idx2")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    o1 = torch.empty_like(xyz1)
    o2 = torch.empty_like(xyz2)
    call("dispu_nn_distance_grad", b, n, ptr(xyz1), m, ptr(xyz2), ptr(g1), ptr(i1), ptr(g2), ptr(i2), ptr(o1), ptr(o2),
         stream_ptr())
    return o1, o2
