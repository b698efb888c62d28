"""Mirror of tf_ops/sampling/tf_sampling.py on torch CUDA tensors, backed by the sm_100a kernels.
Same function names, argument order and meaning as the reference wrappers."""
from __future__ import annotations

import torch

from ._t import call, ptr, req, stream_ptr


def farthest_point_sample(npoint, inp):
    """tf_sampling.py:48 (note: npoint first).  inp (b,n,3) f32 -> (b,npoint) i32, no gradient."""
    inp = req(inp, torch.float32, 3, "inp")
    b, n, c = inp.shape
    if c != 3:
        raise ValueError("FarthestPointSample expects (batch_size,num_points,3) inp shape")  # tf_sampling.cpp:105
    if npoint <= 0:
        raise ValueError("FarthestPointSample expects positive npoint")  # tf_sampling.cpp:99
    out = torch.empty((b, npoint), dtype=torch.int32, device=inp.device)
    temp = None
    from . import _lib
    if n > _lib.lib().dispu_fps_onchip_capacity():
        temp = torch.empty((32, n), dtype=torch.float32, device=inp.device)  # tf_sampling.cpp:115
    call("dispu_fps", b, n, npoint, ptr(inp), ptr(temp), ptr(out), stream_ptr())
    return out


class _GatherPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, idx):
        b, n, _ = inp.shape
        m = idx.shape[1]
        out = torch.empty((b, m, 3), dtype=torch.float32, device=inp.device)
        call("dispu_gather_point", b, n, m, ptr(inp), ptr(idx), ptr(out), stream_ptr())
        ctx.save_for_backward(idx)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, out_g):  # tf_sampling.py:43-47
        (idx,) = ctx.saved_tensors
        out_g = out_g.contiguous()
        b, m, _ = out_g.shape
        inp_g = torch.empty((b, ctx.n, 3), dtype=torch.float32, device=out_g.device)
        call("dispu_gather_point_grad", b, ctx.n, m, ptr(out_g), ptr(idx), ptr(inp_g), stream_ptr())
        return inp_g, None


def gather_point(inp, idx):
    """tf_sampling.py:29.  inp (b,n,3), idx (b,m) i32 -> (b,m,3); differentiable w.r.t. inp."""
    inp = req(inp, torch.float32, 3, "inp")
    idx = req(idx, torch.int32, 2, "idx")
    if inp.shape[2] != 3:
        raise ValueError("GatherPoint expects (batch_size,num_points,3) inp shape")  # tf_sampling.cpp:132
    if idx.shape[0] != inp.shape[0]:
        raise ValueError("GatherPoint expects (batch_size,num_result) idx shape")  # tf_sampling.cpp:136
    return _GatherPoint.apply(inp, idx)


def gather_point_grad(inp, idx, out_g):
    """The GatherPointGrad op (tf_sampling.cpp:55-63)."""
    inp = req(inp, torch.float32, 3, "inp")
    idx = req(idx, torch.int32, 2, "idx")
    out_g = req(out_g, torch.float32, 3, "out_g")
    b, n, _ = inp.shape
    m = idx.shape[1]
    inp_g = torch.empty((b, n, 3), dtype=torch.float32, device=inp.device)
    call("dispu_gather_point_grad", b, n, m, ptr(out_g), ptr(idx), ptr(inp_g), stream_ptr())
    return inp_g


def prob_sample(inp, inpr):
    """tf_sampling.py:13.  inp (b,ncategory) weights, inpr (b,npoints) uniform numbers -> (b,npoints) i32."""
    inp = req(inp, torch.float32, 2, "inp")
    inpr = req(inpr, torch.float32, 2, "inpr")
    b, n = inp.shape
    m = inpr.shape[1]
    if inpr.shape[0] != b:
        raise ValueError("ProbSample expects (batch_size,num_points) inpr shape")  # tf_sampling.cpp:77
    temp = torch.empty((b, n), dtype=torch.float32, device=inp.device)
    out = torch.empty((b, m), dtype=torch.int32, device=inp.device)
    call("dispu_prob_sample", b, n, m, ptr(inp), ptr(inpr), ptr(temp), ptr(out), stream_ptr())
    return out
