"""Mirror of Common/pointnet_util.py (PointNet++ set abstraction / feature propagation) on torch
CUDA tensors.  The reference builds TF variables inside the call; here the layer parameters are
passed in explicitly as lists of (weight (C_in,C_out), bias (C_out,) | None) — weight is the
(1,1,C_in,C_out) kernel of tf_util.conv2d (Common/tf_util.py:120-185) with the unit dims dropped.
Batch norm is not on this path (the live call sites run bn=False / inference); ReLU after every
layer as in tf_util.conv2d's default activation.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._t import call, ptr, req, stream_ptr
from .tf_grouping import group_point, knn_point, query_ball_point
from .tf_interpolate import three_interpolate, three_nn, three_nn_weights
from .tf_sampling import farthest_point_sample, gather_point

MLP_FP32, MLP_TF32X3, MLP_TF32 = 0, 1, 2  # exact fp32 | tcgen05 3xTF32 (fp32-level) | tcgen05 single-pass TF32


def sample_and_group(npoint, radius, nsample, xyz, points, knn=False, use_xyz=True):
    """pointnet_util.py:22-59 -> (new_xyz (b,npoint,3), new_points (b,npoint,nsample,3+c), idx, grouped_xyz).
    This materialises the grouped tensors like the reference; pointnet_sa_module does not."""
    new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
    if knn:
        _, idx = knn_point(nsample, xyz, new_xyz)
    else:
        idx, _ = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = group_point(xyz, idx) - new_xyz.unsqueeze(2)
    if points is not None:
        grouped_points = group_point(points, idx)
        new_points = torch.cat([grouped_xyz, grouped_points], dim=-1) if use_xyz else grouped_points
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def sample_and_group_all(xyz, points, use_xyz=True):
    """pointnet_util.py:62-88."""
    b, n, _ = xyz.shape
    new_xyz = torch.zeros((b, 1, 3), dtype=torch.float32, device=xyz.device)
    idx = torch.arange(n, dtype=torch.int32, device=xyz.device).view(1, 1, n).repeat(b, 1, 1)
    grouped_xyz = xyz.view(b, 1, n, 3)
    if points is not None:
        new_points = torch.cat([xyz, points], dim=2) if use_xyz else points
        new_points = new_points.unsqueeze(1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def grouped_mlp_max(xyz, points, new_xyz, idx, layers, layout=0, precision=MLP_FP32):
    """Fused group -> centre-subtract -> concat -> MLP -> max over nsample (dispu_sa_mlp).
    layers: list of (weight (C_in,C_out), bias (C_out,) or None).  Returns (b,m,C_last)."""
    xyz = req(xyz, torch.float32, 3, "xyz")
    new_xyz = req(new_xyz, torch.float32, 3, "new_xyz")
    idx = req(idx, torch.int32, 3, "idx")
    b, n, _ = xyz.shape
    _, m, ns = idx.shape
    c_pts = 0
    if points is not None:
        points = req(points, torch.float32, 3, "points")
        c_pts = points.shape[2]
    ws = [req(w, torch.float32, 2, f"weight[{i}]") for i, (w, _) in enumerate(layers)]
    bs = [None if bb is None else req(bb, torch.float32, 1, f"bias[{i}]") for i, (_, bb) in enumerate(layers)]
    widths = [3 + c_pts] + [w.shape[1] for w in ws]
    for i, w in enumerate(ws):
        if w.shape[0] != widths[i]:
            raise ValueError(f"weight[{i}] has C_in={w.shape[0]}, expected {widths[i]}")
    L = len(ws)
    widths_c = (C.c_int * (L + 1))(*widths)
    w_c = (C.c_void_p * L)(*[w.data_ptr() for w in ws])
    b_c = (C.c_void_p * L)(*[(bb.data_ptr() if bb is not None else None) for bb in bs])
    out = torch.empty((b, m, widths[-1]), dtype=torch.float32, device=xyz.device)
    call("dispu_sa_mlp", b, n, m, ns, c_pts, layout, ptr(xyz), ptr(points), ptr(new_xyz), ptr(idx), L,
         C.cast(widths_c, C.c_void_p), C.cast(w_c, C.c_void_p), C.cast(b_c, C.c_void_p), ptr(out), precision,
         stream_ptr())
    return out


def pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, mlp2=None, group_all=False, pooling="max",
                       knn=False, use_xyz=True, precision=MLP_FP32):
    """pointnet_util.py:91-149 with explicit parameters: `mlp` is the list of (weight, bias) of the
    per-point layers (the reference passes output widths and creates variables).  max pooling
    (the default and the only live choice); returns (new_xyz, new_points (b,npoint,C), idx)."""
    if pooling != "max":
        raise NotImplementedError("only pooling='max' is on the hot path (pointnet_util.py:133)")
    if mlp2:
        raise NotImplementedError("mlp2 (post-pool 1x1 convs) is a plain dense layer, outside the fused path")
    if not use_xyz and points is not None:
        raise NotImplementedError("use_xyz=False is not used by any live call site")
    if group_all:
        new_xyz, _, idx, _ = sample_and_group_all(xyz, points, use_xyz)
    else:
        new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
        if knn:
            _, idx = knn_point(nsample, xyz, new_xyz)
        else:
            idx, _ = query_ball_point(radius, nsample, xyz, new_xyz)
    new_points = grouped_mlp_max(xyz, points, new_xyz, idx, mlp, layout=0, precision=precision)
    return new_xyz, new_points, idx


def pointnet_sa_module_msg(xyz, points, npoint, radius_list, nsample_list, mlp_list, use_xyz=True,
                           precision=MLP_FP32):
    """pointnet_util.py:152-189: multi-scale grouping; channels are [points, xyz-centre] (:179)."""
    new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
    outs = []
    for radius, nsample, mlp in zip(radius_list, nsample_list, mlp_list):
        idx, _ = query_ball_point(radius, nsample, xyz, new_xyz)
        outs.append(grouped_mlp_max(xyz, points, new_xyz, idx, mlp, layout=1, precision=precision))
    return new_xyz, torch.cat(outs, dim=-1)


def pointnet_fp_module(xyz1, xyz2, points1, points2, mlp=None):
    """pointnet_util.py:192-222: three_nn -> inverse-distance weights -> three_interpolate ->
    concat with points1.  `mlp` (list of (weight, bias)) is applied as plain dense ReLU layers."""
    dist, idx = three_nn(xyz1, xyz2)
    weight = three_nn_weights(dist)
    interpolated = three_interpolate(points2, idx, weight)
    new_points1 = torch.cat([interpolated, points1], dim=2) if points1 is not None else interpolated
    for w, bias in (mlp or []):
        new_points1 = torch.relu(new_points1 @ w + (bias if bias is not None else 0.0))
    return new_points1
