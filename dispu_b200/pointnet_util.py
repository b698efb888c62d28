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


POOLINGS = {"max": 0, "avg": 1, "weighted_avg": 2, "min": 3, "max_and_avg": 4}  # DISPU_POOL_* of include/dispu_b200.h
BN_EPS = 1e-3  # tf.contrib.layers.batch_norm default epsilon (Common/tf_util.py:527)


def _split_layers(layers):
    """layers: list of (weight, bias) or (weight, bias, bn) with bn = (gamma, beta, moving_mean, moving_var) — the
    variables tf_util.conv2d(bn=True) creates (Common/tf_util.py:178-181, :527), applied in inference mode.
    Returns weights, biases, bn scales, bn shifts (folded: y = (x.W + b) * scale + shift)."""
    ws, bs, scs, shs = [], [], [], []
    for i, layer in enumerate(layers):
        w, bb = layer[0], layer[1]
        ws.append(req(w, torch.float32, 2, f"weight[{i}]"))
        bs.append(None if bb is None else req(bb, torch.float32, 1, f"bias[{i}]"))
        if len(layer) > 2 and layer[2] is not None:
            gamma, beta, mean, var = layer[2]
            sc = (gamma / torch.sqrt(var + BN_EPS)).to(torch.float32).contiguous()
            scs.append(sc)
            shs.append((beta - mean * sc).to(torch.float32).contiguous())
        else:
            scs.append(None)
            shs.append(None)
    return ws, bs, scs, shs


def _dense_relu(x, layers):
    """Post-pool / per-point 1x1 convs (tf_util.conv2d on a (b,m,1,C) tensor): relu(bn(x.W + b)) per layer on the
    tcgen05 dense layer (dense.PackedLinear, fp32-level 3xTF32); inference BN is folded into W and b exactly like
    the grouped kernel's affine."""
    from .dense import PackedLinear
    ws, bs, scs, shs = _split_layers(layers)
    for w, bb, sc, sh in zip(ws, bs, scs, shs):
        if sc is not None:
            w = (w * sc.unsqueeze(0)).contiguous()
            bb = sh if bb is None else (bb * sc + sh).contiguous()
        if w.shape[1] <= 256:
            x = PackedLinear(w, bb)(x.contiguous(), relu=True)
        else:
            raise NotImplementedError("dense layers wider than 256 columns are outside this path")
    return x


def grouped_mlp_max(xyz, points, new_xyz, idx, layers, layout=0, precision=MLP_FP32, pooling="max", use_xyz=True):
    """Fused group -> centre-subtract -> concat -> MLP -> pool over nsample (dispu_sa_mlp / dispu_sa_mlp_ex).
    layers: list of (weight (C_in,C_out), bias (C_out,) or None [, bn]).  Returns (b,m,C_last) (2*C_last for
    pooling='max_and_avg').  Forward only: when an input requires grad the unfused differentiable op sequence runs."""
    xyz = req(xyz, torch.float32, 3, "xyz")
    new_xyz = req(new_xyz, torch.float32, 3, "new_xyz")
    idx = req(idx, torch.int32, 3, "idx")
    if pooling not in POOLINGS:
        raise ValueError(f"unknown pooling {pooling!r}")
    b, n, _ = xyz.shape
    _, m, ns = idx.shape
    c_pts = 0
    if points is not None:
        points = req(points, torch.float32, 3, "points")
        c_pts = points.shape[2]
    if not use_xyz and points is None:
        use_xyz = True  # sample_and_group: without points the grouped xyz ARE the features (pointnet_util.py:57-58)
    ws, bs, scs, shs = _split_layers(layers)
    widths = [(3 if use_xyz else 0) + c_pts] + [w.shape[1] for w in ws]
    for i, w in enumerate(ws):
        if w.shape[0] != widths[i]:
            raise ValueError(f"weight[{i}] has C_in={w.shape[0]}, expected {widths[i]}")
    params = [t for t in (xyz, points, *ws, *bs) if t is not None]
    if torch.is_grad_enabled() and any(t.requires_grad for t in params):
        return _grouped_mlp_unfused(xyz, points, new_xyz, idx, ws, bs, scs, shs, layout, pooling, use_xyz)
    L = len(ws)
    widths_c = (C.c_int * (L + 1))(*widths)
    w_c = (C.c_void_p * L)(*[w.data_ptr() for w in ws])
    b_c = (C.c_void_p * L)(*[(bb.data_ptr() if bb is not None else None) for bb in bs])
    out_c = widths[-1] * (2 if pooling == "max_and_avg" else 1)
    out = torch.empty((b, m, out_c), dtype=torch.float32, device=xyz.device)
    plain = pooling == "max" and use_xyz and all(sc is None for sc in scs)
    if plain:
        call("dispu_sa_mlp", b, n, m, ns, c_pts, layout, ptr(xyz), ptr(points), ptr(new_xyz), ptr(idx), L,
             C.cast(widths_c, C.c_void_p), C.cast(w_c, C.c_void_p), C.cast(b_c, C.c_void_p), ptr(out), precision,
             stream_ptr())
    else:
        sc_c = (C.c_void_p * L)(*[(t.data_ptr() if t is not None else None) for t in scs])
        sh_c = (C.c_void_p * L)(*[(t.data_ptr() if t is not None else None) for t in shs])
        call("dispu_sa_mlp_ex", b, n, m, ns, c_pts, layout, 1 if use_xyz else 0, POOLINGS[pooling], ptr(xyz), ptr(points),
             ptr(new_xyz), ptr(idx), L, C.cast(widths_c, C.c_void_p), C.cast(w_c, C.c_void_p), C.cast(b_c, C.c_void_p),
             C.cast(sc_c, C.c_void_p), C.cast(sh_c, C.c_void_p), ptr(out), precision, stream_ptr())
    return out


def _grouped_mlp_unfused(xyz, points, new_xyz, idx, ws, bs, scs, shs, layout, pooling, use_xyz):
    """The reference's op sequence (pointnet_util.py:46-58, :117-141) on the differentiable GroupPoint op: used only
    when gradients are requested (training is outside the fused forward path)."""
    gx = group_point(xyz, idx) - new_xyz.unsqueeze(2)
    if points is not None:
        gp = group_point(points, idx)
        x = (torch.cat([gx, gp], -1) if layout == 0 else torch.cat([gp, gx], -1)) if use_xyz else gp
    else:
        x = gx
    for w, bb, sc, sh in zip(ws, bs, scs, shs):
        x = torch.matmul(x, w)
        if bb is not None:
            x = x + bb
        if sc is not None:
            x = x * sc + sh
        x = torch.relu(x)
    if pooling == "max":
        return x.max(dim=2)[0]
    if pooling == "avg":
        return x.mean(dim=2)
    if pooling == "min":
        return (-x).max(dim=2)[0]
    if pooling == "max_and_avg":
        return torch.cat([x.max(dim=2)[0], x.mean(dim=2)], dim=-1)
    wt = torch.exp(-5.0 * gx.norm(dim=-1, keepdim=True))
    return (x * (wt / wt.sum(dim=2, keepdim=True))).sum(dim=2)


def pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, mlp2=None, group_all=False, pooling="max",
                       knn=False, use_xyz=True, precision=MLP_FP32):
    """pointnet_util.py:91-149 with explicit parameters: `mlp` / `mlp2` are lists of (weight, bias[, bn]) of the
    per-point / per-region layers (the reference passes output widths and creates variables; bn = the inference
    statistics (gamma, beta, moving_mean, moving_var) of tf_util.conv2d(bn=True)).  All five pooling modes of
    :122-141 run inside the fused kernel; mlp2 (:142-147) runs on the tcgen05 dense layer.
    Returns (new_xyz, new_points (b,npoint,C), idx)."""
    if group_all:
        new_xyz, _, idx, _ = sample_and_group_all(xyz, points, use_xyz)
    else:
        new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
        if knn:
            _, idx = knn_point(nsample, xyz, new_xyz)
        else:
            idx, _ = query_ball_point(radius, nsample, xyz, new_xyz)
    new_points = grouped_mlp_max(xyz, points, new_xyz, idx, mlp, layout=0, precision=precision, pooling=pooling,
                                 use_xyz=use_xyz)
    if mlp2:
        new_points = _dense_relu(new_points, mlp2)
    return new_xyz, new_points, idx


def pointnet_sa_module_msg(xyz, points, npoint, radius_list, nsample_list, mlp_list, use_xyz=True,
                           precision=MLP_FP32):
    """pointnet_util.py:152-189: multi-scale grouping; channels are [points, xyz-centre] (:179)."""
    new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
    outs = []
    for radius, nsample, mlp in zip(radius_list, nsample_list, mlp_list):
        idx, _ = query_ball_point(radius, nsample, xyz, new_xyz)
        outs.append(grouped_mlp_max(xyz, points, new_xyz, idx, mlp, layout=1, precision=precision, use_xyz=use_xyz))
    return new_xyz, torch.cat(outs, dim=-1)


def pointnet_fp_module(xyz1, xyz2, points1, points2, mlp=None):
    """pointnet_util.py:192-222: three_nn -> inverse-distance weights -> three_interpolate ->
    concat with points1 -> `mlp` (list of (weight, bias[, bn])) as 1x1 convs with ReLU on the tcgen05 dense layer."""
    dist, idx = three_nn(xyz1, xyz2)
    weight = three_nn_weights(dist)
    interpolated = three_interpolate(points2, idx, weight)
    new_points1 = torch.cat([interpolated, points1], dim=2) if points1 is not None else interpolated
    if mlp:
        if torch.is_grad_enabled() and (new_points1.requires_grad or any(l[0].requires_grad for l in mlp)):
            for layer in mlp:
                w, bias = layer[0], layer[1]
                new_points1 = torch.relu(new_points1 @ w + (bias if bias is not None else 0.0))
        else:
            new_points1 = _dense_relu(new_points1, mlp)
    return new_points1
