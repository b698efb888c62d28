"""numpy front-end of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (dispu_b200/) never does.

Each function mirrors the argument order of the reference Python wrapper it stands for
(tf_ops/*/tf_*.py) and returns numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

GPU = 0  # contraction of the reference CUDA kernels (fmaf chain)
CPU = 1  # contraction of the reference host loops (unfused)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dispu_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def farthest_point_sample(npoint, inp):
    """tf_sampling.py:48 — (npoint, inp[b,n,3]) -> idx[b,npoint] int32."""
    inp = _f(inp)
    b, n, _ = inp.shape
    out = np.empty((b, npoint), np.int32)
    rc = lib().orc_fps(b, n, npoint, _p(inp), _p(out))
    assert rc == 0, rc
    return out


def gather_point(inp, idx):
    """tf_sampling.py:29."""
    inp, idx = _f(inp), _i(idx)
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = np.empty((b, m, 3), np.float32)
    lib().orc_gather_point(b, n, m, _p(inp), _p(idx), _p(out))
    return out


def gather_point_grad(inp, idx, out_g):
    inp, idx, out_g = _f(inp), _i(idx), _f(out_g)
    b, n, _ = inp.shape
    m = idx.shape[1]
    g = np.empty((b, n, 3), np.float32)
    lib().orc_gather_point_grad(b, n, m, _p(out_g), _p(idx), _p(g))
    return g


def query_ball_point(radius, nsample, xyz, new_xyz, mode=GPU, fill_empty=True):
    """tf_grouping.py:9 — returns (idx[b,m,nsample], pts_cnt[b,m])."""
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    b, n, _ = xyz.shape
    m = new_xyz.shape[1]
    r = np.full((b,), radius, np.float32) if np.isscalar(radius) else _f(radius)
    idx = np.full((b, m, nsample), -1, np.int32)
    cnt = np.empty((b, m), np.int32)
    lib().orc_query_ball_point(mode, b, n, m, _p(r), nsample, _p(xyz), _p(new_xyz), _p(idx), _p(cnt),
                               1 if fill_empty else 0)
    return idx, cnt


def group_point(points, idx):
    """tf_grouping.py:44."""
    points, idx = _f(points), _i(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, m, ns, c), np.float32)
    lib().orc_group_point(b, n, c, m, ns, _p(points), _p(idx), _p(out))
    return out


def group_point_grad(points, idx, grad_out):
    points, idx, grad_out = _f(points), _i(idx), _f(grad_out)
    b, n, c = points.shape
    _, m, ns = idx.shape
    g = np.empty((b, n, c), np.float32)
    lib().orc_group_point_grad(b, n, c, m, ns, _p(grad_out), _p(idx), _p(g))
    return g


def knn(k, data, query):
    """+d2 ascending, ties -> lower index.  Returns (dist[b,m,k], idx[b,m,k])."""
    data, query = _f(data), _f(query)
    b, n, c = data.shape
    m = query.shape[1]
    dist = np.empty((b, m, k), np.float32)
    idx = np.empty((b, m, k), np.int32)
    rc = lib().orc_knn(b, n, m, c, k, _p(data), _p(query), _p(dist), _p(idx))
    assert rc == 0, rc
    return dist, idx


def knn_point(k, xyz1, xyz2):
    """tf_grouping.py:116 — returns (val = NEGATED squared distances, idx)."""
    d, i = knn(k, xyz1, xyz2)
    return -d, i


def three_nn(xyz1, xyz2):
    """tf_interpolate.py:8 — (dist[b,n,3] squared, idx[b,n,3])."""
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.empty((b, n, 3), np.float32)
    idx = np.empty((b, n, 3), np.int32)
    lib().orc_three_nn(b, n, m, _p(xyz1), _p(xyz2), _p(dist), _p(idx))
    return dist, idx


def three_interpolate(points, idx, weight):
    """tf_interpolate.py:19."""
    points, idx, weight = _f(points), _i(idx), _f(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.empty((b, n, c), np.float32)
    lib().orc_three_interpolate(b, m, c, n, _p(points), _p(idx), _p(weight), _p(out))
    return out


def three_interpolate_grad(points, idx, weight, grad_out):
    points, idx, weight, grad_out = _f(points), _i(idx), _f(weight), _f(grad_out)
    b, m, c = points.shape
    n = idx.shape[1]
    g = np.empty((b, m, c), np.float32)
    lib().orc_three_interpolate_grad(b, n, c, m, _p(grad_out), _p(idx), _p(weight), _p(g))
    return g


def nn_distance(xyz1, xyz2, mode=GPU):
    """tf_nndistance.py:14 — (dist1, idx1, dist2, idx2)."""
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.empty((b, n), np.float32)
    i1 = np.empty((b, n), np.int32)
    d2 = np.empty((b, m), np.float32)
    i2 = np.empty((b, m), np.int32)
    rc = lib().orc_nn_distance(mode, b, n, m, _p(xyz1), _p(xyz2), _p(d1), _p(i1), _p(d2), _p(i2))
    assert rc == 0, rc
    return d1, i1, d2, i2


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    g1, g2, idx1, idx2 = _f(grad_dist1), _f(grad_dist2), _i(idx1), _i(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    o1 = np.empty((b, n, 3), np.float32)
    o2 = np.empty((b, m, 3), np.float32)
    lib().orc_nn_distance_grad(b, n, m, _p(xyz1), _p(xyz2), _p(g1), _p(idx1), _p(g2), _p(idx2), _p(o1), _p(o2))
    return o1, o2


def chamfer(pred, gt, radius=None, mode=GPU):
    """loss_utils.py:45 + :67 — returns dict(cd, hd, per_cd, per_hd); note gt goes first
    into nn_distance."""
    pred, gt = _f(pred), _f(gt)
    b, npred, _ = pred.shape
    ngt = gt.shape[1]
    r = None if radius is None else _f(np.broadcast_to(np.asarray(radius, np.float32), (b,)))
    cd = C.c_float()
    hd = C.c_float()
    pcd = np.empty((b,), np.float32)
    phd = np.empty((b,), np.float32)
    rc = lib().orc_chamfer(mode, b, ngt, npred, _p(gt), _p(pred), _p(r), C.byref(cd), C.byref(hd), _p(pcd), _p(phd))
    assert rc == 0, rc
    return {"cd": np.float32(cd.value), "hd": np.float32(hd.value), "per_cd": pcd, "per_hd": phd}


def sa_mlp(xyz, points, new_xyz, idx, weights, biases, layout=0):
    """Grouped MLP + max-pool of pointnet_sa_module (pointnet_util.py:117-133) on the tensor
    sample_and_group builds (:41-59).  weights[l] is (C_in, C_out)."""
    xyz, new_xyz, idx = _f(xyz), _f(new_xyz), _i(idx)
    b, n, _ = xyz.shape
    _, m, ns = idx.shape
    c_pts = 0 if points is None else points.shape[2]
    pts = None if points is None else _f(points)
    ws = [_f(w) for w in weights]
    bs = [None if bb is None else _f(bb) for bb in biases]
    widths = np.array([3 + c_pts] + [w.shape[1] for w in ws], np.int32)
    L = len(ws)
    wp = (C.c_void_p * L)(*[w.ctypes.data for w in ws])
    bp = (C.c_void_p * L)(*[(bb.ctypes.data if bb is not None else None) for bb in bs])
    out = np.empty((b, m, int(widths[-1])), np.float32)
    rc = lib().orc_sa_mlp(b, n, m, ns, c_pts, layout, _p(xyz), _p(pts), _p(new_xyz), _p(idx), L, _p(widths),
                          wp, bp, _p(out))
    assert rc == 0, rc
    return out


def repulsion_loss(pred, idx, h=0.001):
    """loss_utils.py:271-298 given the ball-query idx (b,n,nsample)."""
    pred, idx = _f(pred), _i(idx)
    b, n, _ = pred.shape
    out = C.c_float()
    lib().orc_repulsion_loss.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib().orc_repulsion_loss(b, n, idx.shape[2], float(h), _p(pred), _p(idx), C.byref(out))
    assert rc == 0, rc
    return np.float32(out.value)


def selection_sort(dist, k):
    dist = _f(dist)
    b, m, n = dist.shape
    outi = np.empty((b, m, n), np.int32)
    out = np.empty((b, m, n), np.float32)
    lib().orc_selection_sort(b, n, m, k, _p(dist), _p(outi), _p(out))
    return outi, out


# ---- compositions (pure numpy on top of the C oracle) ---------------------------------------

def sample_and_group(npoint, radius, nsample, xyz, points, knn_mode=False, use_xyz=True):
    """pointnet_util.py:22-59 -> (new_xyz, new_points, idx, grouped_xyz)."""
    xyz = _f(xyz)
    new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
    if knn_mode:
        _, idx = knn_point(nsample, xyz, new_xyz)
    else:
        idx, _ = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = group_point(xyz, idx) - new_xyz[:, :, None, :]
    if points is not None:
        gp = group_point(points, idx)
        new_points = np.concatenate([grouped_xyz, gp], -1) if use_xyz else gp
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def sample_and_group_all(xyz, points, use_xyz=True):
    """pointnet_util.py:62-88 -> (new_xyz zeros (b,1,3), new_points (b,1,n,3+c), idx (b,1,n), grouped_xyz)."""
    xyz = _f(xyz)
    b, n, _ = xyz.shape
    new_xyz = np.zeros((b, 1, 3), np.float32)
    idx = np.tile(np.arange(n, dtype=np.int32).reshape(1, 1, n), (b, 1, 1))
    grouped_xyz = xyz.reshape(b, 1, n, 3)
    if points is not None:
        new_points = (np.concatenate([xyz, _f(points)], 2) if use_xyz else _f(points))[:, None]
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def conv2d_1x1(x, w, b=None, bn=None, relu=True):
    """tf_util.conv2d with a [1,1] kernel (Common/tf_util.py:120-185): x.W + bias, optional inference batch norm
    (tf.nn.batch_normalization with the moving statistics, epsilon 1e-3, :527), ReLU.  float64 accumulate, so the
    result is the exact value the fp32 implementations are compared against."""
    y = x.astype(np.float64) @ w.astype(np.float64)
    if b is not None:
        y = y + b.astype(np.float64)
    if bn is not None:
        gamma, beta, mean, var = (np.asarray(t, np.float64) for t in bn)
        y = (y - mean) / np.sqrt(var + 1e-3) * gamma + beta
    return np.maximum(y, 0.0) if relu else y


def sa_module(xyz, points, new_xyz, idx, layers, pooling="max", use_xyz=True, layout=0, mlp2=None):
    """pointnet_sa_module's body after sampling/grouping (pointnet_util.py:117-149) and the _msg channel order (:179),
    literally: grouped tensor -> conv2d per layer -> pooling variant -> mlp2.  layers: list of (w, b[, bn]).
    Returns float64 (b,m,C)."""
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    gx = group_point(xyz, idx) - new_xyz[:, :, None, :]
    if points is not None:
        gp = group_point(_f(points), idx)
        if use_xyz:
            x = np.concatenate([gx, gp], -1) if layout == 0 else np.concatenate([gp, gx], -1)
        else:
            x = gp
    else:
        x = gx
    x = x.astype(np.float64)
    for layer in layers:
        x = conv2d_1x1(x, layer[0], layer[1], layer[2] if len(layer) > 2 else None)
    ns = x.shape[2]
    if pooling == "avg":
        x = x.sum(2, keepdims=True) / ns
    elif pooling == "weighted_avg":  # :124-131
        d = np.sqrt((gx.astype(np.float64) ** 2).sum(-1, keepdims=True))
        e = np.exp(-d * 5)
        x = (x * (e / e.sum(2, keepdims=True))).sum(2, keepdims=True)
    elif pooling == "max":
        x = x.max(2, keepdims=True)
    elif pooling == "min":  # :135-137: max_pool2d(-1 * new_points): the result is NOT negated back
        x = (-x).max(2, keepdims=True)
    elif pooling == "max_and_avg":  # :138-141 (the reference's variable names are swapped; the concat is [max | avg])
        x = np.concatenate([x.max(2, keepdims=True), x.sum(2, keepdims=True) / ns], -1)
    else:
        raise ValueError(pooling)
    for layer in (mlp2 or []):
        x = conv2d_1x1(x, layer[0], layer[1], layer[2] if len(layer) > 2 else None)
    return x[:, :, 0]


def three_nn_weights(dist):
    """pointnet_util.py:205-208."""
    d = np.maximum(_f(dist), np.float32(1e-10))
    inv = (np.float32(1.0) / d).astype(np.float32)
    norm = inv.sum(axis=2, keepdims=True, dtype=np.float32)
    return (inv / norm).astype(np.float32)


# ---- feature-space kNN / EdgeConv dense block / weight-net aggregation ------------

def feature_knn(k, points, queries, skip=0):
    """knn_point_2 (tf_grouping.py:95-114): (k, points (b,n,c), queries (b,m,c)) -> (dist, idx) (b,m,k-skip)."""
    points, queries = _f(points), _f(queries)
    b, n, c = points.shape
    m = queries.shape[1]
    dist = np.empty((b, m, k), np.float32)
    idx = np.empty((b, m, k), np.int32)
    rc = lib().orc_feature_knn(b, n, m, c, k, _p(points), c, _p(queries), c, _p(dist), _p(idx))
    assert rc == 0, rc
    return np.ascontiguousarray(dist[:, :, skip:]), np.ascontiguousarray(idx[:, :, skip:])


def dense_conv(feat, idx, w0, b0, w1, b1, w2, b2):
    """dense_conv (Common/ops.py:1897-1915), literal op order -> (b,n,3g+c)."""
    feat, idx = _f(feat), _i(idx)
    b, n, c = feat.shape
    k = idx.shape[2]
    g = w0.shape[1]
    ws = [_f(w0), _f(w1), _f(w2)]
    bs = [None if x is None else _f(x) for x in (b0, b1, b2)]
    out = np.empty((b, n, 3 * g + c), np.float32)
    rc = lib().orc_dense_conv(b, n, c, k, g, _p(feat), c, _p(idx), _p(ws[0]), _p(bs[0]), _p(ws[1]), _p(bs[1]),
                              _p(ws[2]), _p(bs[2]), _p(out), 3 * g + c)
    assert rc == 0, rc
    return out


def weightnet_aggregate(xyz, idx, G, Ww, bw, bn_scale):
    """Common/ops.py:1062-1067 -> (b,m,C*T)."""
    xyz, idx, G, Ww = _f(xyz), _i(idx), _f(G), _f(Ww)
    bw = None if bw is None else _f(bw)
    b, m, ns, Cc = G.shape
    T = Ww.shape[1]
    out = np.empty((b, m, Cc * T), np.float32)
    lib().orc_weightnet_aggregate.argtypes = [C.c_int] * 5 + [C.c_void_p] * 5 + [C.c_float, C.c_void_p]
    rc = lib().orc_weightnet_aggregate(b, m, ns, Cc, T, _p(xyz), _p(idx), _p(G), _p(Ww), _p(bw), float(bn_scale), _p(out))
    assert rc == 0, rc
    return out


def uniform_loss_term(pcd, idx, expect_len):
    """One percentage term of get_uniform_loss (loss_utils.py:238-269) given the ball-query idx (b,npoint,nsample)."""
    pcd, idx = _f(pcd), _i(idx)
    b, n, _ = pcd.shape
    out = C.c_float()
    lib().orc_uniform_loss_term.argtypes = [C.c_int] * 4 + [C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib().orc_uniform_loss_term(b, n, idx.shape[1], idx.shape[2], float(expect_len), _p(pcd), _p(idx), C.byref(out))
    assert rc == 0, rc
    return np.float32(out.value)


def get_uniform_loss(pcd, percentages=(0.004, 0.006, 0.008, 0.010, 0.012), radius=1.0):
    """loss_utils.py:238-269 composed from the oracle ops."""
    import math
    pcd = _f(pcd)
    b, n, _ = pcd.shape
    npoint = int(n * 0.05)
    new_xyz = gather_point(pcd, farthest_point_sample(npoint, pcd))
    tot = 0.0
    for p in percentages:
        nsample = int(n * p)
        r = math.sqrt(p * radius)
        expect_len = math.sqrt(math.pi * radius ** 2 * p / nsample)
        idx, _ = query_ball_point(r, nsample, pcd, new_xyz)
        tot += float(uniform_loss_term(pcd, idx, expect_len)) * math.pow(p * 100, 2)
    return np.float32(tot / len(percentages))
