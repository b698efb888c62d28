"""Mirror of the grouping helpers of Common/ops.py:110-179 that sit on the hot path."""
from __future__ import annotations

import torch

from . import tf_grouping, tf_sampling


def knn_query(k, support_pts, query_pts):
    """ops.py:110-118.  The reference calls nanoflann on the HOST through tf.py_func; this stays on the GPU.
    support_pts (b,n1,3), query_pts (b,n2,3) -> (b,n2,k) int32."""
    _, idx = tf_grouping.knn(k, support_pts, query_pts)
    return idx


def sampling(npoint, pts, feature=None):
    """ops.py:120-136.  FPS then gather of pts (b,n,3) [and feature (b,n,c)]."""
    fps_idx = tf_sampling.farthest_point_sample(npoint, pts)
    sub = tf_sampling.gather_point(pts, fps_idx)
    if feature is None:
        return sub
    return sub, tf_grouping.group_point(feature, fps_idx.unsqueeze(2)).squeeze(2)


def grouping(feature, K, src_xyz, q_xyz, use_xyz=True, use_knn=True, radius=0.2):
    """ops.py:154-179 -> (grouped_xyz (b,m,K,3), grouped_feature (b,m,K,[3+]c), idx (b,m,K) i32).
    The reference returns idx as (b,m,K,2) gather_nd pairs; the batch column is implicit here."""
    if use_knn:
        idx = knn_query(K, src_xyz, q_xyz)
    else:
        idx, _ = tf_grouping.query_ball_point(radius, K, src_xyz, q_xyz)
    grouped_xyz = tf_grouping.group_point(src_xyz, idx)
    grouped_feature = tf_grouping.group_point(feature, idx)
    if use_xyz:
        grouped_feature = torch.cat([grouped_xyz, grouped_feature], dim=-1)
    return grouped_xyz, grouped_feature, idx


def dilat_group(xyz, points, k, dilation=1, use_xyz=False):
    """ops.py:139-152."""
    _, idx = tf_grouping.knn_point(k * dilation + 1, xyz, xyz)
    idx = idx[:, :, 1::dilation].contiguous()
    grouped_xyz = tf_grouping.group_point(xyz, idx) - xyz.unsqueeze(2)
    if points is not None:
        grouped_points = tf_grouping.group_point(points, idx)
        if use_xyz:
            grouped_points = torch.cat([grouped_xyz, grouped_points], dim=-1)
    else:
        grouped_points = grouped_xyz
    return grouped_xyz, grouped_points, idx


def get_edge_feature_idx(point_cloud, k=16):
    """Neighbour indices of get_edge_feature (ops.py:1866-1868): knn_point_2(k+1) in feature space, self dropped."""
    return tf_grouping.feature_knn(k + 1, point_cloud, point_cloud, skip=1, need_dist=False)[1]


def dense_conv(feature, weights, k=16, idx=None, out=None):
    """dense_conv (ops.py:1897-1915) fused with get_edge_feature (:1856-1877) and the max over the k neighbours.
    feature (b,n,c) [a column slice is fine]; weights = ((w0,b0),(w1,b1),(w2,b2)) with w0 (2c,g), w1 (g+c,g),
    w2 (2g+c,g), g = 24 -> (y (b,n,3g+c), idx (b,n,k)).  `out` may be a column slice of a wider buffer."""
    from ._lib import lib
    from ._t import call, ptr, rows, stream_ptr
    (w0, b0), (w1, b1), (w2, b2) = weights
    b, n, c = feature.shape
    g = w0.shape[1]
    fp, _, ldf = rows(feature, "feature")
    if out is None:
        out = torch.empty((b, n, 3 * g + c), dtype=torch.float32, device=feature.device)
    op, om, ldo = rows(out, "out")
    if om != b * n or out.shape[-1] != 3 * g + c:
        raise ValueError("dense_conv: out must be (b,n,3g+c)")
    nbytes = int(lib().dispu_dense_conv_workspace_bytes(b, n))
    ws = tf_grouping._workspace(nbytes, feature.device)
    if idx is None:  # neighbour search fused with the projection (two launches for the whole block)
        idx = torch.empty((b, n, k), dtype=torch.int32, device=feature.device)
        call("dispu_dense_conv_knn", b, n, c, k, g, fp, ldf, ptr(idx), ptr(w0), ptr(b0), ptr(w1), ptr(b1), ptr(w2), ptr(b2),
             op, ldo, ptr(ws), ws.numel(), stream_ptr())
    else:
        call("dispu_dense_conv", b, n, c, idx.shape[2], g, fp, ldf, ptr(idx), ptr(w0), ptr(b0), ptr(w1), ptr(b1), ptr(w2),
             ptr(b2), op, ldo, ptr(ws), ws.numel(), stream_ptr())
    return out, idx


def weightnet_aggregate(xyz, idx, G, Ww, bw, bn_scale, out=None):
    """PointShuffle2's weight net + transposed product (ops.py:1062-1067): xyz (b,m,3), idx (b,m,ns), G (b,m,ns,C),
    Ww (3,T), bw (T) -> (b,m,C*T) = flatten(transpose(G) . relu(bn(gx.Ww+bw)))."""
    from ._t import call, ptr, req, stream_ptr
    xyz = req(xyz, torch.float32, 3, "xyz")
    idx = req(idx, torch.int32, 3, "idx")
    G = req(G, torch.float32, 4, "G")
    b, m, ns, C = G.shape
    T = Ww.shape[1]
    if out is None:
        out = torch.empty((b, m, C * T), dtype=torch.float32, device=G.device)
    call("dispu_weightnet_aggregate", b, m, ns, C, T, ptr(xyz), ptr(idx), ptr(G), ptr(Ww), ptr(bw), float(bn_scale),
         ptr(out), stream_ptr())
    return out
