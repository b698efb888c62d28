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
