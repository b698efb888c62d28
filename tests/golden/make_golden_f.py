"""Golden vectors for the SURVEY §8(f) rows whose reference implementation is stock TensorFlow (no C++ to compile):
feature-space kNN (tf_grouping.py:61-66,95-114), dense_conv over get_edge_feature (Common/ops.py:1856-1915),
PointShuffle2's weight net + transposed product (Common/ops.py:181-191,1062-1067) and get_uniform_loss
(Common/loss_utils.py:238-269).  TensorFlow is not installable here, so the vectors come from INDEPENDENT numpy
float64 restatements of those expressions (written from the cited lines, not from oracle/), rounded to float32;
tests/test_oracle.py checks the C oracle against them and the -m gpu tests check the kernels against the oracle.

    python tests/golden/make_golden_f.py        -> tests/golden/f_rows.npz
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import unit_ball  # noqa: E402


def knn_point_2(k, points, queries):
    """tf_grouping.py:61-66 + :95-114 in float64: D = r_A - 2 A.B^T + r_B^T, top_k(-D)."""
    a, b = queries.astype(np.float64), points.astype(np.float64)
    r_a = (a * a).sum(2, keepdims=True)
    r_b = (b * b).sum(2, keepdims=True)
    d = r_a - 2 * np.einsum("bic,bjc->bij", a, b) + np.transpose(r_b, (0, 2, 1))
    idx = np.argsort(d, axis=-1, kind="stable")[:, :, :k]
    return np.take_along_axis(d, idx, -1), idx.astype(np.int32)


def dense_conv(feat, idx, ws, bs):
    """Common/ops.py:1897-1915 on get_edge_feature (:1856-1877), float64."""
    b, n, c = feat.shape
    k = idx.shape[2]
    f = feat.astype(np.float64)
    central = np.broadcast_to(f[:, :, None, :], (b, n, k, c))
    neigh = f[np.arange(b)[:, None, None], idx]
    y = np.concatenate([central, neigh - central], -1)
    y = np.concatenate([np.maximum(y @ ws[0] + bs[0], 0), central], -1)
    y = np.concatenate([np.maximum(y @ ws[1] + bs[1], 0), y], -1)
    y = np.concatenate([y @ ws[2] + bs[2], y], -1)
    return y.max(2)


def weightnet_aggregate(xyz, idx, G, Ww, bw, eps=1e-3):
    """Common/ops.py:1062-1067 with weight_net_hidden (:181-191, batch norm at its initial moving statistics)."""
    b = xyz.shape[0]
    gx = xyz[np.arange(b)[:, None, None], idx].astype(np.float64) - xyz[:, :, None, :].astype(np.float64)
    wn = np.maximum((gx @ Ww.astype(np.float64) + bw) / math.sqrt(1.0 + eps), 0)
    return np.einsum("bmsc,bmst->bmct", G.astype(np.float64), wn).reshape(b, xyz.shape[1], -1)


def uniform_term(pcd, idx, expect_len):
    """One percentage term of Common/loss_utils.py:253-266 given the ball-query idx, float64 on float32 distances."""
    b = pcd.shape[0]
    g = pcd[np.arange(b)[:, None, None], idx]  # (b,npoint,ns,3) float32
    diff = g[:, :, :, None, :] - g[:, :, None, :, :]
    d = ((diff * diff)[..., 0] + (diff * diff)[..., 1]) + (diff * diff)[..., 2]  # float32, unfused like tf.reduce_sum of 3 terms
    d1 = np.sort(d, -1)[..., 1].astype(np.float64)
    u = np.sqrt(np.abs(d1 + 1e-8))
    return float((((u - expect_len) ** 2) / (expect_len + 1e-8)).mean())


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    feat = rng.standard_normal((2, 120, 24)).astype(np.float32)
    out["knn_feat"] = feat
    d, i = knn_point_2(17, feat, feat)
    out["knn_d"], out["knn_i"] = d.astype(np.float32), i
    ws = [(rng.standard_normal((ci, 24)) / math.sqrt(ci)).astype(np.float32) for ci in (48, 48, 72)]
    bs = [(rng.standard_normal(24) / 4).astype(np.float32) for _ in range(3)]
    for j in range(3):
        out[f"dc_w{j}"], out[f"dc_b{j}"] = ws[j], bs[j]
    out["dc_idx"] = np.ascontiguousarray(i[:, :, 1:])
    out["dc_out"] = dense_conv(feat, out["dc_idx"], ws, bs).astype(np.float32)
    xyz = unit_ball(rng, 1, 24)
    d3 = ((xyz[:, :, None, :].astype(np.float64) - xyz[:, None, :, :]) ** 2).sum(-1)
    idx3 = np.argsort(d3, axis=-1, kind="stable")[:, :, :16].astype(np.int32)
    G = rng.standard_normal((1, 24, 16, 128)).astype(np.float32)
    Ww = rng.standard_normal((3, 16)).astype(np.float32)
    bw = (rng.standard_normal(16) / 4).astype(np.float32)
    out["wn_xyz"], out["wn_idx"], out["wn_G"], out["wn_Ww"], out["wn_bw"] = xyz, idx3, G, Ww, bw
    out["wn_out"] = weightnet_aggregate(xyz, idx3, G, Ww, bw).astype(np.float32)
    pcd = unit_ball(rng, 1, 1024)
    ns, p = 8, 0.008
    r = math.sqrt(p)
    centres = pcd[:, rng.permutation(1024)[:51]]
    dd = ((centres[:, :, None, :].astype(np.float64) - pcd[:, None, :, :]) ** 2).sum(-1)
    uidx = np.zeros((1, 51, ns), np.int32)
    for q in range(51):  # query_ball_point semantics: first ns in index order, padded with the first hit
        hits = np.nonzero(np.sqrt(dd[0, q]) < r * (1 - 1e-6))[0][:ns]
        row = np.full(ns, hits[0] if len(hits) else 0)
        row[:len(hits)] = hits
        uidx[0, q] = row
    out["un_pcd"], out["un_idx"] = pcd, uidx
    e = math.sqrt(math.pi * p / ns)
    out["un_expect"] = np.float32(e)
    out["un_term"] = np.float32(uniform_term(pcd, uidx, e))
    np.savez_compressed(os.path.join(HERE, "f_rows.npz"), **out)
    print("wrote f_rows.npz", {k: getattr(v, "shape", ()) for k, v in out.items()})


if __name__ == "__main__":
    main()
