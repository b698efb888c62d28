"""Mirror of tf_ops/grouping/tf_grouping.py on torch CUDA tensors, backed by the sm_100a kernels."""
from __future__ import annotations

import numbers

import torch

from . import _lib
from ._t import call, ptr, req, stream_ptr

_ws_cache = {}


def _workspace(nbytes, device):
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def query_ball_point(radius, nsample, xyz, new_xyz):
    """tf_grouping.py:9.  xyz (b,n,3) dataset, new_xyz (b,m,3) queries, radius scalar or (b,) tensor
    (only element 0 is used, as in tf_grouping_g.cu:25) -> idx (b,m,nsample) i32, pts_cnt (b,m) i32."""
    xyz = req(xyz, torch.float32, 3, "xyz")
    new_xyz = req(new_xyz, torch.float32, 3, "new_xyz")
    b, n, c = xyz.shape
    if c != 3:
        raise ValueError("QueryBallPoint expects (batch_size, ndataset, 3) xyz1 shape.")  # tf_grouping.cpp:80
    if new_xyz.shape[0] != b or new_xyz.shape[2] != 3:
        raise ValueError("QueryBallPoint expects (batch_size, npoint, 3) xyz2 shape.")  # tf_grouping.cpp:85
    if nsample <= 0:
        raise ValueError("QueryBallPoint expects positive nsample")  # tf_grouping.cpp:75
    m = new_xyz.shape[1]
    if isinstance(radius, numbers.Real):  # tf_grouping.py:21-26: scalar -> tiled (b,) tensor
        radius = torch.full((max(b, 1),), float(radius), dtype=torch.float32, device=xyz.device)
    radius = req(radius, torch.float32, 1, "radius")
    idx = torch.empty((b, m, nsample), dtype=torch.int32, device=xyz.device)
    cnt = torch.empty((b, m), dtype=torch.int32, device=xyz.device)
    if n >= 2048 and b > 0 and m > 0:  # large clouds: uniform-grid kernel (same results), scratch from the torch allocator
        nbytes = int(_lib.lib().dispu_query_ball_point_workspace_bytes(b, n, m))
        ws = _workspace(nbytes, xyz.device)
        call("dispu_query_ball_point_ws", b, n, m, ptr(radius), nsample, ptr(xyz), ptr(new_xyz), ptr(idx), ptr(cnt),
             ptr(ws), ws.numel(), stream_ptr())
    else:
        call("dispu_query_ball_point", b, n, m, ptr(radius), nsample, ptr(xyz), ptr(new_xyz), ptr(idx), ptr(cnt),
             stream_ptr())
    return idx, cnt


def select_top_k(k, dist):
    """tf_grouping.py:33.  dist (b,m,n) -> (idx (b,m,n) i32, dist_out (b,m,n)); first k columns are the k smallest."""
    dist = req(dist, torch.float32, 3, "dist")
    if k <= 0:
        raise ValueError("SelectionSort expects positive k")  # tf_grouping.cpp:119
    b, m, n = dist.shape
    outi = torch.empty((b, m, n), dtype=torch.int32, device=dist.device)
    out = torch.empty((b, m, n), dtype=torch.float32, device=dist.device)
    call("dispu_selection_sort", b, n, m, k, ptr(dist), ptr(outi), ptr(out), stream_ptr())
    return outi, out


class _GroupPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        b, n, c = points.shape
        _, m, ns = idx.shape
        out = torch.empty((b, m, ns, c), dtype=torch.float32, device=points.device)
        call("dispu_group_point", b, n, c, m, ns, ptr(points), ptr(idx), ptr(out), stream_ptr())
        ctx.save_for_backward(idx)
        ctx.shape = (b, n, c)
        return out

    @staticmethod
    def backward(ctx, grad_out):  # tf_grouping.py:53-57
        (idx,) = ctx.saved_tensors
        b, n, c = ctx.shape
        _, m, ns = idx.shape
        grad_out = grad_out.contiguous()
        g = torch.empty((b, n, c), dtype=torch.float32, device=grad_out.device)
        call("dispu_group_point_grad", b, n, c, m, ns, ptr(grad_out), ptr(idx), ptr(g), stream_ptr())
        return g, None


def group_point(points, idx):
    """tf_grouping.py:44.  points (b,n,c), idx (b,m,nsample) i32 -> (b,m,nsample,c); differentiable w.r.t. points."""
    points = req(points, torch.float32, 3, "points")
    idx = req(idx, torch.int32, 3, "idx")
    if idx.shape[0] != points.shape[0]:
        raise ValueError("GroupPoint expects (batch_size, npoints, nsample) idx shape")  # tf_grouping.cpp:160
    return _GroupPoint.apply(points, idx)


def group_point_grad(points, idx, grad_out):
    """The GroupPointGrad op (tf_grouping.cpp:56-64)."""
    points = req(points, torch.float32, 3, "points")
    idx = req(idx, torch.int32, 3, "idx")
    grad_out = req(grad_out, torch.float32, 4, "grad_out")
    b, n, c = points.shape
    _, m, ns = idx.shape
    g = torch.empty((b, n, c), dtype=torch.float32, device=points.device)
    call("dispu_group_point_grad", b, n, c, m, ns, ptr(grad_out), ptr(idx), ptr(g), stream_ptr())
    return g


def knn(k, xyz1, xyz2):
    """k nearest dataset points xyz1 (b,n,3) of every query xyz2 (b,m,3): (+d2 (b,m,k), idx (b,m,k) i32),
    ascending, ties to the lower index."""
    xyz1 = req(xyz1, torch.float32, 3, "xyz1")
    xyz2 = req(xyz2, torch.float32, 3, "xyz2")
    b, n, c = xyz1.shape
    if c != 3 or xyz2.shape[2] != 3:
        raise NotImplementedError("dispu_knn is the 3-D kernel; feature-space kNN is SURVEY §8(f) rank 3")
    if xyz2.shape[0] != b:
        raise ValueError("knn expects xyz1 and xyz2 to share the batch dimension")
    m = xyz2.shape[1]
    dist = torch.empty((b, m, k), dtype=torch.float32, device=xyz1.device)
    idx = torch.empty((b, m, k), dtype=torch.int32, device=xyz1.device)
    if k > 64:  # patch extraction sizes (k = 256): radix-select kernel, scratch from the torch allocator
        nbytes = int(_lib.lib().dispu_knn_large_workspace_bytes(b, n, m))
        ws = _workspace(nbytes, xyz1.device)
        call("dispu_knn_large", b, n, m, k, ptr(xyz1), ptr(xyz2), ptr(dist), ptr(idx), ptr(ws), ws.numel(), stream_ptr())
    else:
        call("dispu_knn", b, n, m, k, ptr(xyz1), ptr(xyz2), ptr(dist), ptr(idx), stream_ptr())
    return dist, idx


def knn_point(k, xyz1, xyz2):
    """tf_grouping.py:116.  Returns (val, idx) with val = NEGATED squared distances, like top_k(-dist)."""
    dist, idx = knn(k, xyz1, xyz2)
    return -dist, idx


def group_max(points, idx, out=None):
    """Fused group_point + reduce_max over nsample: points (b,n,c), idx (b,m,ns) -> (b,m,c).  Forward only.
    points / out may be column slices of wider buffers."""
    from ._t import rows
    idx = req(idx, torch.int32, 3, "idx")
    if points.dim() != 3:
        raise ValueError("group_max expects points (b,n,c)")
    b, n, c = points.shape
    _, m, ns = idx.shape
    pp, _, ldp = rows(points, "points")
    if out is None:
        out = torch.empty((b, m, c), dtype=torch.float32, device=points.device)
    op, om, ldo = rows(out, "out")
    if om != b * m or out.shape[-1] != c:
        raise ValueError("group_max: out must be (b,m,c)")
    call("dispu_group_max_ld", b, n, c, m, ns, pp, ldp, ptr(idx), op, ldo, stream_ptr())
    return out


def group_affine_relu(P, Q, idx, sign=-1.0):
    """out[b,i,s,:] = relu(P[b,idx[b,i,s],:] + sign*Q[b,i,:]): the first layer of a grouped MLP after hoisting it to
    per-point GEMMs (see dispu_group_affine_relu).  P (b,n,c), Q (b,m,c) or None, idx (b,m,ns) -> (b,m,ns,c)."""
    P = req(P, torch.float32, 3, "P")
    idx = req(idx, torch.int32, 3, "idx")
    if Q is not None:
        Q = req(Q, torch.float32, 3, "Q")
    b, n, c = P.shape
    _, m, ns = idx.shape
    out = torch.empty((b, m, ns, c), dtype=torch.float32, device=P.device)
    call("dispu_group_affine_relu", b, n, c, m, ns, ptr(P), ptr(Q), float(sign), ptr(idx), ptr(out), stream_ptr())
    return out


def feature_knn(k, points, queries, skip=0, need_dist=True):
    """Feature-space kNN on the matmul-form distance matrix (knn_point_2's arithmetic, tf_grouping.py:61-66,95-114):
    points (b,n,c), queries (b,m,c) [column slices of wider buffers are fine] -> (dist, idx) (b,m,k-skip), ascending,
    ties to the lower index; the first `skip` neighbours are dropped (get_edge_feature drops the self match,
    Common/ops.py:1868)."""
    from ._t import rows
    if points.dim() != 3 or queries.dim() != 3 or points.shape[0] != queries.shape[0] or points.shape[2] != queries.shape[2]:
        raise ValueError("feature_knn expects points (b,n,c) and queries (b,m,c)")
    b, n, c = points.shape
    m = queries.shape[1]
    if not (0 < k <= n) or not (0 <= skip < k):
        raise ValueError("feature_knn expects 0 < k <= n and 0 <= skip < k")
    pp, _, ldp = rows(points, "points")
    qp, _, ldq = rows(queries, "queries")
    idx = torch.empty((b, m, k - skip), dtype=torch.int32, device=points.device)
    dist = torch.empty((b, m, k - skip), dtype=torch.float32, device=points.device) if need_dist else None
    call("dispu_feature_knn", b, n, m, c, k, skip, pp, ldp, qp, ldq, ptr(dist), ptr(idx), stream_ptr())
    return dist, idx


def knn_point_2(k, points, queries, sort=True, unique=True):
    """tf_grouping.py:95-114.  Returns (distances (b,m,k), indices (b,m,k)); the reference returns the indices as
    (b,m,k,2) gather_nd pairs whose batch column is implicit here.  `unique` is a no-op in the reference too
    (prepare_for_unique_top_k :89-91 rebinds a local)."""
    return feature_knn(k, points, queries)
