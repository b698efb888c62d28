"""Mirror of the Chamfer / Hausdorff losses of Common/loss_utils.py:45-84, fused into the
NnDistance kernel (no (b,n) distance tensors are written, no TF reduce ops)."""
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
        ws = torch.empty((max(nbytes, 1 << 16),), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def chamfer_hausdorff(pred, gt, radius=None, out=None, per_cloud=None, sum_out=None, max_out=None):
    """One fused pass: returns a (4,) device tensor [cd_mean, hd_max, cd_sum, b] (see dispu_chamfer).
    pred (b,n_pred,3), gt (b,n_gt,3), radius None | float | (b,) tensor.  sum_out / max_out: optional 1-element float32
    views that also receive cd_sum / hd_max (e.g. slots of the buffer a single all-reduce sums, dispu_chamfer_ex)."""
    pred = req(pred, torch.float32, 3, "pred")
    gt = req(gt, torch.float32, 3, "gt")
    b, n_pred, c = pred.shape
    if c != 3 or gt.shape[2] != 3 or gt.shape[0] != b:
        raise ValueError("chamfer expects pred (b,n,3) and gt (b,m,3)")
    n_gt = gt.shape[1]
    if radius is not None and not isinstance(radius, torch.Tensor):
        radius = torch.full((b,), float(radius), dtype=torch.float32, device=pred.device)
    if radius is not None:
        radius = req(radius, torch.float32, 1, "radius")
    if out is None:
        out = torch.empty((4,), dtype=torch.float32, device=pred.device)
    nbytes = int(_lib.lib().dispu_chamfer_workspace_bytes(b, n_gt, n_pred))
    ws = _workspace(nbytes, pred.device)
    if sum_out is None and max_out is None:
        call("dispu_chamfer", b, n_gt, ptr(gt), n_pred, ptr(pred), ptr(radius), ptr(out), ptr(per_cloud), ptr(ws),
             ws.numel(), stream_ptr())
    else:
        for t in (sum_out, max_out):
            if t is not None and (t.dtype != torch.float32 or not t.is_cuda or t.numel() != 1):
                raise ValueError("sum_out / max_out must be 1-element float32 CUDA tensors (views are fine)")
        call("dispu_chamfer_ex", b, n_gt, ptr(gt), n_pred, ptr(pred), ptr(radius), ptr(out), ptr(per_cloud), ptr(sum_out),
             ptr(max_out), ptr(ws), ws.numel(), stream_ptr())
    return out


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


def _radius_arg(radius):
    """None for the reference's default 1.0 (Python scalar); scalars / (b,) tensors pass through.  A tensor is never
    compared with 1.0 (that would be a (b,) bool tensor and a hidden device->host sync)."""
    if isinstance(radius, numbers.Real) and float(radius) == 1.0:
        return None
    return radius


def _nn_pair_differentiable(pred, gt):
    """The reference's own graph (loss_utils.py:51): NnDistance op, gradients through dispu_nn_distance_grad
    (tf_nndistance.py:31-37)."""
    from .tf_nndistance import nn_distance
    fwd, _, bwd, _ = nn_distance(gt, pred)
    return fwd, bwd


def chamfer(pred, gt, radius=1.0, forward_weight=1.0, threshold=None, return_hd=False):
    """loss_utils.py:45.  cd = mean_b((mean fwd + mean bwd)/radius) with nn_distance(gt, pred); radius is a Python
    scalar or the per-cloud (b,) tensor the trainer passes (DisPU/model.py:75).  Inference / metric calls take the
    fused kernel (no distance tensors); when an input requires grad the loss is built from the differentiable
    NnDistance op exactly like the reference, so loss.backward() reaches pred through NnDistanceGrad."""
    if forward_weight != 1.0 or threshold is not None:
        raise NotImplementedError("the live call sites (DisPU/model.py:75-79) use the defaults")
    if _needs_grad(pred, gt):
        fwd, bwd = _nn_pair_differentiable(pred, gt)
        cd = fwd.mean(dim=1) + bwd.mean(dim=1)
        r = _radius_arg(radius)
        return (cd if r is None else cd / r).mean()
    return chamfer_hausdorff(pred, gt, _radius_arg(radius))[0]


def hausdorff_loss(pred, gt, radius=1.0, forward_weight=1.0, threshold=None):
    """loss_utils.py:67.  hd = max_b((max fwd + max bwd)/radius); same two paths as chamfer()."""
    if forward_weight != 1.0 or threshold is not None:
        raise NotImplementedError("the live call sites (DisPU/model.py:75-79) use the defaults")
    if _needs_grad(pred, gt):
        fwd, bwd = _nn_pair_differentiable(pred, gt)
        hd = fwd.max(dim=1)[0] + bwd.max(dim=1)[0]
        r = _radius_arg(radius)
        return (hd if r is None else hd / r).max()
    return chamfer_hausdorff(pred, gt, _radius_arg(radius))[1]


def get_repulsion_loss(pred, nsample=20, radius=0.07, knn=False, use_l1=False, h=0.001, out=None):
    """loss_utils.py:271-298 (the live call: DisPU/model.py:83, defaults).  pred (b,n,3) -> scalar tensor.
    Ball query on the B200 kernel, then one fused pass (no (b,n,nsample,3) tensor, no top_k).  When pred requires
    grad the reference's op sequence is used instead (QueryBallPoint -> differentiable GroupPoint -> top_k -> hinge)."""
    if knn or use_l1:
        raise NotImplementedError("the live call site uses the ball-query / squared-distance form")
    from .tf_grouping import group_point, query_ball_point
    pred = req(pred, torch.float32, 3, "pred")
    b, n, _ = pred.shape
    idx, _ = query_ball_point(radius, nsample, pred.detach(), pred.detach())
    if _needs_grad(pred):
        grouped = group_point(pred, idx) - pred.unsqueeze(2)
        dists = (grouped * grouped).sum(dim=-1)
        val = torch.topk(-dists, 5, dim=-1)[0][:, :, 1:]
        return torch.clamp(h + val, min=0.0).mean()
    if out is None:
        out = torch.empty((2,), dtype=torch.float32, device=pred.device)
    nbytes = int(_lib.lib().dispu_repulsion_workspace_bytes(b, n))
    ws = _workspace(nbytes, pred.device)
    call("dispu_repulsion_loss", b, n, nsample, float(h), ptr(pred), ptr(idx), ptr(out), ptr(ws), ws.numel(), stream_ptr())
    return out[0]


def get_uniform_loss(pcd, percentages=(0.004, 0.006, 0.008, 0.010, 0.012), radius=1.0):
    """loss_utils.py:238-269.  pcd (b,n,3) -> scalar tensor.  Per percentage: FPS(5 % of n) -> ball query
    (r = sqrt(p*radius), nsample = int(n*p)) -> one fused pass for the in-group nearest-neighbour statistic
    (no (b,npoint,nsample,3) tensor, no (nsample,nsample) distance matrices, no top_k)."""
    import math
    from .tf_grouping import query_ball_point
    from .tf_sampling import farthest_point_sample, gather_point
    pcd = req(pcd, torch.float32, 3, "pcd")
    b, n, _ = pcd.shape
    npoint = int(n * 0.05)
    if npoint <= 0:
        raise ValueError("get_uniform_loss needs at least 20 points per cloud")
    grad = _needs_grad(pcd)
    src = pcd.detach()
    new_xyz = gather_point(src, farthest_point_sample(npoint, src))  # the same seeds for every percentage
    total = None
    for p in percentages:
        nsample = int(n * p)
        if nsample <= 0:
            raise ValueError("get_uniform_loss: int(n*p) must be positive")
        r = math.sqrt(p * radius)
        disk_area = math.pi * (radius ** 2) * p / nsample
        expect_len = math.sqrt(disk_area)
        idx, _ = query_ball_point(r, nsample, src, new_xyz)
        if grad:  # the reference's op sequence on the differentiable GroupPoint (loss_utils.py:254-262)
            from .tf_grouping import group_point
            g = group_point(pcd, idx)  # (b,npoint,nsample,3)
            d = ((g.unsqueeze(3) - g.unsqueeze(2)) ** 2).sum(dim=-1)  # knn_point(2, g, g): pairwise squared distances
            var = torch.topk(-d, 2, dim=-1)[0]
            ud = torch.sqrt(torch.abs(-var[..., 1:] + 1e-8)).mean(dim=-1)
            ud = (ud - expect_len) ** 2 / (expect_len + 1e-8)
            term = ud.reshape(-1).mean() * math.pow(p * 100, 2)
            total = term if total is None else total + term
            continue
        out = torch.empty((2,), dtype=torch.float32, device=pcd.device)
        nbytes = int(_lib.lib().dispu_uniform_workspace_bytes(b, npoint, nsample))
        ws = _workspace(nbytes, pcd.device)
        call("dispu_uniform_loss_term", b, n, npoint, nsample, float(expect_len), ptr(pcd), ptr(idx), ptr(out), ptr(ws),
             ws.numel(), stream_ptr())
        term = out[0] * math.pow(p * 100, 2)
        total = term if total is None else total + term
    return total / len(percentages)
