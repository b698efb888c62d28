"""Mirror of the Chamfer / Hausdorff losses of Common/loss_utils.py:45-84, fused into the
NnDistance kernel (no (b,n) distance tensors are written, no TF reduce ops)."""
from __future__ import annotations

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


def chamfer_hausdorff(pred, gt, radius=None, out=None, per_cloud=None):
    """One fused pass: returns a (4,) device tensor [cd_mean, hd_max, cd_sum, b] (see dispu_chamfer).
    pred (b,n_pred,3), gt (b,n_gt,3), radius None | float | (b,) tensor."""
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
    call("dispu_chamfer", b, n_gt, ptr(gt), n_pred, ptr(pred), ptr(radius), ptr(out), ptr(per_cloud), ptr(ws),
         ws.numel(), stream_ptr())
    return out


def chamfer(pred, gt, radius=1.0, forward_weight=1.0, threshold=None, return_hd=False):
    """loss_utils.py:45.  cd = mean_b((mean fwd + mean bwd)/radius) with nn_distance(gt, pred)."""
    if forward_weight != 1.0 or threshold is not None:
        raise NotImplementedError("the live call sites (DisPU/model.py:75-79) use the defaults")
    return chamfer_hausdorff(pred, gt, None if radius == 1.0 else radius)[0]


def hausdorff_loss(pred, gt, radius=1.0, forward_weight=1.0, threshold=None):
    """loss_utils.py:67.  hd = max_b((max fwd + max bwd)/radius)."""
    if forward_weight != 1.0 or threshold is not None:
        raise NotImplementedError("the live call sites (DisPU/model.py:75-79) use the defaults")
    return chamfer_hausdorff(pred, gt, None if radius == 1.0 else radius)[1]
