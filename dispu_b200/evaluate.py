"""The CD / HD metrics of evaluate.py:34-39,153-165 on the fused NnDistance kernel: both clouds are
normalised to the unit sphere, then CD = mean fwd + mean bwd and HD = max fwd + max bwd of the squared
nearest-neighbour distances (nn_distance(pred, gt))."""
from __future__ import annotations

import torch

from . import tf_nndistance
from .inference import load_xyz, normalize_point_cloud


def chamfer_hausdorff_metric(pred, gt):
    """pred (n,3) / (b,n,3), gt (m,3) / (b,m,3) -> dict of python floats (b = 1) or (b,) tensors."""
    single = pred.dim() == 2
    if single:
        pred, gt = pred.unsqueeze(0), gt.unsqueeze(0)
    p, _, _ = normalize_point_cloud(pred)
    g, _, _ = normalize_point_cloud(gt)
    d_fwd, _, d_bwd, _ = tf_nndistance.nn_distance(p.contiguous(), g.contiguous())
    cd_f, cd_b = d_fwd.mean(dim=1), d_bwd.mean(dim=1)
    hd = d_fwd.amax(dim=1) + d_bwd.amax(dim=1)
    out = {"CD": cd_f + cd_b, "hausdorff": hd, "cd_forward": cd_f, "cd_backward": cd_b}
    return {k: float(v[0]) for k, v in out.items()} if single else out


def evaluate_pair(pred_path, gt_path, device="cuda"):
    """One row of evaluate.py's CSV (:153-165) from two .xyz files."""
    return chamfer_hausdorff_metric(load_xyz(pred_path, device), load_xyz(gt_path, device))
