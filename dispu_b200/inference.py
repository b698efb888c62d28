"""The inference patch pipeline of DisPU/model.py:315-381 (Model.pc_prediction / Model.test) on the B200 kernels
 and the on-disk .xyz format it reads and writes.

The reference: FPS patch seeds on the GPU -> .eval() to the host -> sklearn kNN patches on the CPU
(Common/pc_util.py:83-91) -> ~24 serial B=1 sess.run calls -> concatenate -> one-CTA FPS down to the target count ->
np.savetxt.  Here every step stays on the device and the patches go through the generator as ONE batch:
seed FPS (cluster kernel for large clouds) -> large-k kNN patch extraction -> per-patch normalisation ->
batched generator -> de-normalise -> merge -> FPS to the target count.
"""
from __future__ import annotations

import numpy as np
import torch

from . import tf_grouping, tf_sampling


def normalize_point_cloud(pc):
    """Common/pc_util.py:147-161 on a torch tensor (..., P, 3): (pc, centroid, furthest_distance)."""
    centroid = pc.mean(dim=-2, keepdim=True)
    pc = pc - centroid
    furthest = torch.sqrt((pc * pc).sum(dim=-1, keepdim=True)).amax(dim=-2, keepdim=True)
    return pc / furthest, centroid, furthest


def extract_knn_patch(queries, pc, k):
    """Common/pc_util.py:83-91: queries (M,3), pc (P,3) -> (patches (M,k,3), idx (M,k) int32); neighbours ascending
    in distance, ties to the lower index (sklearn's brute-force order is unspecified for exact ties)."""
    _, idx = tf_grouping.knn(k, pc.unsqueeze(0).contiguous(), queries.unsqueeze(0).contiguous())
    idx = idx[0]
    return pc[idx.long()], idx


def pc_prediction(gen, pc, patch_num_point=256, patch_num_ratio=3, passes=1):
    """Model.pc_prediction (DisPU/model.py:315-341) with Model.patch_prediction (:300-313) batched over the patches.
    pc (N,3) float32 CUDA tensor, already normalised like Model.test does (:365).  `passes` = how many times the
    generator is applied (build_model_test, :116-118: 1 for x4, 2 for x16).
    Returns (patches (S,k,3), pred (S,k*4^passes,3), coarse (S,k*4^passes,3)) in the cloud's frame."""
    n = pc.shape[0]
    seed_num = int(n / patch_num_point * patch_num_ratio)
    seeds = tf_sampling.farthest_point_sample(seed_num, pc.unsqueeze(0).contiguous())[0]
    patches, _ = extract_knn_patch(pc[seeds.long()], pc, patch_num_point)
    x, centroid, furthest = normalize_point_cloud(patches)
    x = x.contiguous()
    with torch.no_grad():
        coarse, fine = gen(x)
        for _ in range(passes - 1):
            coarse, fine = gen(fine.contiguous())
    pred = centroid + fine * furthest
    pred_coarse = centroid + coarse * furthest
    return patches, pred, pred_coarse


def upsample(gen, pc, final_ratio=4, patch_num_point=256, patch_num_ratio=3):
    """Model.test for one cloud (DisPU/model.py:361-381): pc (N,3) -> (N*final_ratio, 3), both float32 CUDA tensors."""
    step_ratio = 4
    passes = max(1, int(round(np.log(final_ratio) / np.log(step_ratio))))
    pcn, centroid, furthest = normalize_point_cloud(pc)
    _, pred, _ = pc_prediction(gen, pcn.contiguous(), patch_num_point, patch_num_ratio, passes)
    merged = (pred.reshape(-1, 3) * furthest) + centroid
    out_num = int(pc.shape[0] * final_ratio)
    idx = tf_sampling.farthest_point_sample(out_num, merged.unsqueeze(0).contiguous())[0]
    return merged[idx.long()]


def save_xyz(path, points):
    """np.savetxt(path, pred_pc, fmt='%.6f') (DisPU/model.py:381)."""
    np.savetxt(path, points.detach().cpu().numpy() if isinstance(points, torch.Tensor) else points, fmt="%.6f")


def load_xyz(path, device="cuda"):
    """pc_util.load for .xyz / .txt (Common/pc_util.py:389-392): np.loadtxt -> float32, first three columns."""
    return torch.from_numpy(np.loadtxt(path).astype(np.float32)[:, :3].copy()).to(device)
