"""Runs the C3 set-abstraction kernels a few times (for ncu captures).  C3_B sets the batch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dispu_b200.pointnet_util as PU  # noqa: E402
import dispu_b200.tf_grouping as G  # noqa: E402
import dispu_b200.tf_sampling as S  # noqa: E402

dev = "cuda"
B = int(os.environ.get("C3_B", "128"))
prec = {"x3": PU.MLP_TF32X3, "tf32": PU.MLP_TF32, "fp32": PU.MLP_FP32}[os.environ.get("C3_PREC", "x3")]
rng = np.random.default_rng(3000)
pts = rng.uniform(-1, 1, (B, 9000, 3)).astype(np.float32)
xyz = np.stack([p[(p ** 2).sum(1) < 1][:4096] for p in pts])
x = torch.from_numpy(xyz).to(dev)
g = torch.Generator().manual_seed(11)


def layers(widths):
    return [((torch.randn((widths[i], widths[i + 1]), generator=g) / widths[i] ** 0.5).to(dev), torch.zeros(widths[i + 1], device=dev))
            for i in range(len(widths) - 1)]


l1, l2 = layers([3, 32, 32, 64]), layers([67, 64, 64, 128])
for _ in range(int(os.environ.get("C3_REPS", "3"))):
    new_xyz = S.gather_point(x, S.farthest_point_sample(1024, x))
    idx1, _ = G.query_ball_point(0.2, 32, x, new_xyz)
    f1 = PU.grouped_mlp_max(x, None, new_xyz, idx1, l1, precision=prec)
    new_xyz2 = S.gather_point(new_xyz, S.farthest_point_sample(512, new_xyz))
    idx2, _ = G.query_ball_point(0.4, 32, new_xyz, new_xyz2)
    f2 = PU.grouped_mlp_max(new_xyz, f1, new_xyz2, idx2, l2, precision=prec)
torch.cuda.synchronize()
print("ok", float(f2.sum()))
