"""Launches every op of the path once at the C2 / C5 shapes (for ncu captures)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import unit_ball  # noqa: E402
import dispu_b200.loss_utils as LU  # noqa: E402
import dispu_b200.tf_grouping as G  # noqa: E402
import dispu_b200.tf_interpolate as I  # noqa: E402
import dispu_b200.tf_nndistance as N  # noqa: E402
import dispu_b200.tf_sampling as S  # noqa: E402

dev = "cuda"
rng = np.random.default_rng(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
for rep in range(int(os.environ.get("REPS", "2"))):
    # C2 shapes: b=32, 1024 points
    x = T(unit_ball(rng, 32, 1024)); y = T(unit_ball(rng, 32, 1024))
    G.knn(16, x, x)
    LU.chamfer_hausdorff(x, y)
    N.nn_distance(x, y)
    idx, cnt = G.query_ball_point(0.07, 20, x, x)          # repulsion-loss shape
    G.group_point(torch.randn((32, 1024, 128), device=dev), G.knn(16, x, x)[1])
    S.farthest_point_sample(256, x)
    d, i = I.three_nn(x, x[:, :256].contiguous())
    I.three_interpolate(torch.randn((32, 256, 64), device=dev), i, I.three_nn_weights(d))
    # C5 shapes: b=64
    for n in (1024, 4096, 16384):
        z = T(unit_ball(rng, 64, n))
        f = S.farthest_point_sample(n // 4, z)
        c = S.gather_point(z, f)
        G.query_ball_point((32.0 / n) ** (1 / 3), 32, z, c)
    # inference merge FPS: one big cloud
    S.farthest_point_sample(8192, T(unit_ball(rng, 1, 24576)))
torch.cuda.synchronize()
print("ok")
