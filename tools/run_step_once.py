"""Runs the c2 step (fused generator forward + Chamfer/Hausdorff of coarse and fine) eagerly N times — the
target of the ncu captures under profiles/ (launch list and --set full of single kernels).
Usage: python tools/run_step_once.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dispu_b200 import loss_utils  # noqa: E402
from dispu_b200.generator import Generator  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
inp, gt = bench.synth_batch(0)
x, g = torch.from_numpy(inp).to(dev), torch.from_numpy(gt).to(dev)
gen = Generator(device=dev)
gen.use_streams = False
out = torch.zeros((4,), device=dev)
with torch.no_grad():
    for _ in range(steps):
        coarse, fine = gen(x)
        loss_utils.chamfer_hausdorff(coarse, g, out=out)
        loss_utils.chamfer_hausdorff(fine, g, out=out)
torch.cuda.synchronize()
print("ok", out.tolist())
