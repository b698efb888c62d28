"""Development tool: run one dense layer a few times (for ncu captures).  usage: run_dense_once.py MxKxN [reps]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dispu_b200.dense import PackedLinear  # noqa: E402
M, K, N = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32768x2048x256").split("x"))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x = torch.randn((M, K), device="cuda")
lin = PackedLinear(torch.randn((K, N), device="cuda") / K ** 0.5, torch.zeros((N,), device="cuda"))
y = torch.empty((M, N), device="cuda")
for _ in range(reps):
    lin(x, relu=True, out=y)
torch.cuda.synchronize()
