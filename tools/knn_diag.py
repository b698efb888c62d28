"""Development probe: 3-D kNN (k=16) on the generator's actual coarse cloud vs uniform unit-ball data."""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import unit_ball
import bench
from dispu_b200.generator import Generator
import dispu_b200.tf_grouping as G
import dispu_b200.loss_utils as LU

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

inp, gt = bench.synth_batch(0)
gen = Generator(device=dev)
with torch.no_grad():
    coarse, fine = gen(torch.from_numpy(inp).to(dev))
coarse = coarse.contiguous()
ub = torch.from_numpy(unit_ball(np.random.default_rng(0), 32, 1024)).to(dev)
c = coarse.cpu().numpy()
print("coarse extent", c.min(), c.max(), "std", c.std(), "dup-structure: mean |copy0 - copy1|", float(np.abs(c[:, :256] - c[:, 256:512]).mean()))
for name, x in (("coarse", coarse), ("fine", fine.contiguous()), ("unit_ball", ub), ("gt", torch.from_numpy(gt).to(dev))):
    print(json.dumps({"data": name, "no_tma_env": os.environ.get("DISPU_NO_TMA"), "knn16_ms": timeit(lambda: G.knn(16, x, x)),
                      "chamfer_ms": timeit(lambda: LU.chamfer_hausdorff(x, ub))}), flush=True)
