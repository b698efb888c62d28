"""Development probe: dispu_knn at the c4 second-stage shape (b=32, n=m=4096, k=16) and at c2's (n=m=1024)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from dispu_b200 import tf_grouping
dev = torch.device("cuda", 0)
flush = torch.empty((192 * 1024 * 1024 // 4,), dtype=torch.float32, device=dev)
for n in (1024, 4096):
    xyz = torch.rand((32, n, 3), device=dev) * 2 - 1
    for _ in range(3): tf_grouping.knn(16, xyz, xyz)
    tot = 0.0
    for _ in range(10):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tf_grouping.knn(16, xyz, xyz); e1.record(); e1.synchronize(); tot += e0.elapsed_time(e1)
    print(f"knn b=32 n=m={n} k=16: {tot / 10 * 1e3:.1f} us")
