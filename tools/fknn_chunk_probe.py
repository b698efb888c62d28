"""Development probe: the EdgeConv block (feature kNN + projection + per-edge MLP) at the c4 second-stage shape
(b=32, n=1024) for the chunk sizes of the feature-kNN kernel (DISPU_FK_P=8|16|32 -> 256 / 512 / 1024 candidates per chunk)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from dispu_b200 import ops as OPS
dev = "cuda"
flush = torch.empty((192 * 1024 * 1024 // 4,), dtype=torch.float32, device=dev)
for n, c in ((1024, 48), (1024, 24), (256, 48)):
    feat = torch.randn((32, n, c), device=dev)
    ws = [(torch.randn((ci, 24), device=dev) / ci ** 0.5, None) for ci in (2 * c, 24 + c, 48 + c)]
    for _ in range(3): y, idx = OPS.dense_conv(feat, ws)
    tot = 0.0
    for _ in range(8):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); OPS.dense_conv(feat, ws); e1.record(); e1.synchronize(); tot += e0.elapsed_time(e1)
    print(f"P={os.environ.get('DISPU_FK_P', 'auto')} dense_conv b=32 n={n} c={c}: {tot / 8 * 1e3:.1f} us  checksum {float(y.sum()):.4f} {int(idx.sum())}")
