"""Development probe: PointShuffle2's conv1 as the generator launches it (gather front-end + weight-net epilogue,
b=32, m=1024, ns=16, K=N=128) under DISPU_DENSE_DBG component masks (1 no A loads, 4 no MMA issue, 8 no convert /
tcgen05.st, 64 no epilogue body)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dispu_b200 import tf_grouping  # noqa: E402
from dispu_b200.dense import PackedLinear  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty((192 * 1024 * 1024 // 4,), dtype=torch.float32, device=dev)
b, m, ns, K, N = 32, 1024, 16, 128, 128
xyz = torch.rand((b, m, 3), device=dev) * 2 - 1
idx = tf_grouping.knn(ns, xyz, xyz)[1]
pm = torch.randn((b, m, K), device=dev)
qm = torch.randn((b, m, K), device=dev)
lin = PackedLinear(torch.randn((K, N), device=dev) / K ** 0.5, torch.zeros((N,), device=dev))
ww, bw = torch.randn((3, 16), device=dev), torch.zeros((16,), device=dev)
H = torch.empty((b, m, N * 16), device=dev)
G = torch.empty((b, m, ns, N), device=dev)


def t(fn, iters=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


a = t(lambda: lin(pm, relu=True, gather=(idx, qm, -1.0), weightnet=(xyz, ww, bw, 0.9995), out=H))
c = t(lambda: lin(pm, relu=True, gather=(idx, qm, -1.0), out=G))
print(f"dbg={os.environ.get('DISPU_DENSE_DBG', '0')}: gather+weightnet {a:.1f} us | gather only {c:.1f} us")
