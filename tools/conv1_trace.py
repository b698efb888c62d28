"""Development probe: clock64 trace (DISPU_DENSE_TRACE) of PointShuffle2's conv1 as the generator launches it
(gather front-end + weight-net epilogue), the gather-only form and the plain streamed form of the same shape."""
import ctypes as C
import os
import sys

os.environ["DISPU_DENSE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dispu_b200 import _lib, tf_grouping  # noqa: E402
from dispu_b200.dense import PackedLinear  # noqa: E402

dev = torch.device("cuda", 0)
b, m, ns, K, N = 32, 1024, 16, 128, 128
xyz = torch.rand((b, m, 3), device=dev) * 2 - 1
idx = tf_grouping.knn(ns, xyz, xyz)[1]
pm = torch.randn((b, m, K), device=dev)
qm = torch.randn((b, m, K), device=dev)
lin = PackedLinear(torch.randn((K, N), device=dev) / K ** 0.5, torch.zeros((N,), device=dev))
ww, bw = torch.randn((3, 16), device=dev), torch.zeros((16,), device=dev)
H = torch.empty((b, m, N * 16), device=dev)
G = torch.empty((b, m, ns, N), device=dev)
X = torch.randn((b * m * ns, K), device=dev)
Y = torch.empty((b * m * ns, N), device=dev)


def show(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    buf = np.zeros(1536, np.int64)
    rc = _lib.lib().dispu_dense_debug_trace(buf.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    p, mm, e = buf[:512].reshape(64, 8), buf[512:1024].reshape(64, 8), buf[1024:].reshape(64, 8)
    t0 = p[0, 0]
    print(f"== {name}")
    print("producer it: start(rel) wait_empty convert+tmem_st fence+arrive | MMA: wait_full issue+commit | period")
    for i in range(0, 28):
        if p[i, 0] == 0:
            break
        per = p[i + 1, 0] - p[i, 0] if p[i + 1, 0] else 0
        print(f" {i:3d}: {p[i,0]-t0:8d} {p[i,1]-p[i,0]:6d} {p[i,2]-p[i,1]:6d} {p[i,3]-p[i,2]:6d} | {mm[i,1]-mm[i,0]:6d} {mm[i,2]-mm[i,1]:6d} | {per:6d}")
    print("epilogue block: start(rel) tmem_ld+wait stage+sync passes+stores")
    for i in range(0, 24):
        if e[i, 0] == 0:
            break
        print(f" {i:3d}: {e[i,0]-t0:8d} {e[i,1]-e[i,0]:6d} {e[i,2]-e[i,1]:6d} {e[i,3]-e[i,2]:6d}")


show("conv1 gather + weight net", lambda: lin(pm, relu=True, gather=(idx, qm, -1.0), weightnet=(xyz, ww, bw, 0.9995), out=H))
show("conv1 gather only", lambda: lin(pm, relu=True, gather=(idx, qm, -1.0), out=G))
show("plain M524288 K128 N128", lambda: lin(X, relu=True, out=Y))
