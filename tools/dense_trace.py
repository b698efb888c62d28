"""Development probe: clock64 trace of CTA 0's producer / MMA / epilogue roles for one dense launch
(DISPU_DENSE_TRACE=1).  Prints per-iteration phase durations in cycles."""
import ctypes as C
import os
import sys

os.environ["DISPU_DENSE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dispu_b200 import _lib  # noqa: E402
from dispu_b200.dense import PackedLinear  # noqa: E402

dev = torch.device("cuda", 0)
for (M, K, N) in [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(32768, 2048, 256), (524288, 128, 128), (32768, 1024, 64)]:
    x = torch.randn((M, K), device=dev)
    lin = PackedLinear(torch.randn((K, N), device=dev) / K ** 0.5, torch.zeros((N,), device=dev))
    y = torch.empty((M, N), device=dev)
    for _ in range(3):
        lin(x, relu=True, out=y)
    torch.cuda.synchronize()
    buf = np.zeros(1536, np.int64)
    rc = _lib.lib().dispu_dense_debug_trace(buf.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    p, m, e = buf[:512].reshape(64, 8), buf[512:1024].reshape(64, 8), buf[1024:].reshape(64, 8)
    t0 = p[0, 0]
    print(f"== M{M} K{K} N{N}")
    print("producer it: start(rel) wait_empty convert+tmem_st fence+arrive | MMA: wait_full issue+commit | period")
    for i in range(0, 24):
        if p[i, 0] == 0:
            break
        per = p[i + 1, 0] - p[i, 0] if p[i + 1, 0] else 0
        print(f" {i:3d}: {p[i,0]-t0:8d} {p[i,1]-p[i,0]:6d} {p[i,2]-p[i,1]:6d} {p[i,3]-p[i,2]:6d} | {m[i,1]-m[i,0]:6d} {m[i,2]-m[i,1]:6d} | {per:6d}")
    print("epilogue block: start(rel) tmem_ld+wait stage+sync passes+stores")
    for i in range(0, 16):
        if e[i, 0] == 0:
            break
        print(f" {i:3d}: {e[i,0]-t0:8d} {e[i,1]-e[i,0]:6d} {e[i,2]-e[i,1]:6d} {e[i,3]-e[i,2]:6d}")
