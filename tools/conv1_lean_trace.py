"""Development probe (needs a diagnostic build with marks in the lean gather loop)."""
import ctypes as C, os, sys
os.environ["DISPU_DENSE_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from dispu_b200 import _lib, tf_grouping
from dispu_b200.dense import PackedLinear
dev = torch.device("cuda", 0)
b, m, ns, K, N = 32, 1024, 16, 128, 128
xyz = torch.rand((b, m, 3), device=dev) * 2 - 1
idx = tf_grouping.knn(ns, xyz, xyz)[1]
pm = torch.randn((b, m, K), device=dev); qm = torch.randn((b, m, K), device=dev)
lin = PackedLinear(torch.randn((K, N), device=dev) / K ** 0.5, torch.zeros((N,), device=dev))
ww, bw = torch.randn((3, 16), device=dev), torch.zeros((16,), device=dev)
H = torch.empty((b, m, N * 16), device=dev)
for _ in range(3):
    lin(pm, relu=True, gather=(idx, qm, -1.0), weightnet=(xyz, ww, bw, 0.9995), out=H)
torch.cuda.synchronize()
buf = np.zeros(1536, np.int64)
assert _lib.lib().dispu_dense_debug_trace(buf.ctypes.data_as(C.c_void_p)) == 0
p = buf[:512].reshape(64, 8)
print("dbg", os.environ.get("DISPU_DENSE_DBG", "0"), "it: cp.wait | issue copies | wait empty | convert+STTM | st wait | fence+arrive | period")
for i in range(8, 24):
    if p[i, 0] == 0: break
    print(i, p[i,1]-p[i,0], p[i,2]-p[i,1], p[i,3]-p[i,2], p[i,4]-p[i,3], p[i,5]-p[i,4], p[i,6]-p[i,5], (p[i+1,0]-p[i,0]) if p[i+1,0] else 0)
