"""Development probe: clock64 samples of the tcgen05 EdgeConv kernel (CTA 0, thread 0), DISPU_EDGE_TRACE=1."""
import ctypes as C, os, sys
os.environ["DISPU_EDGE_TRACE"] = "1"
os.environ["DISPU_FK_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dispu_b200 import _lib, ops as OPS
dev = "cuda"
b, n, c = 32, 256, 48
feat = torch.randn((b, n, c), device=dev)
ws = [(torch.randn((ci, 24), device=dev) / ci ** 0.5, None) for ci in (2 * c, 24 + c, 48 + c)]
for _ in range(3):
    y, idx = OPS.dense_conv(feat, ws)
torch.cuda.synchronize()
buf = np.zeros(64, np.int64)
fn = _lib.lib().dispu__edge_debug_trace
fn.argtypes = [C.c_void_p]; fn.restype = C.c_int
assert fn(buf.ctypes.data_as(C.c_void_p)) == 0
t = buf[buf > 0]
print("marks:", len(t))
print("deltas (clk):", (t[1:] - t[:-1]).tolist())
print("total:", int(t[-1] - t[0]))

fn2 = _lib.lib().dispu__fk_debug_trace
fn2.argtypes = [C.c_void_p]; fn2.restype = C.c_int
assert fn2(buf.ctypes.data_as(C.c_void_p)) == 0
t = buf[buf > 0]
print("feature kNN marks:", len(t), "(start | staged | norms | [q staged+rq | distances | selection]* | knn done | proj weights | proj done)")
print("deltas (clk):", (t[1:] - t[:-1]).tolist())
print("total:", int(t[-1] - t[0]))
