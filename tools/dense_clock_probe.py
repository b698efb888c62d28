"""Development probe: does the SM clock hold during a tensor-heavy dense launch?  Times the launch with CUDA events (ns) and
reads CTA 0's clock64 span from the trace buffer (cycles): cycles / ns = effective SM clock."""
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
for (M, K, N) in [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(32768, 2048, 256)]:
    x = torch.randn((M, K), device=dev)
    lin = PackedLinear(torch.randn((K, N), device=dev) / K ** 0.5, torch.zeros((N,), device=dev))
    y = torch.empty((M, N), device=dev)
    for _ in range(5):
        lin(x, relu=True, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for rep in range(10):
        e0.record(); lin(x, relu=True, out=y); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    buf = np.zeros(1536, np.int64)
    rc = _lib.lib().dispu_dense_debug_trace(buf.ctypes.data_as(C.c_void_p))
    p, e = buf[:512].reshape(64, 8), buf[1024:].reshape(64, 8)
    nz = p[:, 0][p[:, 0] > 0]
    per = np.diff(nz)
    ez = e[:, 3][e[:, 3] > 0]
    span = (ez.max() if len(ez) else nz.max()) - nz.min()
    print(f"== M{M} K{K} N{N}: event us min {min(ts):.1f} med {sorted(ts)[5]:.1f}; traced producer its {len(nz)}, median period {np.median(per):.0f} cyc, "
          f"first 64 its span {nz.max()-nz.min()} cyc; last epilogue mark - first producer mark {span} cyc")
    gt = p[:, 4][p[:, 0] > 0]
    print(f"   globaltimer span {gt.max()-gt.min()} ns over {nz.max()-nz.min()} cycles -> {(nz.max()-nz.min())/max(1,(gt.max()-gt.min())):.3f} GHz effective")
    print("   periods:", per[:20].tolist())
    for i in range(8, 16):
        print("   it", i, "wait_empty", p[i,1]-p[i,0], "convert", p[i,2]-p[i,1], "fence+arrive", p[i,3]-p[i,2], "load_chunk", p[i,5]-p[i,3], "tail", p[i,6]-p[i,5], "to next", p[i+1,0]-p[i,6])
