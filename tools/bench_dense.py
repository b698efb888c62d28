"""Device times of the tcgen05 3xTF32 dense layer (dispu_dense_tf32x3) vs torch FP32 SGEMM (+bias,+ReLU) at the
generator's layer shapes.  Development tool."""
import sys, os, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from dispu_b200.dense import PackedLinear
dev = "cuda"
flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
def timeit(fn, iters=8):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
rows = []
for M, K, N in [(524288, 128, 128), (32768, 2048, 256), (32768, 482, 256), (32768, 256, 128), (32768, 256, 256), (32768, 131, 128),
                (8192, 120, 48), (32768, 128, 256), (32768, 64, 3)]:
    x = torch.randn((M, K), device=dev); w = torch.randn((K, N), device=dev) / K ** 0.5; b = torch.randn((N,), device=dev)
    lin = PackedLinear(w, b)
    ms = timeit(lambda: lin(x, relu=True))
    ms_t = timeit(lambda: torch.relu_(torch.addmm(b, x, w)))
    ref = torch.relu(x.double() @ w.double() + b.double())
    err = float((lin(x, relu=True).double() - ref).abs().max()); err_t = float((torch.relu(torch.addmm(b, x, w)).double() - ref).abs().max())
    fl = 2.0 * M * K * N
    r = {"M": M, "K": K, "N": N, "tc_ms": round(ms, 4), "torch_ms": round(ms_t, 4), "tc_tflops": round(fl / ms / 1e9, 1),
         "torch_tflops": round(fl / ms_t / 1e9, 1), "tc_maxerr": err, "torch_maxerr": err_t}
    rows.append(r); print(json.dumps(r), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "bench_dense.json"), "w"), indent=1)
