"""Development probe: times dispu_dense_tf32x3 at the generator's big shapes under DISPU_DENSE_DBG component masks
(1 = no A loads, 2 = no weight TMA, 4 = no MMA issue) to see which pipeline side bounds the kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dispu_b200.dense import PackedLinear  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty((192 * 1024 * 1024 // 4,), dtype=torch.float32, device=dev)


def t(fn, iters=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


out = []
for (M, K, N) in [(32768, 2048, 256), (524288, 128, 128), (32768, 256, 256), (32768, 1024, 64), (32768, 128, 128), (32768, 128, 256)]:
    x = torch.randn((M, K), device=dev)
    lin = PackedLinear(torch.randn((K, N), device=dev) / K ** 0.5, torch.zeros((N,), device=dev))
    y = torch.empty((M, N), device=dev)
    out.append(f"M{M} K{K} N{N}: {t(lambda: lin(x, relu=True, out=y)):.1f} us")
print(f"dbg={os.environ.get('DISPU_DENSE_DBG', '0')}: " + " | ".join(out))
