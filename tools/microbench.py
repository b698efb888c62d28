"""Per-op device timings (CUDA events) of the sm_100a kernels next to the reference's own CUDA
kernels recompiled for sm_100a (oracle/_ref/libref_gpu.so, if present).  Development tool; the
judged numbers come from bench.py.  Usage: python tools/microbench.py [--quick]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import unit_ball  # noqa: E402

import dispu_b200.tf_sampling as S  # noqa: E402
import dispu_b200.tf_grouping as G  # noqa: E402
import dispu_b200.tf_interpolate as I  # noqa: E402
import dispu_b200.tf_nndistance as N  # noqa: E402
import dispu_b200.loss_utils as LU  # noqa: E402
from oracle import ref as R  # noqa: E402

dev = "cuda"
HAVE_REF = R.available("gpu")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def timeit(fn, iters=10, warm=3, default_stream=False):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        if default_stream:
            torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def main():
    quick = "--quick" in sys.argv
    rng = np.random.default_rng(0)
    rows = []

    def rec(name, cfg, ms, ref_ms=None, **extra):
        r = {"op": name, "cfg": cfg, "ms": round(ms, 4), "ref_ms": None if ref_ms is None else round(ref_ms, 4)}
        r.update(extra)
        rows.append(r)
        print(json.dumps(r), flush=True)

    # FPS
    fps_cfgs = [(32, 256, 64), (32, 1024, 256), (128, 4096, 1024), (64, 1024, 256), (64, 4096, 1024), (1, 2048, 24),
                (1, 24576, 8192)]
    if not quick:
        fps_cfgs += [(64, 16384, 4096), (64, 65536, 16384), (1, 98304, 32768)]
    for b, n, m in fps_cfgs:
        x = T(unit_ball(rng, b, n))
        ms = timeit(lambda: S.farthest_point_sample(m, x), iters=5 if n > 8192 else 10)
        ref_ms = None
        if HAVE_REF and (n <= 16384 or quick is False and b == 1):
            temp = torch.empty((32, n), dtype=torch.float32, device=dev)
            out = torch.empty((b, m), dtype=torch.int32, device=dev)
            ref_ms = timeit(lambda: R.gpu_fps(b, n, m, x.data_ptr(), temp.data_ptr(), out.data_ptr()), iters=3, warm=1,
                            default_stream=True)
        sweep = b * (m - 1) * n * 16
        rec("fps", f"b{b} n{n} m{m}", ms, ref_ms, sweep_GBps=round(sweep / ms / 1e6, 1))

    # ball query
    for b, n, m, r, ns in [(32, 1024, 1024, 0.07, 20), (128, 4096, 1024, 0.2, 32), (64, 1024, 256, (32 / 1024) ** (1 / 3), 32),
                           (64, 4096, 1024, (32 / 4096) ** (1 / 3), 32)] + ([] if quick else [(64, 16384, 4096, (32 / 16384) ** (1 / 3), 32)]):
        xyz = unit_ball(rng, b, n)
        x = T(xyz)
        q = T(xyz[:, :m])
        ms = timeit(lambda: G.query_ball_point(r, ns, x, q))
        ref_ms = None
        if HAVE_REF:
            ri = torch.zeros((b, m, ns), dtype=torch.int32, device=dev)
            rc = torch.zeros((b, m), dtype=torch.int32, device=dev)
            rad = torch.full((b,), r, dtype=torch.float32, device=dev)
            ref_ms = timeit(lambda: R.gpu_query_ball_point(b, n, m, rad.data_ptr(), ns, x.data_ptr(), q.data_ptr(),
                                                           ri.data_ptr(), rc.data_ptr()), iters=3, warm=1, default_stream=True)
        alg = b * (12 * n + 12 * m + 4 * m * ns + 4 * m)
        rec("ball_query", f"b{b} n{n} m{m} r{r:.3f} ns{ns}", ms, ref_ms, alg_GBps=round(alg / ms / 1e6, 1))

    # group point
    for b, n, c, m, ns in [(128, 4096, 3, 1024, 32), (128, 4096, 64, 1024, 32), (32, 512, 64, 128, 64)]:
        pts = torch.randn((b, n, c), device=dev)
        idx = torch.randint(0, n, (b, m, ns), device=dev, dtype=torch.int32)
        ms = timeit(lambda: G.group_point(pts, idx))
        ref_ms = None
        if HAVE_REF:
            out = torch.empty((b, m, ns, c), device=dev)
            ref_ms = timeit(lambda: R.gpu_group_point(b, n, c, m, ns, pts.data_ptr(), idx.data_ptr(), out.data_ptr()),
                            iters=3, warm=1, default_stream=True)
        alg = b * (4 * m * ns + 4 * n * c + 4 * m * ns * c)
        rec("group_point", f"b{b} n{n} c{c} m{m} ns{ns}", ms, ref_ms, alg_GBps=round(alg / ms / 1e6, 1))

    # knn
    for b, n, m, k in [(32, 1024, 1024, 16), (32, 4096, 4096, 16), (1, 2048, 24, 256 if False else 16)]:
        x = T(unit_ball(rng, b, n))
        q = x[:, :m].contiguous()
        ms = timeit(lambda: G.knn(k, x, q))
        rec("knn", f"b{b} n{n} m{m} k{k}", ms, None, pairs_per_ns=round(b * n * m / ms / 1e6, 2))

    # three_nn / interpolate
    for b, n, m, c in [(32, 512, 128, 64), (128, 4096, 1024, 64)]:
        x1, x2 = T(unit_ball(rng, b, n)), T(unit_ball(rng, b, m))
        pts = torch.randn((b, m, c), device=dev)
        ms = timeit(lambda: I.three_nn(x1, x2))
        rec("three_nn", f"b{b} n{n} m{m}", ms)
        d, i = I.three_nn(x1, x2)
        w = I.three_nn_weights(d)
        ms = timeit(lambda: I.three_interpolate(pts, i, w))
        alg = b * (24 * n + 12 * n * c + 4 * n * c)
        rec("three_interpolate", f"b{b} n{n} m{m} c{c}", ms, None, alg_GBps=round(alg / ms / 1e6, 1))

    # nn_distance / chamfer
    for b, n, m in [(32, 1024, 1024), (32, 4096, 4096), (1, 8192, 8192), (32, 16384, 1024)]:
        x1, x2 = T(unit_ball(rng, b, n)), T(unit_ball(rng, b, m))
        ms = timeit(lambda: N.nn_distance(x1, x2))
        ref_ms = None
        if HAVE_REF:
            r1 = torch.empty((b, n), device=dev); ri1 = torch.empty((b, n), dtype=torch.int32, device=dev)
            r2 = torch.empty((b, m), device=dev); ri2 = torch.empty((b, m), dtype=torch.int32, device=dev)
            ref_ms = timeit(lambda: R.gpu_nn_distance(b, n, x1.data_ptr(), m, x2.data_ptr(), r1.data_ptr(), ri1.data_ptr(),
                                                      r2.data_ptr(), ri2.data_ptr()), iters=3, warm=1, default_stream=True)
        rec("nn_distance", f"b{b} n{n} m{m}", ms, ref_ms, pairs_per_ns=round(2 * b * n * m / ms / 1e6, 2))
        ms = timeit(lambda: LU.chamfer_hausdorff(x1, x2))
        rec("chamfer_fused", f"b{b} n{n} m{m}", ms, None, pairs_per_ns=round(2 * b * n * m / ms / 1e6, 2))

    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/microbench.json", "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
