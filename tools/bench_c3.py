"""BASELINE.json config 3: B=128, dataset 4096 -> npoint 1024 (r 0.2, ns 32, mlp [32,32,64]) -> npoint 512
(r 0.4, ns 32, mlp [64,64,128]): per-kernel device times of the set-abstraction path, with the grouped MLP
on tcgen05 (3xTF32 / TF32) next to the exact-FP32 kernel and an unfused torch restatement (group -> matmul
chain -> max, the reference's op sequence).  Development tool; writes gpurun_out/bench_c3.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dispu_b200.pointnet_util as PU  # noqa: E402
import dispu_b200.tf_grouping as G  # noqa: E402
import dispu_b200.tf_sampling as S  # noqa: E402

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    B = int(os.environ.get("C3_B", "128"))
    rng = np.random.default_rng(3000)
    pts = rng.uniform(-1, 1, (B, 9000, 3)).astype(np.float32)
    xyz = np.stack([p[(p ** 2).sum(1) < 1][:4096] for p in pts])
    x = torch.from_numpy(xyz).to(dev)
    g = torch.Generator().manual_seed(11)

    def layers(widths):
        return [((torch.randn((widths[i], widths[i + 1]), generator=g) / widths[i] ** 0.5).to(dev), torch.zeros(widths[i + 1], device=dev))
                for i in range(len(widths) - 1)]
    l1, l2 = layers([3, 32, 32, 64]), layers([67, 64, 64, 128])
    res = {}
    res["fps_4096_1024_ms"] = timeit(lambda: S.farthest_point_sample(1024, x))
    fidx = S.farthest_point_sample(1024, x)
    new_xyz = S.gather_point(x, fidx)
    res["ball_4096_1024_ms"] = timeit(lambda: G.query_ball_point(0.2, 32, x, new_xyz))
    idx1, _ = G.query_ball_point(0.2, 32, x, new_xyz)
    rows1 = B * 1024 * 32
    fl1 = 2 * rows1 * (3 * 32 + 32 * 32 + 32 * 64)
    for name, prec in (("tf32x3", PU.MLP_TF32X3), ("tf32", PU.MLP_TF32), ("fp32", PU.MLP_FP32)):
        ms = timeit(lambda: PU.grouped_mlp_max(x, None, new_xyz, idx1, l1, precision=prec), iters=5)
        res[f"sa1_mlp_{name}_ms"] = ms
        res[f"sa1_mlp_{name}_tflops"] = fl1 / ms / 1e9
    f1 = PU.grouped_mlp_max(x, None, new_xyz, idx1, l1, precision=PU.MLP_TF32X3)

    def torch_unfused():
        gx = G.group_point(x, idx1) - new_xyz.unsqueeze(2)
        h = gx
        for w, b in l1:
            h = torch.relu(h @ w + b)
        return h.max(dim=2)[0]
    res["sa1_unfused_torch_ms"] = timeit(torch_unfused, iters=3)
    np.testing.assert_allclose(torch_unfused().cpu().numpy(), f1.cpu().numpy(), rtol=2e-5, atol=2e-5)

    res["fps_1024_512_ms"] = timeit(lambda: S.farthest_point_sample(512, new_xyz))
    new_xyz2 = S.gather_point(new_xyz, S.farthest_point_sample(512, new_xyz))
    res["ball_1024_512_ms"] = timeit(lambda: G.query_ball_point(0.4, 32, new_xyz, new_xyz2))
    idx2, _ = G.query_ball_point(0.4, 32, new_xyz, new_xyz2)
    rows2 = B * 512 * 32
    fl2 = 2 * rows2 * (67 * 64 + 64 * 64 + 64 * 128)
    for name, prec in (("tf32x3", PU.MLP_TF32X3), ("tf32", PU.MLP_TF32), ("fp32", PU.MLP_FP32)):
        ms = timeit(lambda: PU.grouped_mlp_max(new_xyz, f1, new_xyz2, idx2, l2, precision=prec), iters=5)
        res[f"sa2_mlp_{name}_ms"] = ms
        res[f"sa2_mlp_{name}_tflops"] = fl2 / ms / 1e9

    def torch_unfused2():
        gx = G.group_point(new_xyz, idx2) - new_xyz2.unsqueeze(2)
        h = torch.cat([gx, G.group_point(f1, idx2)], dim=-1)
        for w, b in l2:
            h = torch.relu(h @ w + b)
        return h.max(dim=2)[0]
    res["sa2_unfused_torch_ms"] = timeit(torch_unfused2, iters=3)
    res["config"] = f"B={B} n=4096 npoint 1024/512 ns=32"
    res["flops"] = {"sa1": fl1, "sa2": fl2}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_c3.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
