"""Generates the committed golden fixtures from the REAL reference code (oracle/_ref, built by
oracle/build_ref.sh from /root/reference).

  python tests/golden/make_golden.py cpu   # here: reference host loops   -> tests/golden/ref_cpu.npz
  python tests/golden/make_golden.py gpu   # on a GPU box: reference CUDA kernels recompiled for
                                           # sm_100a -> gpurun_out/ref_gpu.npz (then copied to tests/golden/)

Inputs are seeded and stored next to the outputs so the fixtures are self-contained.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import lattice, unit_ball  # noqa: E402
from oracle import ref as R  # noqa: E402


def inputs():
    rng = np.random.default_rng(20261017)
    d = {}
    d["ball_a"] = unit_ball(rng, 2, 600)          # dataset cloud
    d["ball_b"] = unit_ball(rng, 2, 150)          # second cloud / queries
    d["lat_a"] = lattice(rng, 2, 700, q=6)        # exact ties, duplicates
    d["lat_b"] = lattice(rng, 2, 90, q=6)
    d["feat_b"] = rng.standard_normal((2, 150, 12)).astype(np.float32)
    d["feat_a"] = rng.standard_normal((2, 600, 5)).astype(np.float32)
    d["big"] = unit_ball(rng, 1, 3500)            # crosses the reference's 3072-point smem window
    return d


def make_cpu():
    d = inputs()
    out = dict(d)
    for tag, a, b in (("ball", d["ball_a"], d["ball_b"]), ("lat", d["lat_a"], d["lat_b"])):
        d1, i1, d2, i2 = R.cpu_nn_distance(a, b)
        out[f"{tag}_nn_d1"], out[f"{tag}_nn_i1"], out[f"{tag}_nn_d2"], out[f"{tag}_nn_i2"] = d1, i1, d2, i2
        td, ti = R.cpu_three_nn(a, b)
        out[f"{tag}_3nn_d"], out[f"{tag}_3nn_i"] = td, ti
        out[f"{tag}_qbp_r0.3_ns16"] = R.cpu_query_ball_point(0.3, 16, a, b)
        out[f"{tag}_knn16"] = R.cpu_knn_batch(a, a, 16).astype(np.int32)
    td, ti = out["ball_3nn_d"], out["ball_3nn_i"]
    w = np.maximum(td, np.float32(1e-10))
    w = (np.float32(1) / w)
    w = (w / w.sum(2, keepdims=True, dtype=np.float32)).astype(np.float32)
    out["ball_3nn_w"] = w
    out["ball_interp"] = R.cpu_three_interpolate(d["feat_b"], ti, w)
    gi = out["ball_qbp_r0.3_ns16"].copy()
    gi[gi < 0] = 0
    out["ball_group_idx"] = gi
    out["ball_group"] = R.cpu_group_point(d["feat_a"], gi)
    np.savez_compressed(os.path.join(HERE, "ref_cpu.npz"), **out)
    print("wrote ref_cpu.npz", {k: v.shape for k, v in out.items()})


def make_gpu():
    import torch
    d = inputs()
    out = {}
    dev = "cuda"

    def T(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def fps(x, m):
        b, n, _ = x.shape
        xx = T(x); temp = torch.empty((32, n), device=dev); o = torch.empty((b, m), dtype=torch.int32, device=dev)
        torch.cuda.synchronize(); R.gpu_fps(b, n, m, xx.data_ptr(), temp.data_ptr(), o.data_ptr()); torch.cuda.synchronize()
        return o.cpu().numpy()

    out["ball_fps150"] = fps(d["ball_a"], 150)
    out["lat_fps200"] = fps(d["lat_a"], 200)
    out["big_fps64"] = fps(d["big"], 64)
    out["ballb_fps200"] = fps(d["ball_b"], 200)  # m > n
    for tag, a, q in (("ball", d["ball_a"], d["ball_b"]), ("lat", d["lat_a"], d["lat_b"])):
        b, n, _ = a.shape
        m = q.shape[1]
        for r, ns in ((0.3, 16), (0.12, 8)):
            x1, x2 = T(a), T(q)
            idx = torch.zeros((b, m, ns), dtype=torch.int32, device=dev)
            cnt = torch.zeros((b, m), dtype=torch.int32, device=dev)
            rad = torch.full((b,), r, device=dev)
            torch.cuda.synchronize()
            R.gpu_query_ball_point(b, n, m, rad.data_ptr(), ns, x1.data_ptr(), x2.data_ptr(), idx.data_ptr(), cnt.data_ptr())
            torch.cuda.synchronize()
            out[f"{tag}_qbp_r{r}_ns{ns}_idx"] = idx.cpu().numpy()
            out[f"{tag}_qbp_r{r}_ns{ns}_cnt"] = cnt.cpu().numpy()
        x1, x2 = T(a), T(q)
        r1 = torch.empty((b, n), device=dev); i1 = torch.empty((b, n), dtype=torch.int32, device=dev)
        r2 = torch.empty((b, m), device=dev); i2 = torch.empty((b, m), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        R.gpu_nn_distance(b, n, x1.data_ptr(), m, x2.data_ptr(), r1.data_ptr(), i1.data_ptr(), r2.data_ptr(), i2.data_ptr())
        torch.cuda.synchronize()
        out[f"{tag}_nn_d1"], out[f"{tag}_nn_i1"] = r1.cpu().numpy(), i1.cpu().numpy()
        out[f"{tag}_nn_d2"], out[f"{tag}_nn_i2"] = r2.cpu().numpy(), i2.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_gpu.npz"), **out)
    print("wrote gpurun_out/ref_gpu.npz", sorted(out))


if __name__ == "__main__":
    {"cpu": make_cpu, "gpu": make_gpu}[sys.argv[1]]()
