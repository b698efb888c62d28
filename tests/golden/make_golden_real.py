"""Real-cloud fixtures: the test INPUTS the reference ships — data/test/Icosahedron.xyz (whose
symmetry gives exact distance ties: the FPS / kNN tie-break stressor), data/test/fandisk.xyz and the reference's own
4x outputs data/test/output/*_X4.xyz — stored as float32 arrays, plus golden results computed by the REAL reference
code (oracle/_ref, built by oracle/build_ref.sh from /root/reference):

  python tests/golden/make_golden_real.py cpu   # here: host loops (nnsearch, threenn_cpu, query_ball_point_cpu,
                                                # nanoflann kNN) and evaluate.py's CD / HD  -> tests/golden/real_clouds.npz
  python tests/golden/make_golden_real.py gpu   # on the GPU box: reference CUDA kernels recompiled for sm_100a (FPS,
                                                # ball query, NnDistance) -> gpurun_out/real_clouds_gpu.npz (copy to tests/golden/)
The gpu mode reads only real_clouds.npz (the reference checkout does not travel to the GPU box)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402

REF_DATA = "/root/reference/data/test"


def normalize(pc):
    """Common/ops.py:1954-1963 in float32 (what evaluate.py:34-35 applies to both clouds)."""
    c = pc.mean(axis=1, keepdims=True, dtype=np.float32)
    pc = (pc - c).astype(np.float32)
    f = np.sqrt((pc * pc).sum(-1, keepdims=True, dtype=np.float32)).max(axis=1, keepdims=True)
    return (pc / f).astype(np.float32)


def make_cpu():
    out = {}
    for name in ("Icosahedron", "fandisk"):
        out[name] = np.loadtxt(os.path.join(REF_DATA, name + ".xyz"))[:, :3].astype(np.float32)[None]
        out[name + "_X4"] = np.loadtxt(os.path.join(REF_DATA, "output", name + "_X4.xyz"))[:, :3].astype(np.float32)[None]
    for name in ("Icosahedron", "fandisk"):
        x, y = out[name], out[name + "_X4"]
        # evaluate.py:34-39,153-165: normalise both, nn_distance(pred, gt), CD = mean + mean, HD = max + max
        # (here pred = the reference's 4x output, "gt" = its 2048-point input: a results check on real surfaces)
        p, g = normalize(y), normalize(x)
        d_f, i_f = R.cpu_nnsearch(p, g)
        d_b, i_b = R.cpu_nnsearch(g, p)
        out[name + "_eval_cd"] = np.float64(d_f.astype(np.float64).mean() + d_b.astype(np.float64).mean())
        out[name + "_eval_hd"] = np.float64(d_f.max() + d_b.max())
        out[name + "_nn_d1"], out[name + "_nn_i1"], out[name + "_nn_d2"], out[name + "_nn_i2"] = R.cpu_nn_distance(x, y)
        out[name + "_3nn_d"], out[name + "_3nn_i"] = R.cpu_three_nn(y, x)
        out[name + "_qbp_r0.2_ns32"] = R.cpu_query_ball_point(0.2, 32, x, x[:, ::8].copy())
        out[name + "_knn16"] = R.cpu_knn_batch(x, x, 16).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "real_clouds.npz"), **out)
    print("wrote real_clouds.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


def make_gpu():
    import torch
    d = np.load(os.path.join(HERE, "real_clouds.npz"))
    out = {}
    dev = "cuda"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    for name in ("Icosahedron", "fandisk"):
        x, y = T(d[name]), T(d[name + "_X4"])
        for tag, cloud, m in ((name, x, 512), (name + "_X4", y, 2048)):
            n = cloud.shape[1]
            temp = torch.empty((32, n), dtype=torch.float32, device=dev)
            idx = torch.empty((1, m), dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            R.gpu_fps(1, n, m, cloud.data_ptr(), temp.data_ptr(), idx.data_ptr())
            torch.cuda.synchronize()
            out[f"{tag}_fps{m}"] = idx.cpu().numpy()
        q = x[:, ::8].contiguous()
        m, ns = q.shape[1], 32
        rad = torch.full((1,), 0.2, dtype=torch.float32, device=dev)
        qi = torch.zeros((1, m, ns), dtype=torch.int32, device=dev)
        qc = torch.zeros((1, m), dtype=torch.int32, device=dev)
        R.gpu_query_ball_point(1, x.shape[1], m, rad.data_ptr(), ns, x.data_ptr(), q.data_ptr(), qi.data_ptr(), qc.data_ptr())
        torch.cuda.synchronize()
        out[name + "_gpu_qbp_idx"], out[name + "_gpu_qbp_cnt"] = qi.cpu().numpy(), qc.cpu().numpy()
        n, m2 = x.shape[1], y.shape[1]
        r1 = torch.empty((1, n), device=dev); i1 = torch.empty((1, n), dtype=torch.int32, device=dev)
        r2 = torch.empty((1, m2), device=dev); i2 = torch.empty((1, m2), dtype=torch.int32, device=dev)
        R.gpu_nn_distance(1, n, x.data_ptr(), m2, y.data_ptr(), r1.data_ptr(), i1.data_ptr(), r2.data_ptr(), i2.data_ptr())
        torch.cuda.synchronize()
        out[name + "_gpu_nn_d1"], out[name + "_gpu_nn_i1"] = r1.cpu().numpy(), i1.cpu().numpy()
        out[name + "_gpu_nn_d2"], out[name + "_gpu_nn_i2"] = r2.cpu().numpy(), i2.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "real_clouds_gpu.npz"), **out)
    print("wrote gpurun_out/real_clouds_gpu.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    make_gpu() if (len(sys.argv) > 1 and sys.argv[1] == "gpu") else make_cpu()
