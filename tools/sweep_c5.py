"""BASELINE.json config 5: FPS + ball-query sweep, B=64, N in {1k,4k,16k,64k}: FPS m=N/4, ball query of those
m centres into the N points, ns=32, r=(32/N)^(1/3).  Reports per kernel the device time (CUDA events, L2
flushed), the SURVEY.md §8(d) roofline figures against the measured HBM peak, and — when oracle/_ref/libref_gpu.so
is present — the reference's own CUDA kernels recompiled for sm_100a on the same inputs.
With torchrun (WORLD_SIZE>1) every rank runs B/WORLD_SIZE clouds (batch-sharded, no collective) and rank 0
reports the max-over-ranks time.  Writes gpurun_out/sweep_c5[_N].json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import unit_ball  # noqa: E402
import dispu_b200.tf_grouping as G  # noqa: E402
import dispu_b200.tf_sampling as S  # noqa: E402
from oracle import ref as R  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"]
    flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    B = 64 // world
    with_ref = R.available("gpu") and world == 1 and "--no-ref" not in sys.argv
    sizes = [1024, 4096, 16384, 65536]
    if "--quick" in sys.argv:
        sizes = sizes[:3]

    def timeit(fn, iters, default_stream=False):
        fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts))], device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    rows = []
    for n in sizes:
        m, ns, r = n // 4, 32, (32.0 / n) ** (1.0 / 3.0)
        rng = np.random.default_rng(5000 + int(np.log2(n)) + 100 * rank)
        x = torch.from_numpy(unit_ball(rng, B, n)).to(dev)
        it = 5 if n <= 16384 else 2
        fps_ms = timeit(lambda: S.farthest_point_sample(m, x), it)
        centres = S.gather_point(x, S.farthest_point_sample(m, x))
        bq_ms = timeit(lambda: G.query_ball_point(r, ns, x, centres), it)
        Bg = B * world
        sweep = Bg * (m - 1) * n * 16
        comp = Bg * (12 * n + 4 * m)
        bq_alg = Bg * (12 * n + 12 * m + 4 * m * ns + 4 * m)
        row = {"n": n, "m": m, "B_total": Bg, "gpus": world,
               "fps_ms": fps_ms, "fps_sweep_GBps": sweep / fps_ms / 1e6, "fps_sweep_frac_of_hbm": sweep / fps_ms / 1e6 / hbm,
               "fps_compulsory_GBps": comp / fps_ms / 1e6, "fps_point_updates_per_ns": Bg * (m - 1) * n / fps_ms / 1e6,
               "ball_ms": bq_ms, "ball_alg_GBps": bq_alg / bq_ms / 1e6, "ball_frac_of_hbm": bq_alg / bq_ms / 1e6 / hbm}
        if with_ref and n <= 16384:
            temp = torch.empty((32, n), device=dev); out = torch.empty((B, m), dtype=torch.int32, device=dev)
            row["ref_fps_ms"] = timeit(lambda: R.gpu_fps(B, n, m, x.data_ptr(), temp.data_ptr(), out.data_ptr()), 1, True)
            ri = torch.zeros((B, m, ns), dtype=torch.int32, device=dev); rc = torch.zeros((B, m), dtype=torch.int32, device=dev)
            rad = torch.full((B,), r, device=dev)
            row["ref_ball_ms"] = timeit(lambda: R.gpu_query_ball_point(B, n, m, rad.data_ptr(), ns, x.data_ptr(), centres.data_ptr(),
                                                                      ri.data_ptr(), rc.data_ptr()), 1, True)
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump({"hbm_peak_GBps": hbm, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", f"sweep_c5_{world}gpu.json"), "w"), indent=1)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
