"""BASELINE.json configs[3] (c4): two-stage 16x upsampling, 256 -> 1024 -> 4096 points per patch (DisPU/model.py:116-118:
the generator applied twice with shared weights), B patches per GPU, Chamfer/Hausdorff of the final cloud against a
4096-point ground truth, batch-sharded over the ranks with ONE all-reduce of the Chamfer partial sum.
    python tools/bench_c4.py [--b 32] [--steps 10]            (or under torch.distributed.run for N GPUs)
Prints one JSON line (rank 0): input points / s over all ranks, ms per step (max over ranks)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from dispu_b200 import loss_utils  # noqa: E402
from dispu_b200.generator import Generator  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--b", type=int, default=32)
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
rng = np.random.default_rng(4000 + rank)
inp, _ = bench.synth_batch(rank, b=args.b)
gt = np.empty((args.b, 4096, 3), np.float32)
for i in range(args.b):
    c = rng.uniform(-1, 1, (16384, 3)).astype(np.float32)
    gt[i] = c[(c ** 2).sum(1) < 1.0][:4096]
x, g = torch.from_numpy(inp).to(dev), torch.from_numpy(gt).to(dev)
gen = Generator(device=dev)
out = torch.zeros((4,), device=dev)
part = torch.zeros((1,), device=dev)


def step():
    with torch.no_grad():
        _, fine1 = gen(x)                      # 256 -> 1024
        _, fine2 = gen(fine1.contiguous())     # 1024 -> 4096
        loss_utils.chamfer_hausdorff(fine2, g, out=out)
        part[0] = out[2]
    if world > 1:
        dist.all_reduce(part)


for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "c4: two-stage 16x, 256->1024->4096, B=%d per GPU, Chamfer vs 4096-pt gt" % args.b,
                      "n_gpus": world, "ms_per_step": float(ms[0]), "input_points_per_s": args.b * 256 * world / (float(ms[0]) * 1e-3),
                      "cd_sum": float(part[0]), "eager": True}))
if world > 1:
    dist.destroy_process_group()
