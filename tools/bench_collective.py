"""Latency of the path's one collective alone: K back-to-back all-reduces of 2 floats captured in one CUDA graph, this
library's NVLink peer-memory kernel vs ncclAllReduce.  Launch under torch.distributed.run (one rank per GPU)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dispu_b200.dist import PeerAllReduce  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
K = 200
peer = PeerAllReduce(dev)
x = torch.ones(2, dtype=torch.float32, device=dev)
res = {}
for name, fn in (("peer", lambda: peer(x)), ("nccl", lambda: dist.all_reduce(x))):
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(K):
                x.fill_(1.0 / world)  # keeps the values bounded; one tiny launch on both arms
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = float(t[0]) * 1e3 / (5 * K)
    assert abs(float(x[0]) - 1.0) < 1e-5, (name, x.tolist())
if rank == 0:
    print(json.dumps({"world": world, "us_per_allreduce_incl_fill": res,
                      "note": f"{K} x (fill_ + all-reduce of 2 floats) per graph, 5 replays, device time, max over ranks"}))
dist.barrier()
peer.close()
dist.destroy_process_group()
