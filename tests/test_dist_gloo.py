"""world_size-2 gloo test (CPU) of the N>1 host logic: batch sharding and the single all-reduce of
the Chamfer partial sums reproduce the single-process batch mean / max."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import unit_ball
from dispu_b200.dist import allreduce_chamfer, allreduce_hausdorff, shard_range
from oracle import oracle as O


def test_shard_range_partitions():
    for total in (0, 1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, pred, gt, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(pred.shape[0], rank, world)
    o = O.chamfer(pred[lo:hi], gt[lo:hi])  # stands in for dispu_chamfer on this rank's shard
    partial = torch.tensor([float(o["per_cd"].astype(np.float64).sum()), 2.0 * float(o["per_cd"].astype(np.float64).sum())])
    mean, count = allreduce_chamfer(partial, hi - lo)
    hd = allreduce_hausdorff(torch.tensor([float(o["hd"])]))
    q.put((rank, mean.tolist(), count, hd.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_chamfer_allreduce():
    rng = np.random.default_rng(5)
    b = 5  # uneven split: 3 + 2
    pred, gt = unit_ball(rng, b, 64), unit_ball(rng, b, 96)
    want = O.chamfer(pred, gt)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, pred, gt, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mean, count, hd in res:
        assert count == b
        np.testing.assert_allclose(mean[0], want["cd"], rtol=1e-6)
        np.testing.assert_allclose(mean[1], 2 * want["cd"], rtol=1e-6)
        np.testing.assert_allclose(hd[0], want["hd"], rtol=1e-6)
