"""The path's one collective as this library's NVLink peer-memory kernel (csrc/peer_reduce.cu, dispu_b200.dist.PeerAllReduce).
Bit-exact against the rank-ordered float32 sum numpy computes; the mailbox protocol (epoch parity, CUDA-graph replay) is
exercised on one device with two 'ranks' on two streams, and across real devices when the box has more than one."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch

from dispu_b200 import _lib


def test_peer_allreduce_needs_process_group():
    from dispu_b200.dist import PeerAllReduce
    with pytest.raises(RuntimeError):
        PeerAllReduce()  # no torch.distributed group in this process: refused, never a silent CPU sum


def _ordered_sum(vals):
    acc = np.zeros(vals.shape[1], np.float32)
    for r in range(vals.shape[0]):
        acc = (acc + vals[r]).astype(np.float32)
    return acc


def _mailboxes(n):
    lib = _lib.lib()
    boxes = []
    for _ in range(n):
        p, h = C.c_void_p(), C.create_string_buffer(64)
        _lib.check("dispu_peer_alloc", lib.dispu_peer_alloc(C.byref(p), h))
        boxes.append(p)
    return boxes


@pytest.mark.gpu
@pytest.mark.parametrize("world,n", [(1, 2), (2, 2), (4, 6), (8, 2)])
def test_peer_allreduce_ranks_on_one_device(world, n):
    """`world` ranks as `world` streams of ONE device sharing plain device mailboxes (no IPC): every rank's kernel
    waits for all the others, so the launches must overlap — they are one warp each."""
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    boxes = _mailboxes(world)
    try:
        peers = torch.tensor([b.value for b in boxes], dtype=torch.int64, device=dev)
        epochs = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(world)]
        rng = np.random.default_rng(world * 10 + n)
        streams = [torch.cuda.Stream(dev) for _ in range(world)]
        ins = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(world)]
        outs = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(world)]
        for step in range(5):  # odd and even epochs, both halves of the mailbox reused
            vals = (rng.standard_normal((world, n)) * 10.0 ** int(rng.integers(-3, 4))).astype(np.float32)
            for r in range(world):
                ins[r].copy_(torch.from_numpy(vals[r]))
            torch.cuda.synchronize()
            for r in range(world):
                _lib.check("dispu_peer_allreduce_sum",
                           lib.dispu_peer_allreduce_sum(world, r, peers.data_ptr(), epochs[r].data_ptr(), ins[r].data_ptr(),
                                                        outs[r].data_ptr(), n, streams[r].cuda_stream))
            torch.cuda.synchronize()
            want = _ordered_sum(vals)
            for r in range(world):
                assert np.array_equal(outs[r].cpu().numpy(), want), (step, r)
                assert int(epochs[r].item()) == step + 1
    finally:
        torch.cuda.synchronize()
        for b in boxes:
            lib.dispu_peer_free(b)


@pytest.mark.gpu
def test_peer_allreduce_in_place_under_cuda_graph():
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    (box,) = _mailboxes(1)
    try:
        peers = torch.tensor([box.value], dtype=torch.int64, device=dev)
        epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        x = torch.zeros(2, dtype=torch.float32, device=dev)
        s = torch.cuda.Stream(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                _lib.check("dispu_peer_allreduce_sum",
                           lib.dispu_peer_allreduce_sum(1, 0, peers.data_ptr(), epoch.data_ptr(), x.data_ptr(), x.data_ptr(), 2,
                                                        s.cuda_stream))
            for i in range(4):
                x.copy_(torch.tensor([1.5 + i, -2.25], device=dev))
                g.replay()
                s.synchronize()
                assert x.tolist() == [1.5 + i, -2.25]
        assert int(epoch.item()) == 4  # the epoch advances on the device: a replay needs no host work
    finally:
        torch.cuda.synchronize()
        lib.dispu_peer_free(box)


@pytest.mark.gpu
def test_peer_allreduce_missing_peer_returns_nan_not_hang():
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    boxes = _mailboxes(2)
    try:
        peers = torch.tensor([b.value for b in boxes], dtype=torch.int64, device=dev)
        epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        x = torch.ones(2, dtype=torch.float32, device=dev)
        _lib.check("dispu_peer_allreduce_sum",
                   lib.dispu_peer_allreduce_sum(2, 0, peers.data_ptr(), epoch.data_ptr(), x.data_ptr(), x.data_ptr(), 2, 0))
        torch.cuda.synchronize()  # rank 1 never calls: rank 0 gives up after its bounded spin
        assert torch.isnan(x).all()
    finally:
        for b in boxes:
            lib.dispu_peer_free(b)


def test_peer_allreduce_argument_checks():
    lib = _lib.lib()
    assert lib.dispu_peer_mailbox_bytes() >= 2 * 32 * 32
    one = C.c_void_p(16)
    assert lib.dispu_peer_allreduce_sum(0, 0, one, one, one, one, 2, None) == _lib.EINVAL
    assert lib.dispu_peer_allreduce_sum(2, 2, one, one, one, one, 2, None) == _lib.EINVAL
    assert lib.dispu_peer_allreduce_sum(64, 0, one, one, one, one, 2, None) == _lib.EUNSUPPORTED
    assert lib.dispu_peer_allreduce_sum(2, 0, one, one, one, one, 7, None) == _lib.EUNSUPPORTED
    assert lib.dispu_peer_allreduce_sum(2, 0, None, one, one, one, 2, None) == _lib.ENULL


# ---------------------------------------------------------------- real devices, one process per GPU (CUDA IPC + NVLink)
def _worker(rank, world, port, q):
    import torch.distributed as dist
    from dispu_b200.dist import PeerAllReduce
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    red = PeerAllReduce()
    dev = torch.device("cuda", rank)
    x = torch.zeros(2, dtype=torch.float32, device=dev)
    s = torch.cuda.Stream(dev)
    got = []
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            red(x)
        for step in range(6):
            x.copy_(torch.tensor([0.1 * (rank + 1) * (step + 1), 1e-3 * rank - step], dtype=torch.float32))
            g.replay() if step % 2 else red(x)
            s.synchronize()
            got.append(x.tolist())
    q.put((rank, got))
    dist.barrier()
    red.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_peer_allreduce_across_devices():
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("one visible GPU: the cross-device leg runs under `gpurun --gpus N` / the driver's scaling run")
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for step in range(6):
        vals = np.array([[np.float32(0.1 * (r + 1) * (step + 1)), np.float32(1e-3 * r - step)] for r in range(world)], np.float32)
        want = _ordered_sum(vals)
        for r in range(world):
            assert np.array_equal(np.array(res[r][step], np.float32), want), (step, r)
