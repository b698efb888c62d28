"""Multi-GPU plumbing of the path: one process per GPU, contiguous batch shards, and the ONE
collective the path has — an all-reduce of the per-rank Chamfer partial sums.
Every op of the path is independent per cloud (tf_sampling_g.cu:113, tf_grouping_g.cu:4-8,
tf_nndistance_g.cu:8), so nothing else crosses ranks.  torch.distributed (NCCL on GPUs, gloo in
the CPU tests) is plumbing only."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` clouds for `rank`; the first total % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_chamfer(partial: torch.Tensor, local_clouds: int, group=None):
    """partial: (k,) per-rank SUMS over local clouds of per-cloud losses (e.g. [sum cd_coarse, sum cd_fine],
    out[2] of dispu_chamfer).  Returns the global batch means (loss_utils.py:63 reduce_mean over the
    whole batch) and the global cloud count.  One all-reduce: the count rides in the same buffer."""
    buf = torch.cat([partial.reshape(-1).to(torch.float32),
                     torch.tensor([float(local_clouds)], dtype=torch.float32, device=partial.device)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    total = buf[-1]
    return buf[:-1] / total, int(round(float(total)))


def allreduce_hausdorff(hd_local: torch.Tensor, group=None):
    """hausdorff_loss is a max over the batch (loss_utils.py:83); only logged by the trainer."""
    out = hd_local.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.MAX, group=group)
    return out


class PeerAllReduce:
    """The path's one collective as this library's own NVLink peer-memory kernel (csrc/peer_reduce.cu) instead of an
    ncclAllReduce launch: each rank's `dispu_peer_allreduce_sum` stores its <= 6 floats into every peer's mailbox
    (CUDA-IPC-mapped device memory), waits for the peers' values and sums them in rank order — one ~NVLink-latency kernel
    per rank, CUDA-graph capturable, bit-identical on all ranks.  torch.distributed only carries the 64-byte IPC handles
    once, here.  Single node only (every pair of ranks must have P2P access); raises otherwise — the caller then keeps
    dist.all_reduce."""

    MAX_VALUES = 6

    def __init__(self, device=None, group=None):
        import ctypes as C
        import socket

        from . import _lib

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerAllReduce: torch.distributed is not initialised")
        if not torch.cuda.is_available():
            raise RuntimeError("PeerAllReduce: needs CUDA devices (peer memory over NVLink); there is no CPU path")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._lib = _lib
        self._own = C.c_void_p()
        self._opened = []
        with torch.cuda.device(self.device):
            handle = C.create_string_buffer(64)
            _lib.check("dispu_peer_alloc", _lib.lib().dispu_peer_alloc(C.byref(self._own), handle))
            mine = (socket.gethostname(), self.device.index, bytes(handle.raw))
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            ptrs = []
            err = None
            for r, (host, index, h) in enumerate(everyone):
                if r == self.rank:
                    ptrs.append(self._own.value)
                    continue
                if host != mine[0]:
                    err = f"rank {r} is on host {host!r}, this rank on {mine[0]!r}: peer memory is single-node"
                    break
                if index == self.device.index:
                    err = f"ranks {r} and {self.rank} share cuda:{index}: one process per GPU"
                    break
                if not torch.cuda.can_device_access_peer(self.device.index, index):
                    err = f"cuda:{self.device.index} has no P2P access to cuda:{index}"
                    break
                p = C.c_void_p()
                rc = _lib.lib().dispu_peer_open(h, C.byref(p))
                if rc != 0:
                    err = f"cudaIpcOpenMemHandle of rank {r}'s mailbox failed (rc={rc})"
                    break
                self._opened.append(p)
                ptrs.append(p.value)
            # every rank must take the same decision, or the survivors would wait for a peer that fell back
            flags = [None] * self.world
            dist.all_gather_object(flags, err, group=group)
            bad = [f"rank {r}: {e}" for r, e in enumerate(flags) if e]
            if bad:
                self.close()
                raise RuntimeError("PeerAllReduce unavailable: " + "; ".join(bad))
            self._peers = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
            self._epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
            torch.cuda.synchronize(self.device)
        dist.barrier(group=group)  # every mailbox is mapped and zeroed before anyone's first store
        # self-test before anyone relies on it: IPC mappings that open but do not carry stores (no NVLink / P2P disabled
        # in the container) would otherwise surface as NaNs in the middle of a run.  Collective decision again.
        with torch.cuda.device(self.device):
            probe = torch.tensor([float(self.rank + 1), 1.0], dtype=torch.float32, device=self.device)
            self(probe)
            ok = probe.tolist() == [self.world * (self.world + 1) / 2.0, float(self.world)]
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            self.close()
            raise RuntimeError(f"PeerAllReduce unavailable: self-test failed on ranks {[r for r, f in enumerate(flags) if not f]}")

    def __call__(self, values: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """out[i] = sum over ranks of values[i] (out defaults to values: in place).  Enqueued on the current stream."""
        from ._t import call, stream_ptr

        if not (values.is_cuda and values.dtype == torch.float32 and values.is_contiguous() and values.dim() == 1):
            raise ValueError("PeerAllReduce: expected a contiguous 1-D float32 CUDA tensor")
        if not 1 <= values.numel() <= self.MAX_VALUES:
            raise ValueError(f"PeerAllReduce: 1..{self.MAX_VALUES} values (the path's collective is a handful of floats)")
        out = values if out is None else out
        if out.shape != values.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != values.device:
            raise ValueError("PeerAllReduce: out must match values")
        call("dispu_peer_allreduce_sum", self.world, self.rank, self._peers.data_ptr(), self._epoch.data_ptr(),
             values.data_ptr(), out.data_ptr(), values.numel(), stream_ptr())
        return out

    def close(self):
        lib = self._lib.lib()
        for p in self._opened:
            lib.dispu_peer_close(p)
        self._opened = []
        if self._own:
            lib.dispu_peer_free(self._own)
            self._own = None
