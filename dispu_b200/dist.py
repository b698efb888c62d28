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
