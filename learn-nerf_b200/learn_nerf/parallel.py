"""Ray-sharded data parallelism (one process per GPU, torch.distributed / NCCL).

The reference is single-device; every ray is independent through render_rays
(render.py:39-91) and the only cross-ray reductions are the loss means (train.py:141-142)
and the tree norms (train.py:92-104).  So N GPUs each take a contiguous slice of the ray
batch, compute local-mean losses and their gradients, and exchange ONE flat fp32 gradient
buffer per step with an all-reduce (sum); the 1/world factor is folded into the fused
Adam kernel.  Rendering needs no collective at all.
"""
from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n rays: the first n % world ranks get one extra ray."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_rays(batch: torch.Tensor) -> torch.Tensor:
    """This rank's slice of a [N, ...] ray batch (rows of an image stay contiguous)."""
    rank, ws = world()
    a, b = shard_bounds(batch.shape[0], rank, ws)
    return batch[a:b]


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of the flat gradient buffer (no-op for world 1).  Returns the
    buffer; the caller scales by 1/world (lnrf_adam_step's grad_scale)."""
    if world()[1] > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def mean_scalars_(values: torch.Tensor) -> torch.Tensor:
    """Average small logging scalars (per-rank mean losses) over ranks, in place."""
    ws = world()[1]
    if ws > 1:
        dist.all_reduce(values, op=dist.ReduceOp.SUM)
        values /= ws
    return values


def gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Concatenate per-rank row blocks (e.g. rendered image rows) on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [b - a for a, b in (shard_bounds(n_total, r, ws) for r in range(ws))]
    rows = max(sizes)  # all_gather wants equal shapes: pad the short shards
    padded = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(ws)]
    dist.all_gather(bufs, padded)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)
