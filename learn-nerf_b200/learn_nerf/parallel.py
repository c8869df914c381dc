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


class PeerGrads:
    """The flat gradient buffer of every rank, mapped into this process over NVLink
    (torch symmetric memory is only the allocator / rendezvous: plumbing).  ``buffer`` is this
    rank's own gradient tensor, ``ptrs`` the device addresses of all ranks' buffers in rank
    order, ``barrier()`` a cross-rank barrier enqueued on the current stream."""

    def __init__(self, numel: int, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem
        self.buffer = symm_mem.empty(numel, dtype=torch.float32, device=device)
        self.buffer.zero_()
        self._hdl = symm_mem.rendezvous(self.buffer, dist.group.WORLD.group_name)
        self.ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        if len(self.ptrs) != dist.get_world_size() or any(p == 0 for p in self.ptrs):
            raise RuntimeError("symmetric-memory rendezvous returned no peer mappings")

    def barrier(self):
        self._hdl.barrier()


def peer_grads(numel: int, device: torch.device):
    """PeerGrads when the job is multi-GPU over NCCL and peer mappings are available, else None
    (single GPU, the gloo CPU tests, LNRF_ALLREDUCE=nccl, or no NVLink P2P): the caller then
    uses ncclAllReduce followed by the plain Adam kernel."""
    import os
    if world()[1] <= 1 or device.type != "cuda" or dist.get_backend() != "nccl":
        return None
    if os.environ.get("LNRF_ALLREDUCE", "peer").lower() in ("nccl", "none"):
        return None
    try:
        pg = PeerGrads(numel, device)
        ok = torch.ones(1, device=device)
    except Exception as e:  # noqa: BLE001  (every rank must agree on the path: vote below)
        import warnings
        warnings.warn(f"peer gradient mapping unavailable ({e!r}); using ncclAllReduce")
        pg, ok = None, torch.zeros(1, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return pg if float(ok) > 0 else None


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
