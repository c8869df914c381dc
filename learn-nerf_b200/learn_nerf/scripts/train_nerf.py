"""Host-side mirror of the model selection and the training loop of
learn_nerf/scripts/train_nerf.py (``create_model`` :141-170, the loop in ``main`` :72-133).
Not a CLI: the functions take the values the reference reads from argparse.

With torch.distributed initialised every rank draws the same global batches from the feeder and
trains on its contiguous shard of each batch (``parallel.shard_rays``); gradients are exchanged
once per step inside ``TrainLoop`` (fused NVLink peer all-reduce + Adam, NCCL fallback).
"""
import os
from typing import Any, Callable, Dict, Iterator, Optional, Tuple

import torch

from .. import parallel, prng
from ..dataset import ModelMetadata, NeRFDataset
from ..instant_ngp import InstantNGPModel, InstantNGPRefNERFModel
from ..model import ModelBase, NeRFModel
from ..ref_nerf import RefNERFModel
from ..train import TrainLoop


def create_model(metadata: ModelMetadata, instant_ngp: bool = False, ref_nerf: bool = False,
                 precision: str = "bf16") -> Tuple[ModelBase, ModelBase, Dict[str, Any]]:
    """train_nerf.py:141-170: (coarse, fine, train_kwargs) for the four model families.
    ``precision`` selects the NeRF MLP path (bf16 tcgen05 or fp32); the others are fp32."""
    if instant_ngp:
        cls = (lambda **kw: InstantNGPRefNERFModel(sh_degree=4, **kw)) if ref_nerf else InstantNGPModel
        mk = lambda levels: cls(table_sizes=[2 ** 18] * levels, grid_sizes=[2 ** (4 + i // 2) for i in range(levels)],
                                bbox_min=list(metadata.bbox_min), bbox_max=list(metadata.bbox_max))
        return mk(6), mk(16), dict(adam_eps=1e-15, adam_b1=0.9, adam_b2=0.99)
    if ref_nerf:
        return RefNERFModel(sh_degree=4), RefNERFModel(sh_degree=4), {}
    return NeRFModel(precision=precision), NeRFModel(precision=precision), {}


def train(data: NeRFDataset, save_path: str, key=0, lr: float = 1e-4, batch_size: int = 4096, coarse_samples: int = 64,
          fine_samples: int = 128, save_interval: int = 1000, max_steps: Optional[int] = None, instant_ngp: bool = False,
          ref_nerf: bool = False, precision: str = "bf16", density_penalty: Optional[float] = None,
          density_penalty_batch_size: int = 128, shuffle_dir: Optional[str] = None, device="cuda",
          log: Callable[[str], None] = print) -> TrainLoop:
    """The loop of train_nerf.py:72-133 (without the optional test set): feeder -> step_fn -> log ->
    periodic atomic checkpoint."""
    init_key, key = prng.split(key)
    coarse, fine, train_kwargs = create_model(data.metadata, instant_ngp, ref_nerf, precision)
    loop = TrainLoop(coarse, fine, init_rng=init_key, lr=lr, coarse_ts=coarse_samples, fine_ts=fine_samples,
                     density_penalty=density_penalty, density_penalty_batch_size=density_penalty_batch_size,
                     device=device, **train_kwargs)
    if os.path.exists(save_path):
        log(f"loading from checkpoint: {save_path}")
        loop.load(save_path)
    step_fn = loop.step_fn(list(data.metadata.bbox_min), list(data.metadata.bbox_max))
    data_key, _test_data_key, key = prng.split(key, 3)
    rank = parallel.world()[0]
    batches: Iterator[torch.Tensor] = data.iterate_batches(shuffle_dir or save_path + ".shuffled", data_key, batch_size,
                                                           device=device, ray_device=device)
    for i, batch in enumerate(batches):
        if max_steps is not None and i >= max_steps:
            break
        step_key, _test_key, key = prng.split(key, 3)
        losses = step_fn(step_key, parallel.shard_rays(batch).contiguous())
        if rank == 0:
            log(f"step {i}: " + " ".join(f"{k}={float(v):.05}" for k, v in losses.items()))
            if i and i % save_interval == 0:
                loop.save(save_path)
    return loop
