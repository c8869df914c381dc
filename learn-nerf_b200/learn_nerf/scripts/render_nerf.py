"""Host-side mirror of the rendering loop of learn_nerf/scripts/render_nerf.py (RenderSession.
render_view, :85-97): rays of a view -> chunks of ``batch_size`` -> ``render_rays`` -> uint8 image.

Ray generation, rendering and the uint8 conversion all run on the device; with
torch.distributed initialised the image rows are sharded over the ranks (no collective in the
render itself, one all_gather of the uint8 rows if ``gather`` is set).
"""
from typing import Optional

import numpy as np
import torch

from .. import _native, parallel, prng
from ..dataset import CameraView
from ..render import NeRFRenderer


class _ChunkGraph:
    """One captured ``render_rays`` call for a fixed chunk size: rays are copied into a static
    buffer, the pre-split Threefry keys of the chunk are uploaded (16 bytes), the graph is replayed
    and the colours are read from its static output.  With the reference's default batch_size of
    1024 a view is 8 launches per 0.3 ms of GPU work: launch-bound when issued eagerly."""

    def __init__(self, renderer: NeRFRenderer, n: int, device):
        self.n = n
        self.dev_words = torch.zeros(4, dtype=torch.int32, device=device)
        self.ring = _native.PinnedRing(4, depth=32)
        self.rays = torch.zeros(n, 2, 3, device=device)
        keys = (prng.DeviceKey(self.dev_words[0:2]), prng.DeviceKey(self.dev_words[2:4]))
        torch.cuda.synchronize(device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up: workspace caches, weight packing
                renderer.render_rays(keys, self.rays)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        # The weight re-pack (bf16 path) must be PART of the graph: replays then always read the
        # current fp32 parameters, also after in-place updates (Adam steps, TrainLoop.load) that no
        # eager forward has seen.  Costs one 5 us launch per chunk.
        for tree in (renderer.coarse_params, renderer.fine_params):
            if hasattr(tree, "mark_updated"):
                tree.mark_updated()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = renderer.render_rays(keys, self.rays)["fine"]["outputs"]

    def run(self, key, rays: torch.Tensor) -> torch.Tensor:
        kc, kf = prng.split(key)  # render.py:55
        words = np.array([kc.k0, kc.k1, kf.k0, kf.k1], dtype=np.uint32).view(np.int32)
        self.ring.upload(words, self.dev_words)  # pinned ring: asynchronous and race-free
        self.rays.copy_(rays, non_blocking=True)
        self.graph.replay()
        return self.out


def render_view(renderer: NeRFRenderer, view: CameraView, width: int, height: int, batch_size: int = 1024,
                key=0, device="cuda", gather: bool = False, shard: bool = True, cuda_graph: bool = False
                ) -> torch.Tensor:
    """-> uint8 [rows, width, 3] (this rank's rows; the whole [height, width, 3] if ``gather``).
    ``cuda_graph``: replay one captured ``render_rays`` per full chunk (cached on the renderer).  The
    graph re-packs the bf16 weights from the fp32 parameter buffers on every replay, so in-place
    parameter updates are picked up; only swapping the buffers for OTHER tensors invalidates it."""
    rank, world = parallel.world() if shard else (0, 1)
    row0, row1 = parallel.shard_bounds(height, rank, world)
    rays = view.bare_rays(width, height, device=device, row0=row0, rows=row1 - row0)
    n = rays.shape[0]
    colors = torch.empty(n, 3, device=rays.device)
    chunk_graph = None
    if cuda_graph and n >= batch_size:
        cache = renderer.__dict__.setdefault("_chunk_graphs", {})
        chunk_graph = cache.get((batch_size, str(rays.device)))
        if chunk_graph is None:
            chunk_graph = cache[(batch_size, str(rays.device))] = _ChunkGraph(renderer, batch_size, rays.device)
    for i in range(0, n, batch_size):  # render_nerf.py:88-92
        key, this_key = prng.split(key)
        if chunk_graph is not None and i + batch_size <= n:
            colors[i:i + batch_size] = chunk_graph.run(this_key, rays[i:i + batch_size])
            continue
        colors[i:i + batch_size] = renderer.render_rays(this_key, rays[i:i + batch_size])["fine"]["outputs"]
    image = _native.rgb_to_u8(colors).view(row1 - row0, width, 3)  # :93-96
    if gather and world > 1:
        image = parallel.gather_rows(image, height)
    return image
