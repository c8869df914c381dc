"""Host-side mirror of the rendering loop of learn_nerf/scripts/render_nerf.py (RenderSession.
render_view, :85-97): rays of a view -> chunks of ``batch_size`` -> ``render_rays`` -> uint8 image.

Ray generation, rendering and the uint8 conversion all run on the device; with
torch.distributed initialised the image rows are sharded over the ranks (no collective in the
render itself, one all_gather of the uint8 rows if ``gather`` is set).
"""
from typing import Optional

import torch

from .. import _native, parallel, prng
from ..dataset import CameraView
from ..render import NeRFRenderer


def render_view(renderer: NeRFRenderer, view: CameraView, width: int, height: int, batch_size: int = 1024,
                key=0, device="cuda", gather: bool = False, shard: bool = True) -> torch.Tensor:
    """-> uint8 [rows, width, 3] (this rank's rows; the whole [height, width, 3] if ``gather``)."""
    rank, world = parallel.world() if shard else (0, 1)
    row0, row1 = parallel.shard_bounds(height, rank, world)
    rays = view.bare_rays(width, height, device=device, row0=row0, rows=row1 - row0)
    n = rays.shape[0]
    colors = torch.empty(n, 3, device=rays.device)
    for i in range(0, n, batch_size):  # render_nerf.py:88-92
        key, this_key = prng.split(key)
        colors[i:i + batch_size] = renderer.render_rays(this_key, rays[i:i + batch_size])["fine"]["outputs"]
    image = _native.rgb_to_u8(colors).view(row1 - row0, width, 3)  # :93-96
    if gather and world > 1:
        image = parallel.gather_rows(image, height)
    return image
