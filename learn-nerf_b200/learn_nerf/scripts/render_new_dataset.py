"""Host-side mirror of the per-frame rendering of learn_nerf/scripts/render_new_dataset.py (:85-133):
rays of a view -> chunks -> ``render_rays`` -> uint8 colour image AND the 16-bit z-depth image derived
from the fine level's ``coords`` / ``alphas``.  Not a CLI (camera sampling and PNG writing stay with
the caller); everything between the camera and the two images runs on the device.
"""
from typing import Tuple

import torch

from .. import _native, parallel, prng
from ..dataset import CameraView
from ..render import NeRFRenderer


def render_view_with_depth(renderer: NeRFRenderer, view: CameraView, size: int, batch_size: int = 1024, key=0,
                           max_depth: float = 4.0, device="cuda", shard: bool = True, gather: bool = False
                           ) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (uint8 [rows, size, 3] colours, int32 [rows, size] depth holding the reference's uint32 values
    ``(z * 0xFFFF).astype(uint32)``), for this rank's rows (all rows if ``gather`` or one process).
    z = clip(where(alpha > 0.9, ((coords - origin) @ direction) / (alpha + 1e-8), max_depth), 0,
    max_depth) / max_depth (render_new_dataset.py:100-117)."""
    rank, world = parallel.world() if shard else (0, 1)
    row0, row1 = parallel.shard_bounds(size, rank, world)
    rays = view.bare_rays(size, size, device=device, row0=row0, rows=row1 - row0)
    n = rays.shape[0]
    colors = torch.empty(n, 3, device=rays.device)
    depth = torch.empty(n, dtype=torch.int32, device=rays.device)
    for i in range(0, n, batch_size):  # :92-95
        key, this_key = prng.split(key)
        fine = renderer.render_rays(this_key, rays[i:i + batch_size])["fine"]
        colors[i:i + batch_size] = fine["outputs"]
        _, d32 = _native.z_depth(fine["coords"], fine["alphas"], view.camera_origin, view.camera_direction, max_depth)
        depth[i:i + batch_size] = d32
    image = _native.rgb_to_u8(colors).view(row1 - row0, size, 3)  # :119-121
    depth = depth.view(row1 - row0, size)  # :123-125
    if gather and world > 1:
        image = parallel.gather_rows(image, size)
        depth = parallel.gather_rows(depth, size)
    return image, depth
