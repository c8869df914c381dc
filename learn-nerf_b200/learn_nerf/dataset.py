"""Host-side mirror of the ray-generation part of learn_nerf/dataset.py: CameraView.

Only what feeds the render path is mirrored (camera JSON I/O and ``bare_rays``,
dataset.py:15-78); image loading and the on-disk shuffle are host I/O outside the hot path.
``bare_rays`` runs on the device (lnrf_bare_rays) and is bit-exact with the oracle restatement.
"""
import json
import math
from dataclasses import dataclass
from typing import Tuple

import torch

from . import _native

Vec3 = Tuple[float, float, float]


@dataclass
class CameraView:
    """dataset.py:15-78."""

    camera_direction: Vec3
    camera_origin: Vec3
    x_axis: Vec3
    y_axis: Vec3
    x_fov: float
    y_fov: float

    @classmethod
    def from_json(cls, path: str, **kwargs) -> "CameraView":
        with open(path, "rb") as f:
            info = json.load(f)
        return cls(camera_direction=tuple(info["z"]), camera_origin=tuple(info["origin"]),
                   x_axis=tuple(info["x"]), y_axis=tuple(info["y"]), x_fov=float(info["x_fov"]),
                   y_fov=float(info["y_fov"]), **kwargs)

    def to_json(self) -> str:
        return json.dumps(dict(z=self.camera_direction, origin=self.camera_origin, x=self.x_axis,
                               y=self.y_axis, x_fov=self.x_fov, y_fov=self.y_fov))

    def bare_rays(self, width: int, height: int, device="cuda", row0: int = 0, rows: int = None
                  ) -> torch.Tensor:
        """All rays of the view in raster order, [N,2,3] (origin, direction), dataset.py:52-78.
        ``row0`` / ``rows`` select a block of image rows (how a view is sharded over GPUs)."""
        rows = height - row0 if rows is None else rows
        return _native.bare_rays(self.camera_origin, self.x_axis, self.y_axis, self.camera_direction,
                                 math.tan(self.x_fov / 2), math.tan(self.y_fov / 2), width, height, row0,
                                 rows, device)
