"""Host-side mirror of learn_nerf/dataset.py: CameraView / NeRFView / NeRFDataset and the
pre-shuffled on-disk ray feeder.

``bare_rays`` runs on the device (lnrf_bare_rays) and is bit-exact with the oracle restatement.
The feeder (``NeRFDataset.iterate_batches`` / ``ShuffledDataset``, dataset.py:130-263) keeps the
reference's two-stage shuffle and its shard file format (raw float32 ``[N,3,3]`` records, a
``done`` marker), so shard directories are interchangeable; the shard assignment and the
permutations come from the Threefry streams in ``prng`` (restated from memory: the ORDER may differ
from what a given JAX version produces, the reference's own test properties hold,
tests/test_cpu_host.py).  With a CUDA ``device`` the batches are staged through pinned buffers
and copied on a side stream one batch ahead of the consumer.
"""
import json
import math
import os
from dataclasses import dataclass
from typing import Iterator, List, Optional, Tuple

import numpy as np
import torch

from . import _native, prng

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


@dataclass
class NeRFView(CameraView):
    """dataset.py:81-101."""

    def image(self) -> np.ndarray:
        """[H, W, 3] uint8 RGB."""
        raise NotImplementedError

    def rays(self, device="cuda") -> torch.Tensor:
        """[N,3,3] rows (origin, direction, colour in [-1,1]), raster order (dataset.py:88-101)."""
        img = np.asarray(self.image())
        bare = self.bare_rays(img.shape[1], img.shape[0], device=device)
        colors = torch.from_numpy(img.reshape(-1, 3).astype(np.float32) / np.float32(127.5) - np.float32(1.0))
        return torch.cat([bare, colors.to(bare.device)[:, None]], dim=1)


@dataclass
class FileNeRFView(NeRFView):
    """dataset.py:104-111."""

    image_path: str = ""

    def image(self) -> np.ndarray:
        from PIL import Image
        rgba = np.array(Image.open(self.image_path).convert("RGBA"))
        # premultiplied alpha (:109-111); np.round == jnp.round (half to even)
        return np.round(rgba[:, :, :3] * (rgba[:, :, 3:] / 255)).astype(np.uint8)


@dataclass
class ModelMetadata:
    """dataset.py:114-127."""

    bbox_min: Vec3
    bbox_max: Vec3

    @classmethod
    def from_json(cls, path: str) -> "ModelMetadata":
        with open(path, "rb") as f:
            metadata = json.load(f)
        return cls(bbox_min=tuple(metadata["min"]), bbox_max=tuple(metadata["max"]))


@dataclass
class NeRFDataset:
    """dataset.py:130-160."""

    metadata: ModelMetadata
    views: List[NeRFView]

    def iterate_batches(self, dir_path: str, key, batch_size: int, repeat: bool = True, num_shards: int = 32,
                        device=None, ray_device="cuda") -> Iterator[torch.Tensor]:
        """Shuffled [N,3,3] ray batches (dataset.py:134-160).  ``device``: where the batches are
        delivered (None = CPU tensors; a CUDA device = pinned staging + one-batch-ahead copies);
        ``ray_device``: where ``view.rays()`` is evaluated while the shards are first written."""
        with ShuffledDataset(dir_path, self, key, num_shards=num_shards, ray_device=ray_device) as sd:
            yield from sd.iterate_batches(batch_size, repeat=repeat, device=device)


class ShuffledDataset:
    """Pre-shuffled rays of a NeRFDataset on disk (dataset.py:163-263): a two-stage shuffle.  Stage
    one scatters every view's rays over ``num_shards`` files with a random shard id per ray; stage
    two, at read time, visits the shards in a random order and permutes the rays inside each.
    Files ``<dir>/0 .. <dir>/<num_shards-1>`` hold raw float32 ``[N,3,3]`` records and ``<dir>/done``
    marks a finished directory, exactly as the reference writes them."""

    RECORD_FLOATS = 9

    def __init__(self, dir_path: str, dataset: NeRFDataset, key, num_shards: int = 32, ray_device="cuda"):
        self.num_shards = num_shards
        self.shard_key, self.shuffle_key = prng.split(key)
        self._paths = [os.path.join(dir_path, str(i)) for i in range(num_shards)]
        marker = os.path.join(dir_path, "done")
        # Multi-process runs (one rank per GPU sharing the directory): rank 0 alone writes the shards,
        # into a scratch directory that is renamed into place once complete, and every rank waits at a
        # barrier before it reads -- no rank can ever see a truncated or half-written shard.
        from . import parallel
        rank, world = parallel.world()
        if rank == 0 and not os.path.exists(marker):
            tmp_dir = dir_path.rstrip("/") + f".tmp.{os.getpid()}"
            os.makedirs(tmp_dir, exist_ok=True)
            final_paths, self._paths = self._paths, [os.path.join(tmp_dir, str(i)) for i in range(num_shards)]
            self._write_shards(dataset, ray_device)
            self._paths = final_paths
            with open(os.path.join(tmp_dir, "done"), "wb") as f:
                f.write(b"done\n")
            if os.path.isdir(dir_path):  # an unfinished directory of an earlier run: replace its files
                for name in os.listdir(tmp_dir):
                    if name != "done":
                        os.replace(os.path.join(tmp_dir, name), os.path.join(dir_path, name))
                os.replace(os.path.join(tmp_dir, "done"), marker)  # the marker appears last
                os.rmdir(tmp_dir)
            else:
                os.makedirs(os.path.dirname(os.path.abspath(dir_path)), exist_ok=True)
                os.rename(tmp_dir, dir_path)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        if not os.path.exists(marker):
            raise FileNotFoundError(f"{marker}: the shuffled shards were not written (is {dir_path} shared "
                                    "between the ranks?)")

    # ---- stage one
    def _write_shards(self, dataset: NeRFDataset, ray_device):
        key = self.shard_key
        files = [open(p, "wb") for p in self._paths]
        try:
            for view in dataset.views:
                rays = view.rays(device=ray_device) if _takes_device(view) else view.rays()
                rays = rays.detach().cpu().numpy() if isinstance(rays, torch.Tensor) else np.asarray(rays)
                rays = np.ascontiguousarray(rays, dtype=np.float32)
                key, view_key = prng.split(key)
                shard_of = prng.randint_host(view_key, rays.shape[0], 0, self.num_shards)
                order = np.argsort(shard_of, kind="stable")  # rays of one shard stay in raster order
                counts = np.bincount(shard_of, minlength=self.num_shards)
                start = 0
                for shard, count in enumerate(counts.tolist()):
                    if count:
                        files[shard].write(rays[order[start:start + count]].tobytes())
                    start += count
        finally:
            for f in files:
                f.close()

    # ---- stage two
    def _read_shard(self, shard: int) -> np.ndarray:
        return np.fromfile(self._paths[shard], dtype=np.float32).reshape(-1, 3, 3)

    def _host_batches(self, batch_size: int, repeat: bool) -> Iterator[np.ndarray]:
        key = self.shuffle_key
        carry: List[np.ndarray] = []  # rays not yet handed out, oldest first
        carried = 0
        while True:
            key, order_key = prng.split(key)
            for shard in prng.permutation_host(order_key, self.num_shards).tolist():
                key, perm_key = prng.split(key)
                rays = self._read_shard(shard)
                carry.append(rays[prng.permutation_host(perm_key, rays.shape[0])])
                carried += rays.shape[0]
                if carried >= batch_size:
                    pool = np.concatenate(carry, axis=0) if len(carry) > 1 else carry[0]
                    full = (carried // batch_size) * batch_size
                    for a in range(0, full, batch_size):
                        yield pool[a:a + batch_size]
                    carry, carried = ([pool[full:]] if full < carried else []), carried - full
            if not repeat:
                break
        if carried:
            yield np.concatenate(carry, axis=0) if len(carry) > 1 else carry[0]

    def iterate_batches(self, batch_size: int, repeat: bool = False, device=None) -> Iterator[torch.Tensor]:
        """Batches of ``batch_size`` rays (the last one may be short when ``repeat`` is False)."""
        dev = torch.device(device) if device is not None else None
        if dev is None or dev.type != "cuda":
            for b in self._host_batches(batch_size, repeat):
                yield torch.from_numpy(np.ascontiguousarray(b))
            return
        # GPU prefetch: two pinned staging buffers, copies on a side stream, one batch in flight
        stream = torch.cuda.Stream(device=dev)
        pinned = [torch.empty(batch_size, 3, 3).pin_memory() for _ in range(2)]
        copied = [None, None]  # per staging buffer: event of the last copy that read it
        pending = None         # (device tensor, event) of the batch in flight

        def deliver(item):
            t, ev = item
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)      # the consumer's stream sees the finished copy
            t.record_stream(cur)
            return t

        for i, b in enumerate(self._host_batches(batch_size, repeat)):
            if copied[i & 1] is not None:
                copied[i & 1].synchronize()  # the copy that last read this pinned buffer has finished
            buf = pinned[i & 1][: b.shape[0]]
            buf.copy_(torch.from_numpy(np.ascontiguousarray(b)))
            with torch.cuda.stream(stream):
                t = buf.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            copied[i & 1] = ev
            if pending is not None:
                yield deliver(pending)
            pending = (t, ev)
        if pending is not None:
            yield deliver(pending)

    def __enter__(self):
        return self

    def __exit__(self, *args):
        return None  # shard files are opened per read


def _takes_device(view) -> bool:
    import inspect
    return "device" in inspect.signature(view.rays).parameters


def load_dataset(directory: str) -> NeRFDataset:
    """dataset.py:266-288: X.png + X.json per view, metadata.json with the bounding box."""
    dataset = NeRFDataset(metadata=ModelMetadata.from_json(os.path.join(directory, "metadata.json")), views=[])
    for img_name in sorted(os.listdir(directory)):
        if img_name.startswith(".") or not img_name.endswith(".png"):
            continue
        img_path = os.path.join(directory, img_name)
        dataset.views.append(FileNeRFView.from_json(img_path[: -len(".png")] + ".json", image_path=img_path))
    return dataset
