"""CPU-only tests: the C-ABI library loads and exports every symbol include/lnrf.h declares,
the host-side helpers (PRNG keys, sharding) behave, and the N>1 data-parallel recipe
(local mean loss -> all-reduce(sum) -> 1/world) equals the single-process gradient, checked
with two gloo processes driving the CPU oracle."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lnrf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lnrf_[a-z0-9_]+)\s*\(", text)))


def test_cabi_library_exports_every_declared_symbol():
    from learn_nerf import _native
    lib = _native.load()  # dlopen works without a GPU
    declared = _declared_symbols()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in include/lnrf.h but not exported by liblnrf.so: {missing}"
    out = subprocess.run(["nm", "-D", "--defined-only", _native.lib_path()], capture_output=True,
                         text=True, check=True).stdout
    exported = set(re.findall(r" T (lnrf_[a-z0-9_]+)", out))
    assert set(declared) <= exported
    assert lib.lnrf_version() >= 1
    assert _native.nerf_param_count() == 593_924
    offs = _native.nerf_param_offsets()
    assert offs[0] == 0 and offs[1] == 60 * 256 and all(o % 4 == 0 for o in offs)
    assert _native.nerf_param_floats() % 4 == 0


def test_no_cpu_fallback():
    """The product path refuses CPU tensors instead of silently computing elsewhere."""
    from learn_nerf import _native
    with pytest.raises(_native.LnrfError):
        _native.sample_coarse(torch.zeros(2, 2, 3), [-1] * 3, [1] * 3, torch.zeros(2, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "learn-nerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_prng_keys():
    """Host key bookkeeping equals the numpy threefry oracle (and with it JAX's documented values)."""
    from learn_nerf import prng
    from oracle import prng_np
    a, b = prng.split(0)
    assert (a.k0, a.k1, b.k0, b.k1) == (4146024105, 967050713, 2718843009, 1272950319)
    for seed in (0, 7, 2 ** 40 + 3):
        ref = prng_np.split(prng_np.prng_key(seed), 3)
        got = prng.split(seed, 3)
        assert [(k.k0, k.k1) for k in got] == [tuple(int(x) for x in r) for r in ref]
        f = prng.fold_in(seed, 1234)
        assert (f.k0, f.k1) == tuple(int(x) for x in prng_np.fold_in(prng_np.prng_key(seed), 1234))
    assert prng.split(7) == prng.split(prng.PRNGKey(7))


def test_shard_bounds_cover_and_balance():
    from learn_nerf.parallel import shard_bounds
    for n in (0, 1, 7, 4096, 640_000):
        for ws in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "learn-nerf_b200"))
sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import make_rays, make_uniforms
from learn_nerf import parallel
from oracle import models_torch as M, train_torch as T
torch.set_num_threads(2)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, ws = parallel.world()
n = 24
batch, uc, uf = make_rays(n, seed=3), make_uniforms(n, 16, 4), make_uniforms(n, 32, 5)
nerf = M.NeRFModel()
params = T.init_params(nerf, nerf, 2)
a, b = parallel.shard_bounds(n, rank, ws)
g, ld, _ = T.grads(nerf, nerf, params, [-1] * 3, [1] * 3, batch[a:b], uc[a:b], uf[a:b], 16, 32)
flat = torch.cat([t.reshape(-1) for _, t in M.tree_leaves(g)])
parallel.allreduce_sum_(flat)
flat /= ws                                   # what lnrf_adam_step's grad_scale does
losses = parallel.mean_scalars_(torch.tensor([ld["coarse"], ld["fine"]]))
rows = parallel.gather_rows(torch.from_numpy(batch[a:b]), n)
if rank == 0:
    gf, ldf, _ = T.grads(nerf, nerf, params, [-1] * 3, [1] * 3, batch, uc, uf, 16, 32)
    full = torch.cat([t.reshape(-1) for _, t in M.tree_leaves(gf)])
    rel = float((flat - full).norm() / full.norm())
    assert rel < 1e-4, rel
    assert abs(float(losses[0]) - ldf["coarse"]) < 1e-5 and abs(float(losses[1]) - ldf["fine"]) < 1e-5
    assert torch.equal(rows, torch.from_numpy(batch))
    print("DP_OK", rel)
dist.destroy_process_group()
"""


def test_data_parallel_recipe_two_gloo_processes(tmp_path):
    """world_size 2 on CPU (gloo): sharded local-mean gradients, summed and divided by world,
    equal the full-batch gradient; gathered rows equal the batch; scalar means agree."""
    script = tmp_path / "worker.py"
    port = 29500 + (os.getpid() % 2000)
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "DP_OK" in outs[0]


def test_nerf_dataset_iterate_batches(tmp_path):
    """The reference's own test of the ray feeder (learn_nerf/test_dataset.py:19-86) restated for
    the host mirror: two 10x10 dummy views, batch_size 51 without repeat -> 4 batches, the last one
    of 200 - 3 * 51 rays; every view contributes exactly its pixel count of rays with its origin,
    their mean direction is the camera axis and their mean colour the image mean; plus: the shard
    directory is reusable (same stream on a second pass), and ``repeat`` keeps going past one epoch.
    (bare_rays is a device kernel: the dummy views take their rays from the oracle restatement.)"""
    from dataclasses import dataclass
    import math
    from learn_nerf.dataset import ModelMetadata, NeRFDataset, NeRFView, ShuffledDataset
    from oracle import render_np

    @dataclass
    class DummyView(NeRFView):
        dummy_image: np.ndarray = None

        def image(self):
            return self.dummy_image

        def bare_rays(self, width, height, device="cpu", row0=0, rows=None):
            return torch.from_numpy(render_np.bare_rays(self.camera_direction, self.camera_origin, self.x_axis,
                                                        self.y_axis, self.x_fov, self.y_fov, width, height))

    rs = np.random.RandomState(1337)
    views = [DummyView(camera_direction=(0.0, 1.0, 0.0), camera_origin=(2.0, 2.0, 2.0), x_axis=(-1.0, 0.0, 0.0),
                       y_axis=(0.0, 0.0, 1.0), x_fov=math.radians(60.0), y_fov=math.radians(60.0),
                       dummy_image=rs.randint(0, 256, (10, 10, 3)).astype(np.uint8)),
             DummyView(camera_direction=(1.0, 0.0, 0.0), camera_origin=(-2.0, 2.0, 2.0), x_axis=(-0.0, 0.0, -1.0),
                       y_axis=(0.0, 1.0, 0.0), x_fov=math.radians(60.0), y_fov=math.radians(60.0),
                       dummy_image=rs.randint(0, 256, (10, 10, 3)).astype(np.uint8))]
    dataset = NeRFDataset(metadata=ModelMetadata(bbox_min=(0.0, 0.0, 0.0), bbox_max=(1.0, 1.0, 1.0)), views=views)
    shard_dir = str(tmp_path / "shards")
    batches = list(dataset.iterate_batches(shard_dir, 1234, batch_size=51, repeat=False, ray_device="cpu"))
    assert len(batches) == 4, "unexpected number of batches"
    assert batches[-1].shape[0] == 200 - 51 * 3, "unexpected last batch size"
    assert all(b.shape[1:] == (3, 3) and b.dtype == torch.float32 for b in batches)
    combined = torch.cat(batches, dim=0).numpy()
    for view in views:
        origin = np.asarray(view.camera_origin, np.float32)
        mask = np.abs(combined[:, 0] - origin).sum(-1) < 1e-5
        assert int(mask.sum()) == 100, f"unexpected number of samples with origin {origin}"
        rays = combined[mask]
        mean_dir = rays[:, 1].mean(0)
        mean_dir /= np.linalg.norm(mean_dir)
        assert abs(float((mean_dir * np.asarray(view.camera_direction, np.float32)).sum()) - 1) < 1e-5
        want = (view.dummy_image.astype(np.float32) / 127.5 - 1).mean(axis=(0, 1))
        assert float(np.abs(rays[:, 2].mean(0) - want).mean()) < 1e-5, "invalid colors for rays"
    # the rays are shuffled (not in raster order), every ray appears exactly once
    ref = np.concatenate([v.rays(device="cpu").numpy() for v in views])
    assert not np.array_equal(combined, ref)
    key = lambda a: sorted(map(bytes, a.reshape(len(a), -1)))
    assert key(combined) == key(ref)
    # second pass over the finished shard directory: identical stream; shard files are raw fp32 [N,3,3]
    again = torch.cat(list(dataset.iterate_batches(shard_dir, 1234, batch_size=51, repeat=False, ray_device="cpu")))
    assert np.array_equal(again.numpy(), combined)
    assert os.path.exists(os.path.join(shard_dir, "done"))
    assert sum(os.path.getsize(os.path.join(shard_dir, str(i))) for i in range(32)) == 200 * 36
    with ShuffledDataset(shard_dir, dataset, 1234) as sd:
        it = sd.iterate_batches(64, repeat=True)
        got = [next(it) for _ in range(5)]  # 320 rays > one epoch of 200
    assert all(b.shape == (64, 3, 3) for b in got)


def test_public_header_has_no_process_wide_switches():
    """re-entrant boundary, no mutable globals after lnrf_init: the tuning / debug setters
    of round 1 are gone from the ABI."""
    import re
    hdr = open(os.path.join(ROOT, "include", "lnrf.h")).read()
    assert not re.search(r"lnrf_set_\w+", hdr)
