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
