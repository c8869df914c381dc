"""Pins the oracle to the REAL reference: runs unixpickle/learn-nerf (JAX/Flax/optax) and writes
tests/golden/jax_reference.npz.  Needs `jax`, `flax`, `optax` and a checkout of the reference:

    LNRF_REFERENCE_PATH=/root/reference JAX_PLATFORMS=cpu python tests/golden/make_golden_jax.py

This image has no JAX (and no network), so the fixture could not be generated here: the parity of
the oracle is UNPINNED until someone runs this one command on a box with JAX and commits the .npz.
`tests/test_jax_parity.py` then checks the oracle restatement (oracle/) against it without needing
JAX, and also runs this generator live whenever `import jax, flax, optax` succeeds.

Everything the fixture holds is produced by the reference's own public API on seeded inputs:
  prng/*     jax.random.PRNGKey / split / fold_in / uniform                 (render.py:55,142)
  render/*   NeRFRenderer.t_range, RaySamples.stratified_sampling / fine_sampling / termination_probs,
             NeRFRenderer.render_rays with random-init NeRFModel            (render.py:39-343)
  models/*   NeRFModel / InstantNGPModel / RefNERFModel / InstantNGPRefNERFModel .apply, with the
             flax-initialised parameter trees saved leaf by leaf            (model.py, instant_ngp.py, ref_nerf.py)
  train/*    TrainLoop.losses, jax.grad of it, and one step_fn call         (train.py:78-165)
"""
import importlib
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from helpers import BBOX_MAX, BBOX_MIN, make_rays  # noqa: E402


def jax_available() -> bool:
    try:
        import flax  # noqa: F401
        import jax  # noqa: F401
        import optax  # noqa: F401
        return True
    except Exception:  # noqa: BLE001
        return False


def reference_path():
    for cand in (os.environ.get("LNRF_REFERENCE_PATH"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "learn_nerf")):
            return cand
    return None


def load_reference():
    """Import the reference package under the alias ``ref_learn_nerf`` (this repo's host mirror is
    also called learn_nerf): its modules only use relative imports."""
    base = reference_path()
    if base is None:
        raise RuntimeError("reference checkout not found (set LNRF_REFERENCE_PATH)")
    pkg_dir = os.path.join(base, "learn_nerf")
    if "ref_learn_nerf" not in sys.modules:
        spec = importlib.util.spec_from_file_location("ref_learn_nerf", os.path.join(pkg_dir, "__init__.py"),
                                                      submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_learn_nerf"] = mod
        spec.loader.exec_module(mod)
    return {name: importlib.import_module(f"ref_learn_nerf.{name}")
            for name in ("render", "model", "instant_ngp", "ref_nerf", "train")}


def _flatten(prefix, tree, out):
    for k, v in tree.items():
        if hasattr(v, "items"):
            _flatten(f"{prefix}/{k}", v, out)
        else:
            out[f"{prefix}/{k}"] = np.asarray(v)


def _key_words(key):
    import jax
    try:
        return np.asarray(jax.random.key_data(key), np.uint32)
    except Exception:  # noqa: BLE001  (old JAX: raw uint32[2] keys)
        return np.asarray(key, np.uint32)


GRIDS = [2 ** (4 + i // 2) for i in range(16)]


def generate(path=None):
    import jax
    import jax.numpy as jnp
    ref = load_reference()
    out = {}
    # ---------------------------------------------------------------- prng
    out["prng/key0"] = _key_words(jax.random.PRNGKey(0))
    out["prng/key_big"] = _key_words(jax.random.PRNGKey(2 ** 31 + 12345))
    k7 = jax.random.PRNGKey(7)
    out["prng/split7"] = np.stack([_key_words(k) for k in jax.random.split(k7)])
    out["prng/split7_3"] = np.stack([_key_words(k) for k in jax.random.split(k7, 3)])
    out["prng/fold7_5"] = _key_words(jax.random.fold_in(k7, 5))
    out["prng/uniform3_5x7"] = np.asarray(jax.random.uniform(jax.random.PRNGKey(3), (5, 7)))
    out["prng/uniform3_33"] = np.asarray(jax.random.uniform(jax.random.PRNGKey(3), (33,)))  # odd count
    # ---------------------------------------------------------------- renderer pieces
    R = ref["render"]
    n = 32
    batch = make_rays(n, seed=101, miss_frac=0.25)
    rays = jnp.asarray(batch[:, :2])
    bmin, bmax = jnp.asarray(BBOX_MIN), jnp.asarray(BBOX_MAX)
    nerf = ref["model"].NeRFModel()
    ex = jnp.zeros((1, 3))
    kc, kf = jax.random.split(jax.random.PRNGKey(21))
    pc = nerf.init(dict(params=kc), ex, ex)["params"]
    pf = nerf.init(dict(params=kf), ex, ex)["params"]
    _flatten("render/params/coarse", pc, out)
    _flatten("render/params/fine", pf, out)
    bg = jnp.asarray([-1.0, 0.25, 0.5])
    rend = R.NeRFRenderer(coarse=nerf, fine=nerf, coarse_params=pc, fine_params=pf, background=bg, bbox_min=bmin,
                          bbox_max=bmax, coarse_ts=64, fine_ts=128)
    key = jax.random.PRNGKey(11)
    t_min, t_max, mask = rend.t_range(rays)
    ck, fk = jax.random.split(key)
    cs = R.RaySamples.stratified_sampling(t_min=t_min, t_max=t_max, mask=mask, count=64, key=ck)
    res = rend.render_rays(key, rays)
    fs = cs.fine_sampling(count=128, key=fk, densities=res["coarse"]["densities"])
    out.update({"render/batch": batch, "render/background": np.asarray(bg), "render/t_min": np.asarray(t_min),
                "render/t_max": np.asarray(t_max), "render/mask": np.asarray(mask),
                "render/coarse_ts": np.asarray(cs.ts), "render/fine_ts": np.asarray(fs.ts),
                "render/coarse_probs": np.asarray(cs.termination_probs(res["coarse"]["densities"]))})
    for lv in ("coarse", "fine"):
        for k in ("outputs", "alphas", "coords", "densities", "rgbs"):
            out[f"render/{lv}/{k}"] = np.asarray(res[lv][k])
    # hard fine-sampling inputs: zero density, one spike, tiny density (flat CDF bins)
    dens = np.abs(np.random.RandomState(5).gamma(0.5, 4.0, (n, 64))).astype(np.float32)
    dens[0] = 0.0
    dens[1] = 0.0
    dens[1, 10] = 1e6
    dens[2, :5] = 1e4
    dens[3] = 1e-12
    out["render/hard_densities"] = dens
    out["render/hard_fine_ts"] = np.asarray(cs.fine_sampling(count=128, key=fk, densities=jnp.asarray(dens)).ts)
    # ---------------------------------------------------------------- models
    rs = np.random.RandomState(31)
    x = rs.uniform(-1.1, 1.1, (64, 3)).astype(np.float32)
    d = rs.randn(64, 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    out["models/x"], out["models/d"] = x, d
    zoo = {
        "nerf": nerf,
        "ngp": ref["instant_ngp"].InstantNGPModel(table_sizes=[2 ** 14] * 16, grid_sizes=GRIDS, bbox_min=bmin,
                                                  bbox_max=bmax),
        "ngp_smooth": ref["instant_ngp"].InstantNGPModel(table_sizes=[2 ** 14] * 6, grid_sizes=GRIDS[:6],
                                                         bbox_min=bmin, bbox_max=bmax, table_smooth=True),
        "refnerf": ref["ref_nerf"].RefNERFModel(sh_degree=4),
        "ngpref": ref["instant_ngp"].InstantNGPRefNERFModel(table_sizes=[2 ** 14] * 16, grid_sizes=GRIDS,
                                                            bbox_min=bmin, bbox_max=bmax, sh_degree=4),
    }
    for i, (name, model) in enumerate(zoo.items()):
        p = model.init(dict(params=jax.random.PRNGKey(40 + i)), ex, ex)["params"]
        if "MultiresHashTableEncoding_0" in p:  # O(1) tables so that the encoding matters
            p = jax.tree_util.tree_map(lambda a: a, p)
            p = dict(p)
            p["MultiresHashTableEncoding_0"] = jax.tree_util.tree_map(lambda a: a * 1e4,
                                                                      p["MultiresHashTableEncoding_0"])
        _flatten(f"models/{name}/params", p, out)
        dens_o, rgb_o, aux = model.apply(dict(params=p), jnp.asarray(x), jnp.asarray(d))
        out[f"models/{name}/density"], out[f"models/{name}/rgb"] = np.asarray(dens_o), np.asarray(rgb_o)
        for k, v in aux.items():
            out[f"models/{name}/aux/{k}"] = np.asarray(v)
    # ---------------------------------------------------------------- train step
    Tm = ref["train"]
    loop = Tm.TrainLoop(nerf, nerf, jax.random.PRNGKey(51), lr=1e-3, coarse_ts=64, fine_ts=128)
    tb = jnp.asarray(make_rays(48, seed=103))
    skey = jax.random.PRNGKey(52)
    _flatten("train/params0", loop.state.params, out)
    (total, ld), grad = jax.value_and_grad(lambda p: loop.losses(skey, bmin, bmax, tb, p), has_aux=True)(
        loop.state.params)
    out["train/batch"] = np.asarray(tb)
    out["train/total"] = np.asarray(total)
    for k, v in ld.items():
        out[f"train/loss/{k}"] = np.asarray(v)
    _flatten("train/grad", grad, out)
    logs = loop.step_fn(bmin, bmax)(skey, tb)
    for k, v in logs.items():
        out[f"train/logs/{k}"] = np.asarray(v)
    _flatten("train/params1", loop.state.params, out)
    out["meta/jax_version"] = np.asarray(jax.__version__)
    path = path or os.path.join(HERE, "jax_reference.npz")
    np.savez_compressed(path, **out)
    return path


if __name__ == "__main__":
    if not jax_available():
        sys.exit("jax / flax / optax are not importable: cannot generate the reference fixture here")
    print("wrote", generate())
