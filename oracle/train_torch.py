"""torch-CPU restatement of learn_nerf/train.py: losses, grads, Adam.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (no JAX here).

``losses`` follows train.py:114-165; the density-penalty branch (:153-184) takes its random
points as explicit inputs (``density_points = (coords[B,3], dirs[B,3])``), like the uniforms.  ``adam_update`` restates optax.adam as used at train.py:59
[recalled: optax is absent]: m <- b1 m + (1-b1) g; v <- b2 v + (1-b2) g^2;
theta <- theta - lr * (m / (1-b1^t)) / (sqrt(v / (1-b2^t)) + eps), t from 1.
Sample positions come from oracle.render_np (bit-exact fp32 stage); the
differentiable part runs in ``dtype`` (fp32 or fp64).
"""
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import render_np
from .models_torch import as_numpy_model_fn, tree_leaves, tree_map
from .render_torch import render_level


def default_loss_weights() -> Dict[str, float]:  # train.py:187-191
    return dict(normal_mse=3e-4, neg_normal=0.1)


def init_params(coarse, fine, seed: int):
    """train.py:47-58 (PRNG differs: torch.Generator instead of jax keys)."""
    gen = torch.Generator().manual_seed(seed)
    return dict(coarse=coarse.init(gen), fine=fine.init(gen),
                background=torch.tensor([-1.0, -1.0, -1.0]))


def losses(coarse, fine, params, bbox_min, bbox_max, batch: np.ndarray, u_coarse: np.ndarray,
           u_fine: np.ndarray, coarse_ts: int, fine_ts: int,
           loss_weights: Optional[Dict[str, float]] = None, dtype=torch.float32,
           fixed_fine_ts: Optional[np.ndarray] = None, density_penalty: Optional[float] = None,
           density_points=None):
    """-> (total_loss tensor, loss_dict, render_out).  ``batch`` [N,3,3] numpy fp32.

    ``fixed_fine_ts`` lets a test inject fine sample positions computed elsewhere
    (e.g. by the CUDA path) so that gradient comparisons are not perturbed by
    1-ulp density differences moving samples.
    """
    loss_weights = loss_weights or default_loss_weights()
    batch = np.asarray(batch, np.float32)
    rays = batch[:, :2]
    t_min, t_max, mask = render_np.ray_t_range(bbox_min, bbox_max, rays)
    cs = render_np.RaySamples.stratified_sampling(t_min, t_max, mask, coarse_ts, u_coarse)

    p = tree_map(lambda t: t.to(dtype), params)
    tb = torch.from_numpy(rays).to(dtype)
    tmin, tmax = torch.from_numpy(t_min).to(dtype), torch.from_numpy(t_max).to(dtype)
    tmask = torch.from_numpy(mask)
    c_out, c_aux = render_level(coarse, p["coarse"], p["background"], tb,
                                torch.from_numpy(cs.ts).to(dtype), tmin, tmax, tmask)
    if fixed_fine_ts is None:
        fs = cs.fine_sampling(fine_ts, u_fine, c_out["densities"].detach().float().numpy())
        f_ts = fs.ts
    else:
        f_ts = np.asarray(fixed_fine_ts, np.float32)
    f_out, f_aux = render_level(fine, p["fine"], p["background"], tb,
                                torch.from_numpy(f_ts).to(dtype), tmin, tmax, tmask)
    targets = torch.from_numpy(batch[:, 2]).to(dtype)
    coarse_loss = torch.mean((c_out["outputs"] - targets) ** 2)  # :141
    fine_loss = torch.mean((f_out["outputs"] - targets) ** 2)  # :142
    loss_dict = dict(coarse=coarse_loss, fine=fine_loss)
    total = coarse_loss + fine_loss
    for name, l in c_aux.items():  # :146-148
        loss_dict[f"coarse_{name}"] = l
        total = total + loss_weights[name] * l
    for name, l in f_aux.items():  # :149-151
        loss_dict[f"fine_{name}"] = l
        total = total + loss_weights[name] * l
    if density_penalty is not None:  # :153-163, fine model first
        coords = torch.from_numpy(np.asarray(density_points[0], np.float32)).to(dtype)
        dirs = torch.from_numpy(np.asarray(density_points[1], np.float32)).to(dtype)
        for prefix, model in (("fine", fine), ("coarse", coarse)):
            densities, _, _ = model.apply(p[prefix], coords, dirs)  # :182
            penalty = torch.mean(densities)  # :183
            loss_dict[f"{prefix}_density"] = penalty
            total = total + density_penalty * penalty
    render_out = dict(coarse=c_out, fine=f_out, coarse_aux=c_aux, fine_aux=f_aux,
                      coarse_ts=cs.ts, fine_ts=f_ts, t_min=t_min, t_max=t_max, mask=mask)
    return total, loss_dict, (p, render_out)


def grads(coarse, fine, params, *args, **kwargs):
    """jax.grad(losses, has_aux=True)(params), train.py:89-90 -> (grad tree, loss_dict, extra)."""
    leaf_params = tree_map(lambda t: t.detach().clone().requires_grad_(True), params)
    total, loss_dict, (p, render_out) = losses(coarse, fine, leaf_params, *args, **kwargs)
    leaves = [t for _, t in tree_leaves(leaf_params)]
    gs = torch.autograd.grad(total, leaves, allow_unused=True)
    it = iter(gs)
    grad_tree = tree_map(lambda t: None, leaf_params)

    def fill(tree, ref):
        for k in sorted(ref.keys()):
            if isinstance(ref[k], dict):
                fill(tree[k], ref[k])
            else:
                g = next(it)
                tree[k] = torch.zeros_like(ref[k]) if g is None else g
    fill(grad_tree, leaf_params)
    return grad_tree, {k: float(v.detach()) for k, v in loss_dict.items()}, render_out


def tree_norm(tree) -> float:  # train.py:92-97
    return float(torch.sqrt(sum((t.double() ** 2).sum() for _, t in tree_leaves(tree))))


class AdamState:
    def __init__(self, params):
        self.m = tree_map(torch.zeros_like, params)
        self.v = tree_map(torch.zeros_like, params)
        self.count = 0


def adam_update(params, grad_tree, state: AdamState, lr: float, b1=0.9, b2=0.999, eps=1e-7):
    """optax.adam + apply_updates (train.py:59,106) -> new params tree."""
    state.count += 1
    t = state.count
    bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
    state.m = _zip_map(lambda m, g: b1 * m + (1 - b1) * g, state.m, grad_tree)
    state.v = _zip_map(lambda v, g: b2 * v + (1 - b2) * g * g, state.v, grad_tree)
    upd = _zip_map(lambda m, v: (m / bc1) / (torch.sqrt(v / bc2) + eps), state.m, state.v)
    return _zip_map(lambda p, u: p - lr * u, params, upd)


def _zip_map(fn, a, b):
    if isinstance(a, dict):
        return {k: _zip_map(fn, a[k], b[k]) for k in a}
    return fn(a, b)


def train_step(coarse, fine, params, state: AdamState, lr, adam_kwargs, *args, **kwargs
               ) -> Tuple[dict, Dict[str, float]]:
    """One TrainLoop.step_fn call, train.py:85-106."""
    g, loss_dict, _ = grads(coarse, fine, params, *args, **kwargs)
    loss_dict["grad_norm"] = tree_norm(g)
    loss_dict["param_norm"] = tree_norm(params)
    new_params = adam_update(params, g, state, lr, **adam_kwargs)
    return new_params, loss_dict
