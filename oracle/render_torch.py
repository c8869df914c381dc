"""torch restatement of learn_nerf/render.py for autograd oracles (any dtype).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (no JAX here).

Same maths as oracle.render_np but with torch ops so gradients w.r.t. model
parameters / densities / rgbs / background come from autograd (the reference
gets them from jax.grad, train.py:90).  Sample positions ``ts`` never carry
gradient (render.py:76 stop_gradient; ts depends only on rays, bbox, uniforms),
so the samplers here run under no_grad and may be fed from oracle.render_np.
"""
from typing import Dict

import torch


def starts_ends(ts: torch.Tensor, t_min: torch.Tensor, t_max: torch.Tensor):
    t_mid = (ts[:, 1:] + ts[:, :-1]) / 2  # render.py:259-265
    return (torch.cat([t_min[:, None], t_mid], dim=1), torch.cat([t_mid, t_max[:, None]], dim=1))


def termination_probs(ts, t_min, t_max, densities):
    """render.py:270-287 -> [N,T+1]."""
    s, e = starts_ends(ts, t_min, t_max)
    density_dt = densities * (e - s)
    acc = torch.cumsum(density_dt, dim=1)
    acc_prev = torch.cat([torch.zeros_like(acc[:, :1]), acc], dim=1)
    surv = torch.exp(-acc_prev)
    term = torch.cat([1 - torch.exp(-density_dt), torch.ones_like(acc[:, :1])], dim=1)
    return surv * term


def composite(ts, t_min, t_max, mask, densities, values, background):
    """RaySamples.render_rays, render.py:155-176."""
    probs = termination_probs(ts, t_min, t_max, densities)
    colors = torch.cat([values, background[None, None].expand(values.shape[0], 1, 3)], dim=1)
    return torch.where(mask[:, None], torch.sum(probs[..., None] * colors, dim=1), background[None])


def render_level(model, params, background, batch, ts, t_min, t_max, mask):
    """Free function render_rays, render.py:293-343."""
    n, t = ts.shape
    pts = batch[:, :1] + batch[:, 1:2] * ts[:, :, None]  # :318 / :153
    dirs = batch[:, 1:2].expand(n, t, 3)  # :319
    de, rgb, aux = model.apply(params, pts.reshape(-1, 3), dirs.reshape(-1, 3))
    de = de.reshape(n, t)
    rgb = rgb.reshape(n, t, 3)
    outputs = composite(ts, t_min, t_max, mask, de, rgb, background)
    probs = termination_probs(ts, t_min, t_max, de)
    alphas = torch.where(mask[:, None], 1 - probs[:, -1:], torch.zeros_like(probs[:, -1:]))
    coords = composite(ts, t_min, t_max, mask, de, pts, torch.zeros(3, dtype=ts.dtype))
    aux_mean: Dict[str, torch.Tensor] = {}
    for k, v in aux.items():  # render.py:192-209
        per_ray = torch.sum(v.reshape(n, t) * probs[:, :-1], dim=-1)
        aux_mean[k] = torch.mean(torch.where(mask, per_ray, torch.zeros_like(per_ray)))
    return dict(outputs=outputs, rgbs=rgb, densities=de, alphas=alphas, coords=coords), aux_mean
