"""numpy fp32 restatement of learn_nerf/render.py, one rounding per reference op.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (no JAX here).

Every array is float32 and every arithmetic op is a separate numpy call, so each
intermediate is rounded exactly once, in the order the reference source writes
them (XLA-CPU may fuse differently; that cannot be checked here).  ``cumsum`` is
strict left-to-right (np.cumsum).  ``exp`` is oracle.expf.expf (shared with the
CUDA kernels).  Uniform random numbers are explicit inputs instead of PRNG keys
(reference: jax.random.uniform, render.py:142).
"""
from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import numpy as np

from .expf import expf

F = np.float32


def ray_t_range(bbox_min, bbox_max, rays, min_t_range=1e-3, epsilon=1e-8):
    """render.py:346-389 (vmapped as in render.py:93-111).

    rays [N,2,3] -> (t_min[N], t_max[N], mask[N] bool).
    """
    rays = np.asarray(rays, F)
    bbox = np.stack([np.asarray(bbox_min, F), np.asarray(bbox_max, F)])  # [2,3]
    origin = rays[:, 0]  # [N,3]
    direction = rays[:, 1]
    offsets = (bbox[None] - origin[:, None]).astype(F)  # [N,2,3]      :368
    denom = (direction + F(epsilon)).astype(F)  # :369
    with np.errstate(divide="ignore", invalid="ignore"):
        ts = (offsets / denom[:, None]).astype(F)
    lo = np.min(ts, axis=1)  # [N,3]                                   :372-378
    hi = np.max(ts, axis=1)
    min_t = np.maximum(F(0), np.max(lo, axis=1)).astype(F)  # :381
    max_t = np.min(hi, axis=1).astype(F)  # :382
    max_t_clipped = np.maximum(max_t, (min_t + F(min_t_range)).astype(F))  # :383
    mask = min_t < max_t  # :386
    t_min = np.where(mask, min_t, F(0)).astype(F)  # :387
    t_max = np.where(mask, max_t_clipped, F(min_t_range)).astype(F)
    return t_min, t_max, mask


@dataclass
class RaySamples:
    """render.py:114-290."""

    t_min: np.ndarray
    t_max: np.ndarray
    mask: np.ndarray
    ts: np.ndarray

    @classmethod
    def stratified_sampling(cls, t_min, t_max, mask, count, uniforms):
        """render.py:121-143; ``uniforms`` [N,count] in [0,1) replaces the key."""
        t_min = np.asarray(t_min, F)
        t_max = np.asarray(t_max, F)
        u = np.asarray(uniforms, F)
        bin_size = ((t_max - t_min).astype(F) / F(count)).astype(F)[:, None]  # :138
        steps = np.arange(0, count, dtype=F)[None]
        bin_starts = ((steps * bin_size).astype(F) + t_min[:, None]).astype(F)  # :139-141
        randoms = (u * bin_size).astype(F)  # :142
        return cls(t_min=t_min, t_max=t_max, mask=np.asarray(mask, bool),
                   ts=(bin_starts + randoms).astype(F))  # :143

    def points(self, rays):
        """render.py:145-153 -> [N,T,3]."""
        rays = np.asarray(rays, F)
        return (rays[:, :1] + (rays[:, 1:] * self.ts[:, :, None]).astype(F)).astype(F)

    def starts(self):  # render.py:259-261
        t_mid = ((self.ts[:, 1:] + self.ts[:, :-1]).astype(F) / F(2)).astype(F)
        return np.concatenate([self.t_min[:, None], t_mid], axis=1)

    def ends(self):  # render.py:263-265
        t_mid = ((self.ts[:, 1:] + self.ts[:, :-1]).astype(F) / F(2)).astype(F)
        return np.concatenate([t_mid, self.t_max[:, None]], axis=1)

    def deltas(self):  # render.py:267-268
        return (self.ends() - self.starts()).astype(F)

    def termination_probs(self, densities):
        """render.py:270-287 -> [N,T+1]; last column is the escape probability."""
        densities = np.asarray(densities, F)
        density_dt = (densities * self.deltas()).astype(F)  # :271
        acc_cur = np.cumsum(density_dt, axis=1, dtype=F)  # :275 sequential
        acc_prev = np.concatenate([np.zeros_like(acc_cur[:, :1]), acc_cur], axis=1)
        prob_survive = expf(-acc_prev)  # :279
        prob_terminate = np.concatenate(
            [(F(1) - expf(-density_dt)).astype(F), np.ones_like(acc_cur[:, :1])], axis=1
        )  # :283-285
        return (prob_survive * prob_terminate).astype(F)  # :287

    def render_rays(self, densities, rgbs, background):
        """render.py:155-176 -> [N,3].  Sum over T+1 is left-to-right in fp32."""
        probs = self.termination_probs(densities)
        background = np.asarray(background, F)
        colors = np.concatenate(
            [np.asarray(rgbs, F), np.tile(background[None, None], [len(self.ts), 1, 1])], axis=1
        )
        prod = (probs[..., None] * colors).astype(F)
        acc = np.zeros([len(self.ts), 3], F)
        for t in range(prod.shape[1]):
            acc = (acc + prod[:, t]).astype(F)
        return np.where(self.mask[:, None], acc, background[None]).astype(F)

    def render_alpha(self, densities):
        """render.py:178-190 -> [N,1]."""
        probs = self.termination_probs(densities)
        return np.where(self.mask[:, None], (F(1) - probs[:, -1:]).astype(F), F(0)).astype(F)

    def average_aux_losses(self, densities, aux: Dict[str, np.ndarray]):
        """render.py:192-209 -> dict of scalars."""
        probs = self.termination_probs(densities)[:, :-1]
        out = {}
        for k, v in aux.items():
            per_ray = np.sum((np.asarray(v, F) * probs).astype(F), axis=-1, dtype=F)
            out[k] = F(np.mean(np.where(self.mask, per_ray, F(0)), dtype=np.float64))
        return out

    def fine_sampling(self, count, uniforms, densities, combine=True, eps=1e-8,
                      return_indices=False):
        """render.py:211-257; ``uniforms`` [N,count] replaces the key.

        jnp.interp is restated per its documented formula [recalled, JAX absent]:
        i = clip(searchsorted(xp, x, 'right'), 1, len-1), linear blend with the
        dx == 0 guard, then the x < xp[0] / x > xp[-1] clamps.
        """
        w = (self.termination_probs(densities)[:, :-1] + F(eps)).astype(F)  # :232
        xs = np.cumsum(w, axis=1, dtype=F)  # :235
        xs = np.concatenate([np.zeros_like(xs[:, :1]), xs], axis=1)  # :236
        xs = (xs / xs[:, -1:]).astype(F)  # :237
        ys = np.concatenate([self.t_min[:, None], self.ends()], axis=1)  # :238-241
        inputs = RaySamples.stratified_sampling(
            np.zeros_like(self.t_min), np.ones_like(self.t_max), self.mask, count, uniforms
        ).ts  # :244-250
        new_ts, idx = _interp_rows(inputs, xs, ys)  # :251
        if combine:
            new_ts = np.sort(np.concatenate([self.ts, new_ts], axis=1), axis=1)  # :253-255
        out = RaySamples(t_min=self.t_min, t_max=self.t_max, mask=self.mask, ts=new_ts.astype(F))
        if return_indices:
            return out, idx
        return out


def _interp_rows(x, xp, fp) -> Tuple[np.ndarray, np.ndarray]:
    """vmap(jnp.interp)(x[N,M], xp[N,K], fp[N,K]) in fp32; also returns i[N,M]."""
    n, m = x.shape
    k = xp.shape[1]
    out = np.empty([n, m], F)
    idx = np.empty([n, m], np.int32)
    eps0 = np.spacing(np.finfo(F).eps)
    for r in range(n):
        i = np.clip(np.searchsorted(xp[r], x[r], side="right"), 1, k - 1)
        df = (fp[r, i] - fp[r, i - 1]).astype(F)
        dx = (xp[r, i] - xp[r, i - 1]).astype(F)
        delta = (x[r] - xp[r, i - 1]).astype(F)
        dx0 = np.abs(dx) <= eps0
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = (delta / np.where(dx0, F(1), dx)).astype(F)
        f = np.where(dx0, fp[r, i - 1], (fp[r, i - 1] + (ratio * df).astype(F)).astype(F))
        f = np.where(x[r] < xp[r, 0], fp[r, 0], f)
        f = np.where(x[r] > xp[r, -1], fp[r, -1], f)
        out[r] = f
        idx[r] = i
    return out, idx


ModelFn = Callable[[np.ndarray, np.ndarray], Tuple[np.ndarray, np.ndarray, Dict[str, np.ndarray]]]


def render_rays(model_fn: ModelFn, background, batch, ts: RaySamples):
    """Free function render.py:293-343.  ``model_fn(x[M,3], d[M,3])`` plays
    ``model.apply(dict(params=params), x, d)`` and returns numpy arrays."""
    batch = np.asarray(batch, F)
    all_points = ts.points(batch)  # :318
    direction_batch = np.tile(batch[:, 1:2], [1, all_points.shape[1], 1])  # :319
    densities, rgbs, aux = model_fn(all_points.reshape([-1, 3]), direction_batch.reshape([-1, 3]))
    densities = np.asarray(densities, F).reshape(all_points.shape[:-1])
    rgbs = np.asarray(rgbs, F).reshape(all_points.shape)
    aux = {k: np.asarray(v, F).reshape(densities.shape) for k, v in aux.items()}
    outputs = ts.render_rays(densities, rgbs, background)  # :329
    alphas = ts.render_alpha(densities)  # :330
    coords = ts.render_rays(densities, all_points, np.zeros([3], F))  # :331
    aux_mean = ts.average_aux_losses(densities, aux)  # :332
    return dict(outputs=outputs, rgbs=rgbs, densities=densities, alphas=alphas, coords=coords), aux_mean


@dataclass
class NeRFRenderer:
    """render.py:11-111 with explicit uniforms instead of a PRNG key."""

    coarse: ModelFn
    fine: ModelFn
    background: np.ndarray
    bbox_min: np.ndarray
    bbox_max: np.ndarray
    coarse_ts: int
    fine_ts: int
    min_t_range: float = 1e-3

    def t_range(self, batch, epsilon=1e-8):
        return ray_t_range(self.bbox_min, self.bbox_max, batch, self.min_t_range, epsilon)

    def render_rays(self, u_coarse, u_fine, batch, return_samples: Optional[dict] = None):
        """render.py:39-91.  u_coarse [N,coarse_ts], u_fine [N,fine_ts]."""
        t_min, t_max, mask = self.t_range(batch)
        coarse_ts = RaySamples.stratified_sampling(t_min, t_max, mask, self.coarse_ts, u_coarse)
        coarse_out, coarse_aux = render_rays(self.coarse, self.background, batch, coarse_ts)
        fine_ts = coarse_ts.fine_sampling(self.fine_ts, u_fine, coarse_out["densities"])
        fine_out, fine_aux = render_rays(self.fine, self.background, batch, fine_ts)
        if return_samples is not None:
            return_samples["coarse"] = coarse_ts
            return_samples["fine"] = fine_ts
        return dict(coarse=coarse_out, fine=fine_out, coarse_aux=coarse_aux, fine_aux=fine_aux)


# --------------------------------------------------------------------------- rays / images
def bare_rays(camera_direction, camera_origin, x_axis, y_axis, x_fov, y_fov, width, height):
    """CameraView.bare_rays, dataset.py:52-78 -> [H*W, 2, 3] fp32, raster order.

    jnp.linspace(-1, 1, num) is restated as start + i * ((stop - start) / (num - 1)) in fp32 with
    the last point set to stop [recalled]; math.tan(fov / 2) is a Python double that JAX rounds
    to fp32 when it multiplies the fp32 array."""
    import math
    f = np.float32

    def lin(num):
        if num == 1:
            return np.array([-1.0], f)
        step = f(2.0) / f(num - 1)
        out = (f(-1.0) + np.arange(num, dtype=f) * step).astype(f)
        out[-1] = f(1.0)
        return out

    z = np.asarray(camera_direction, f)
    ys = (f(math.tan(y_fov / 2)) * lin(height))[:, None, None] * np.asarray(y_axis, f)
    xs = (f(math.tan(x_fov / 2)) * lin(width))[None, :, None] * np.asarray(x_axis, f)
    directions = ((xs + ys) + z).reshape(-1, 3).astype(f)
    sq = directions * directions
    norm = np.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2]).astype(f)
    directions = (directions / norm[:, None]).astype(f)
    origins = np.tile(np.asarray(camera_origin, f)[None], (height * width, 1))
    return np.stack([origins, directions], axis=1)


def rgb_to_u8(colors):
    """render_nerf.py:93-96: ((colors + 1) * 127.5).astype(uint8), colours clamped to [-1, 1]."""
    c = np.clip(np.asarray(colors, np.float32), -1.0, 1.0)
    return ((c + np.float32(1.0)) * np.float32(127.5)).astype(np.uint8)


def z_depth(coords, alphas, camera_origin, camera_direction, max_depth):
    """scripts/render_new_dataset.py:100-117 -> (z[N] fp32 in [0,1], (z * 0xFFFF).astype(uint32)).
    ``(coords - origin) @ direction`` is a left-to-right fp32 dot product."""
    f = np.float32
    c = (np.asarray(coords, f) - np.asarray(camera_origin, f)).astype(f)
    d = np.asarray(camera_direction, f)
    prod = (c * d).astype(f)
    dot = ((prod[:, 0] + prod[:, 1]).astype(f) + prod[:, 2]).astype(f)
    a = np.asarray(alphas, f).reshape(-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.where(a > f(0.9), (dot / (a + f(1e-8)).astype(f)).astype(f), f(max_depth)).astype(f)
    z = (np.clip(z, f(0.0), f(max_depth)) / f(max_depth)).astype(f)
    return z, (z * f(0xFFFF)).astype(np.uint32)
