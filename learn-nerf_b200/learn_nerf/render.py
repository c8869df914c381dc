"""Host-side mirror of learn_nerf/render.py: NeRFRenderer, RaySamples, render_rays,
ray_t_range.  Same names, argument order, dict keys and shapes as the reference
(render.py:11-389); arrays are torch CUDA tensors and every stage is a liblnrf
kernel (K1 sampling, K2 MLP, K3 compositing, K4 fine sampling).

``key`` arguments accept a :class:`learn_nerf.prng.PRNGKey`/int seed, or explicit
uniforms (a tensor of the right shape) -- the parity entry point.
"""
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple, Union

import torch

from . import _native, prng
from .model import ModelBase

KeyOrUniforms = Union[prng.PRNGKey, int, torch.Tensor]


def _uniforms(key: KeyOrUniforms, shape, device) -> torch.Tensor:
    if isinstance(key, torch.Tensor):
        if tuple(key.shape) != tuple(shape):
            raise ValueError(f"explicit uniforms have shape {tuple(key.shape)}, expected {tuple(shape)}")
        return key.to(device=device, dtype=torch.float32).contiguous()
    return prng.uniform(key, shape, device)


@dataclass
class NeRFRenderer:
    """render.py:11-111."""

    coarse: ModelBase
    fine: ModelBase
    coarse_params: Any
    fine_params: Any
    background: torch.Tensor
    bbox_min: torch.Tensor
    bbox_max: torch.Tensor
    coarse_ts: int
    fine_ts: int

    min_t_range: float = 1e-3

    def render_rays(self, key, batch: torch.Tensor, _save: bool = False) -> Dict[str, Dict[str, torch.Tensor]]:
        """render.py:39-91.  ``key``: PRNG key / seed, or a pair (u_coarse[N,Tc], u_fine[N,Tf])."""
        batch = _native._f32c(batch.contiguous(), "batch")
        if isinstance(key, (tuple, list)) and isinstance(key[0], (torch.Tensor, prng.DeviceKey)):
            coarse_key, fine_key = key  # explicit uniforms, or pre-split device-resident keys
        else:
            coarse_key, fine_key = prng.split(key)  # :55
        # t_range (:53) and stratified_sampling (:57-63) are one fused launch (K1)
        u_c = _uniforms(coarse_key, (batch.shape[0], self.coarse_ts), batch.device)
        t_min, t_max, mask, ts_c = _native.sample_coarse(batch, _vec3(self.bbox_min),
                                                         _vec3(self.bbox_max), u_c, self.min_t_range)
        coarse_ts = RaySamples(t_min=t_min, t_max=t_max, mask=mask, ts=ts_c)
        coarse_out, coarse_aux = render_rays(model=self.coarse, params=self.coarse_params,
                                             background=self.background, batch=batch, ts=coarse_ts,
                                             _save="coarse" if _save else None)
        fine_ts = coarse_ts.fine_sampling(count=self.fine_ts, key=fine_key,
                                          densities=coarse_out["densities"].detach())
        fine_out, fine_aux = render_rays(model=self.fine, params=self.fine_params,
                                         background=self.background, batch=batch, ts=fine_ts,
                                         _save="fine" if _save else None)
        return dict(coarse=coarse_out, fine=fine_out, coarse_aux=coarse_aux, fine_aux=fine_aux)

    def t_range(self, batch: torch.Tensor, epsilon: float = 1e-8):
        """render.py:93-111 -> (t_min[N], t_max[N], mask[N] bool)."""
        n = batch.shape[0]
        u = torch.zeros(n, 1, device=batch.device)
        t_min, t_max, mask, _ = _native.sample_coarse(batch, _vec3(self.bbox_min), _vec3(self.bbox_max),
                                                      u, self.min_t_range, epsilon)
        return t_min, t_max, mask.bool()


def _vec3(v):
    if isinstance(v, torch.Tensor):
        return [float(x) for x in v.detach().cpu().tolist()]
    return [float(x) for x in v]


@dataclass
class RaySamples:
    """render.py:114-290."""

    t_min: torch.Tensor
    t_max: torch.Tensor
    mask: torch.Tensor
    ts: torch.Tensor

    def _mask_u8(self) -> torch.Tensor:
        return self.mask.to(torch.uint8) if self.mask.dtype != torch.uint8 else self.mask

    @classmethod
    def stratified_sampling(cls, t_min, t_max, mask, count: int, key: KeyOrUniforms) -> "RaySamples":
        """render.py:121-143."""
        n = t_min.shape[0]
        u = _uniforms(key, (n, count), t_min.device)
        ts = _native.stratified(t_min.contiguous(), t_max.contiguous(), u)
        return cls(t_min=t_min, t_max=t_max, mask=mask, ts=ts)

    def points(self, rays: torch.Tensor) -> torch.Tensor:
        """render.py:145-153 (host helper; the kernels form points in-line)."""
        return rays[:, :1] + (rays[:, 1:] * self.ts[:, :, None])

    def starts(self) -> torch.Tensor:
        """render.py:259-261 -> [N,T]."""
        return _native.ray_intervals(self.ts, self.t_min.contiguous(), self.t_max.contiguous(), ("starts",))["starts"]

    def ends(self) -> torch.Tensor:
        """render.py:263-265 -> [N,T]."""
        return _native.ray_intervals(self.ts, self.t_min.contiguous(), self.t_max.contiguous(), ("ends",))["ends"]

    def deltas(self) -> torch.Tensor:
        """render.py:267-268 -> [N,T]."""
        return _native.ray_intervals(self.ts, self.t_min.contiguous(), self.t_max.contiguous(), ("deltas",))["deltas"]

    def termination_probs(self, densities: torch.Tensor) -> torch.Tensor:
        """render.py:270-287 -> [N,T+1]; the last column is the probability of reaching the background."""
        return _native.termination_probs(self.ts, self.t_min.contiguous(), self.t_max.contiguous(),
                                         densities.contiguous())

    def average_aux_losses(self, densities: torch.Tensor, aux: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """render.py:192-209 -> {name: scalar}."""
        return average_aux_losses(self, densities.contiguous(), aux)

    def render_rays(self, densities, rgbs, background, _rays=None) -> torch.Tensor:
        """render.py:155-176 -> [N,3]."""
        rays = _rays if _rays is not None else _dummy_rays(self.ts)
        out, _, _ = _native.composite_fwd(rays, self.ts, self.t_min, self.t_max, self._mask_u8(),
                                          densities.contiguous(), rgbs.contiguous(),
                                          background.contiguous(), want_aux=False)
        return out

    def render_alpha(self, densities) -> torch.Tensor:
        """render.py:178-190 -> [N,1]."""
        n, T = self.ts.shape
        zeros = torch.zeros(n, T, 3, device=self.ts.device)
        _, alphas, _ = _native.composite_fwd(_dummy_rays(self.ts), self.ts, self.t_min, self.t_max,
                                             self._mask_u8(), densities.contiguous(), zeros,
                                             torch.zeros(3, device=self.ts.device))
        return alphas

    def fine_sampling(self, count: int, key: KeyOrUniforms, densities: torch.Tensor,
                      combine: bool = True, eps: float = 1e-8) -> "RaySamples":
        """render.py:211-257."""
        n = self.ts.shape[0]
        u = _uniforms(key, (n, count), self.ts.device)
        if combine:
            new_ts = _native.sample_fine(self.ts, densities.contiguous(), self.t_min, self.t_max, u, eps)
        else:
            _, _, new_ts = _native.sample_fine(self.ts, densities.contiguous(), self.t_min, self.t_max,
                                               u, eps, debug=True)
        return RaySamples(t_min=self.t_min, t_max=self.t_max, mask=self.mask, ts=new_ts)


def _dummy_rays(ts: torch.Tensor) -> torch.Tensor:
    return torch.zeros(ts.shape[0], 2, 3, device=ts.device)


def render_rays(model: ModelBase, params: Any, background: torch.Tensor, batch: torch.Tensor,
                ts: RaySamples, _save: Optional[str] = None
                ) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """Free function render.py:293-343 -> (dict(outputs, rgbs, densities, alphas, coords), aux)."""
    n, T = ts.ts.shape
    ctx = None
    if hasattr(model, "apply_rays"):
        densities, rgbs, aux, ctx = model.apply_rays(params, batch, ts.ts, save=_save is not None,
                                                     slot=_save)
    else:  # any other ModelBase: the reference seam model.apply(dict(params=params), x, d)
        all_points = ts.points(batch)
        direction_batch = batch[:, 1:2].expand(n, T, 3)
        densities, rgbs, aux = model.apply(dict(params=params), all_points.reshape(-1, 3),
                                           direction_batch.reshape(-1, 3))
        densities = densities.reshape(n, T).contiguous()
        rgbs = rgbs.reshape(n, T, 3).contiguous()
        aux = {k: v.reshape(n, T) for k, v in aux.items()}
    outputs, alphas, coords = _native.composite_fwd(batch, ts.ts, ts.t_min, ts.t_max, ts._mask_u8(),
                                                    densities, rgbs, background.contiguous())
    aux_mean = {}
    if aux:
        aux_mean = ts.average_aux_losses(densities, aux)
    out = dict(outputs=outputs, rgbs=rgbs, densities=densities, alphas=alphas, coords=coords)
    if _save is not None:
        out["_ctx"] = ctx
        out["_ts"] = ts
        out["_aux"] = aux
    return out, aux_mean


def _pack_aux(aux: Dict[str, torch.Tensor]):
    """Up to three per-sample aux values as the colour channels of one [N,T,3] tensor."""
    names = sorted(aux)
    if len(names) > 3:
        raise _native.LnrfError("at most three aux losses per model are supported")
    first = aux[names[0]]
    cols = torch.zeros(first.shape + (3,), device=first.device)
    for i, k in enumerate(names):
        cols[..., i] = aux[k]
    return names, cols


def average_aux_losses(ts: RaySamples, densities, aux: Dict[str, torch.Tensor]):
    """render.py:192-209 via the compositing kernel: sum_t v*p is one colour channel of a
    composite with the aux values as colours and a zero background; then the mean over rays."""
    names, cols = _pack_aux(aux)
    zero_bg = torch.zeros(3, device=ts.ts.device)
    comp, _, _ = _native.composite_fwd(_dummy_rays(ts.ts), ts.ts, ts.t_min, ts.t_max, ts._mask_u8(),
                                       densities, cols, zero_bg, want_aux=False)
    return {k: comp[:, i].mean() for i, k in enumerate(names)}


def ray_t_range(bbox: torch.Tensor, ray: torch.Tensor, min_t_range: float = 1e-3,
                epsilon: float = 1e-8):
    """Single-ray form, render.py:346-389 -> (ts[2], mask scalar)."""
    t_min, t_max, mask, _ = _native.sample_coarse(ray.reshape(1, 2, 3).contiguous(), _vec3(bbox[0]),
                                                  _vec3(bbox[1]),
                                                  torch.zeros(1, 1, device=ray.device), min_t_range,
                                                  epsilon)
    return torch.stack([t_min[0], t_max[0]]), mask[0].bool()
