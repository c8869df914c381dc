"""Host-side mirror of learn_nerf/model.py: ModelBase, NeRFModel, sinusoidal_emb.

``NeRFModel.apply(dict(params=params), x, d)`` keeps the reference signature and
return contract (model.py:12-27); the arithmetic runs in liblnrf.so
(lnrf_nerf_mlp_fwd / _bwd).  Parameters are a Flax-shaped tree
(``Dense_i/{kernel[in,out], bias[out]}``) whose leaves are views into one flat fp32
CUDA buffer laid out as the C ABI expects (include/lnrf.h).
"""
import math
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import torch

from . import _native
from .prng import KeyLike, _as_key


class ParamTree(dict):
    """Nested dict of tensors that are views of ``flat`` (one contiguous buffer)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.flat: Optional[torch.Tensor] = None
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None
        self._manual_version = 0

    def mark_updated(self):
        """Call after the flat buffer was modified by a native kernel (Adam)."""
        self._manual_version += 1

    def version_key(self):
        return (self.flat._version, self._manual_version, self.flat.data_ptr())


class ModelBase:
    """model.py:7-27: ``(x[N,3], d[N,3]) -> (density[N,1], rgb[N,3], aux{name: [N]})``."""

    def init(self, rngs, x=None, d=None, device=None) -> Dict[str, Any]:
        raise NotImplementedError

    def apply(self, variables: Dict[str, Any], x: torch.Tensor, d: torch.Tensor
              ) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, torch.Tensor]]:
        raise NotImplementedError

    def __call__(self, x, d):
        raise NotImplementedError("call model.apply(dict(params=params), x, d)")


def sinusoidal_emb(coords: torch.Tensor, freqs: int) -> torch.Tensor:
    """model.py:65-77 (host-side torch helper; the kernels compute this in-line)."""
    coeffs = 2.0 ** torch.arange(freqs, dtype=torch.float32, device=coords.device)
    inputs = coords[..., None] * coeffs
    combined = torch.cat([torch.sin(inputs), torch.cos(inputs)], dim=-1)
    return combined.reshape(combined.shape[:-2] + (-1,))


# flax nn.Dense default kernel_init = lecun_normal() = variance_scaling(1.0, "fan_in", "truncated_normal"):
# stddev = sqrt(1 / fan_in) / 0.87962566103423978 (the std of a unit normal truncated at +-2), applied to
# jax.random.truncated_normal(key, -2, 2) = sqrt(2) * erfinv(U(erf(-sqrt 2), erf(sqrt 2))).
_TRUNC_STD = 0.87962566103423978
_ERF_SQRT2 = math.erf(math.sqrt(2.0))


def init_flat_(flat: torch.Tensor, dense_kernels, tables, key):
    """Random-init a flat parameter buffer (one Threefry uniform stream over the whole buffer, evaluated on
    the host, one upload): Dense kernels get lecun_normal as Flax does (see above),
    hash tables 1e-4 * (2 U - 1) (instant_ngp.py:181-186), biases and padding stay zero.
    ``dense_kernels``: [(offset, fan_in, count)], ``tables``: [(offset, count)].
    The per-parameter key derivation of flax (path-hashed fold_in) is not reproduced: the stream is
    ``uniform(key, [len(flat)])`` indexed by buffer position."""
    import numpy as np
    from . import prng
    n = flat.numel()
    std = np.zeros(n, np.float32)
    tab = np.zeros(n, np.float32)
    for off, fan_in, count in dense_kernels:
        std[off: off + count] = math.sqrt(1.0 / fan_in) / _TRUNC_STD
    for off, count in tables:
        tab[off: off + count] = 1e-4
    # Runs on the HOST (numpy Threefry stream + scipy erfinv), then one H2D copy: initialisation is not the
    # hot path, and torch's CUDA erfinv is a run-time-compiled (NVRTC "jiterator") kernel.
    from scipy.special import erfinv
    k = prng._as_key(key)
    bits = prng.random_bits_host(k, n)
    u = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)  # jax.random.uniform
    # truncated normal by inversion; clamped inside the open interval (-2, 2) as jax.random does
    dense = std > 0
    host = np.zeros(n, np.float32)
    z = np.clip(math.sqrt(2.0) * erfinv(u[dense].astype(np.float64) * (2.0 * _ERF_SQRT2) - _ERF_SQRT2), -1.9999999, 1.9999999)
    host[dense] = (z * std[dense]).astype(np.float32)
    if tables:
        host += ((u * np.float32(2.0) - np.float32(1.0)) * tab).astype(np.float32)
    flat.copy_(torch.from_numpy(host))
    return flat


def _rng_key(rngs):
    """``model.init(dict(params=rng), ...)`` / ``model.init(rng, ...)`` -> the PRNG key."""
    if isinstance(rngs, dict):
        rngs = rngs["params"]
    return _as_key(rngs)


@dataclass
class NeRFModel(ModelBase):
    """model.py:30-62.  ``precision``: "fp32" (split-fp16 tcgen05 GEMMs, 1e-5) or "bf16" (fused tcgen05 kernels, 2e-2)."""

    input_layers: int = 5
    mid_layers: int = 4
    hidden_dim: int = 256
    color_layer_dim: int = 128
    x_freqs: int = 10
    d_freqs: int = 4
    precision: str = "fp32"

    # ------------------------------------------------------------------ layout
    def _check_arch(self):
        if (self.input_layers, self.mid_layers, self.hidden_dim, self.color_layer_dim, self.x_freqs,
                self.d_freqs) != (5, 4, 256, 128, 10, 4):
            raise _native.LnrfError("liblnrf implements the default NeRFModel architecture only "
                                    "(5+4 x 256, colour 128, x_freqs 10, d_freqs 4)")
        if self.precision not in _native.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_native.PRECISIONS)}")

    def layer_dims(self):
        xe, de, h = 6 * self.x_freqs, 6 * self.d_freqs, self.hidden_dim
        dims = [(xe, h)] + [(h, h)] * (self.input_layers - 1)
        dims += [(h + xe, h)] + [(h, h)] * (self.mid_layers - 1)
        dims += [(h, 1), (h + de, self.color_layer_dim), (self.color_layer_dim, 3)]
        return dims

    def param_floats(self) -> int:
        self._check_arch()
        return _native.nerf_param_floats()

    def param_count(self) -> int:
        return sum(a * b + b for a, b in self.layer_dims())

    def bind(self, flat: torch.Tensor) -> ParamTree:
        """Flax-shaped tree of views into ``flat`` (len == param_floats())."""
        self._check_arch()
        offs = _native.nerf_param_offsets()
        tree = ParamTree()
        for i, (a, b) in enumerate(self.layer_dims()):
            tree[f"Dense_{i}"] = dict(kernel=flat[offs[2 * i]: offs[2 * i] + a * b].view(a, b),
                                      bias=flat[offs[2 * i + 1]: offs[2 * i + 1] + b])
        tree.flat = flat
        return tree

    def flatten_params(self, params: Dict[str, Any], device=None) -> ParamTree:
        """Accept a plain nested dict (e.g. an un-pickled checkpoint) and re-home it."""
        if isinstance(params, ParamTree) and params.flat is not None:
            return params
        first = params["Dense_0"]["kernel"]
        device = device or (first.device if isinstance(first, torch.Tensor) else "cuda")
        flat = torch.zeros(self.param_floats(), device=device)
        tree = self.bind(flat)
        for name, leaf in tree.items():
            for k in ("kernel", "bias"):
                leaf[k].copy_(torch.as_tensor(params[name][k], dtype=torch.float32))
        return tree

    def init(self, rngs, x=None, d=None, device=None, flat: Optional[torch.Tensor] = None):
        """Mirror of ``model.init(dict(params=rng), x, d)`` -> ``{"params": tree}``."""
        device = torch.device(device or (x.device if isinstance(x, torch.Tensor) else "cuda"))
        if flat is None:
            flat = torch.zeros(self.param_floats(), device=device)
        else:
            flat.zero_()
        tree = self.bind(flat)
        offs = _native.nerf_param_offsets()
        init_flat_(flat, [(offs[2 * i], a, a * b) for i, (a, b) in enumerate(self.layer_dims())], [],
                   _rng_key(rngs))
        return {"params": tree}

    # ------------------------------------------------------------------ native calls
    def _packed(self, tree: ParamTree) -> Optional[torch.Tensor]:
        if self.precision != "bf16":
            return None
        key = tree.version_key()
        if tree._packed is None or tree._packed_key != key:
            if tree._packed is None:
                nbytes = _native.nerf_packed_bytes()
                raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=tree.flat.device)
                shift = (-raw.data_ptr()) % 1024
                tree._packed = raw[shift: shift + nbytes]
            _native.nerf_pack_weights(tree.flat, tree._packed)
            tree._packed_key = key
        return tree._packed

    def _workspace(self, m: int, save: bool, device, slot=None) -> Optional[torch.Tensor]:
        nbytes = _native.nerf_mlp_workspace_bytes(m, _native.PRECISIONS[self.precision], save)
        if nbytes == 0:
            return None
        cache = self.__dict__.setdefault("_ws_cache", {})
        key = (str(device), bool(save), slot)
        ws = cache.get(key)
        if ws is None or ws.numel() < nbytes:
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            shift = (-raw.data_ptr()) % 1024  # the bf16 stash holds 1024-byte aligned tile images
            ws = raw[shift: shift + nbytes]
            cache[key] = ws
        return ws

    def _forward(self, tree: ParamTree, x, d, rays, ts, n: int, T: int, save: bool, slot=None):
        self._check_arch()
        dev = tree.flat.device
        m = n * T
        dens = torch.empty(m, device=dev)
        rgb = torch.empty(m, 3, device=dev)
        ws = self._workspace(m, save, dev, slot)
        _native.nerf_mlp_fwd(tree.flat, self._packed(tree), x, d, rays, ts, n, T,
                             _native.PRECISIONS[self.precision], save, ws, dens, rgb)
        return dens, rgb, ws

    def apply(self, variables, x: torch.Tensor, d: torch.Tensor):
        """model.apply(dict(params=params), x[N,3], d[N,3]) -> (density[N,1], rgb[N,3], {})."""
        tree = self.flatten_params(variables["params"])
        x = _native._f32c(x.contiguous(), "x")
        d = _native._f32c(d.contiguous(), "d")
        dens, rgb, _ = self._forward(tree, x, d, None, None, x.shape[0], 1, save=False)
        return dens[:, None], rgb, {}

    def apply_rays(self, params, rays: torch.Tensor, ts: torch.Tensor, save: bool = False, slot=None):
        """Fused seam used by render_rays: points/directions are formed in-kernel from
        rays[N,2,3] and ts[N,T] (render.py:318-324).  Returns dens[N,T], rgb[N,T,3], aux, ctx.
        ``slot`` names the saved-activation workspace so that one model instance can hold
        several forward passes (coarse and fine) until their backward runs."""
        tree = self.flatten_params(params)
        n, T = ts.shape
        dens, rgb, ws = self._forward(tree, None, None, _native._f32c(rays, "rays"),
                                      _native._f32c(ts, "ts"), n, T, save, slot)
        ctx = dict(tree=tree, ws=ws, m=n * T, dens=dens, rgb=rgb) if save else None
        return dens.view(n, T), rgb.view(n, T, 3), {}, ctx

    def backward_rays(self, ctx, d_dens: torch.Tensor, d_rgb: torch.Tensor, d_flat: torch.Tensor,
                      d_aux=None):
        """Accumulate dL/dparams into ``d_flat`` (same layout as the flat params)."""
        tree = ctx["tree"]
        _native.nerf_mlp_bwd(tree.flat, self._packed(tree), ctx["m"],
                             _native.PRECISIONS[self.precision], ctx["ws"], ctx["dens"], ctx["rgb"],
                             d_dens.reshape(-1), d_rgb.reshape(-1, 3), d_flat)
