"""Host-side mirror of learn_nerf/ref_nerf.py: RefNERFModel (sh_degree = 4).

Same constructor fields as the reference (ref_nerf.py:80-90) and the same model contract
(``(x[N,3], d[N,3]) -> (density[N,1], rgb[N,3], {"normal_mse": [N], "neg_normal": [N]})``,
ref_nerf.py:34-77).  The arithmetic -- spatial MLP, the input-gradient normals, integrated
directional encoding, directional block, sRGB colour, aux losses and the full backward incl.
the second-order term through the normals -- runs in liblnrf.so (lnrf_refnerf_fwd / _bwd).
Parameters keep the Flax names ``Dense_0..10`` as views into one flat fp32 buffer.
"""
import math
from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch

from . import _native
from .model import ModelBase, ParamTree, _rng_key, init_flat_

HARMONIC_COUNTS = [1, 3, 5, 7, 9, 11, 13, 15]  # ref_nerf.py:15


@dataclass
class RefNERFModel(ModelBase):
    """ref_nerf.py:80-107 on RefNERFBase (:19-77)."""

    sh_degree: int = 4
    input_layers: int = 5
    mid_layers: int = 4
    hidden_dim: int = 256
    color_layer_dim: int = 128
    x_freqs: int = 10
    d_freqs: int = 4
    precision: str = "fp32"

    aux_names = ("normal_mse", "neg_normal")

    def _check_arch(self):
        assert 1 <= self.sh_degree <= 8  # ref_nerf.py:154
        if (self.sh_degree, self.input_layers, self.mid_layers, self.hidden_dim, self.color_layer_dim,
                self.x_freqs) != (4, 5, 4, 256, 128, 10):
            raise _native.LnrfError("liblnrf implements RefNERFModel(sh_degree=4) with the default "
                                    "5+4 x 256 spatial block and a 128-wide directional block only")
        if self.precision != "fp32":
            raise _native.LnrfError("RefNERFModel runs on the fp32 path only")

    def layer_dims(self):
        xe, h = 6 * self.x_freqs, self.hidden_dim
        dims = [(xe, h)] + [(h, h)] * (self.input_layers - 1)
        dims += [(h + xe, h)] + [(h, h)] * (self.mid_layers - 1)
        enc = sum(HARMONIC_COUNTS[: self.sh_degree])
        dims += [(h + enc + 1, self.color_layer_dim), (self.color_layer_dim, 3)]
        return dims

    def param_floats(self) -> int:
        self._check_arch()
        return _native.refnerf_param_floats()

    def param_count(self) -> int:
        return sum(a * b + b for a, b in self.layer_dims())

    def bind(self, flat: torch.Tensor) -> ParamTree:
        self._check_arch()
        offs = _native.refnerf_param_offsets()
        tree = ParamTree()
        for i, (a, b) in enumerate(self.layer_dims()):
            tree[f"Dense_{i}"] = dict(kernel=flat[offs[2 * i]: offs[2 * i] + a * b].view(a, b),
                                      bias=flat[offs[2 * i + 1]: offs[2 * i + 1] + b])
        tree.flat = flat
        return tree

    def flatten_params(self, params: Dict[str, Any], device=None) -> ParamTree:
        if isinstance(params, ParamTree) and params.flat is not None:
            return params
        first = params["Dense_0"]["kernel"]
        device = device or (first.device if isinstance(first, torch.Tensor) else "cuda")
        tree = self.bind(torch.zeros(self.param_floats(), device=device))
        for name, leaf in tree.items():
            for k in ("kernel", "bias"):
                leaf[k].copy_(torch.as_tensor(params[name][k], dtype=torch.float32))
        return tree

    def init(self, rngs, x=None, d=None, device=None, flat: Optional[torch.Tensor] = None):
        device = torch.device(device or (x.device if isinstance(x, torch.Tensor) else "cuda"))
        if flat is None:
            flat = torch.zeros(self.param_floats(), device=device)
        else:
            flat.zero_()
        tree = self.bind(flat)
        offs = _native.refnerf_param_offsets()
        init_flat_(flat, [(offs[2 * i], a, a * b) for i, (a, b) in enumerate(self.layer_dims())], [],
                   _rng_key(rngs))
        return {"params": tree}

    # ------------------------------------------------------------------ native calls
    def _workspace(self, m: int, save: bool, device, slot=None) -> torch.Tensor:
        nbytes = _native.refnerf_workspace_bytes(m, save)
        cache = self.__dict__.setdefault("_ws_cache", {})
        key = (str(device), bool(save), slot)
        ws = cache.get(key)
        if ws is None or ws.numel() < nbytes:
            raw = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
            shift = (-raw.data_ptr()) % 256
            ws = raw[shift: shift + nbytes]
            cache[key] = ws
        return ws

    def _forward(self, tree, x, d, rays, ts, n, T, save, slot=None):
        dev = tree.flat.device
        m = n * T
        dens = torch.empty(m, device=dev)
        rgb = torch.empty(m, 3, device=dev)
        aux_mse = torch.empty(m, device=dev)
        aux_neg = torch.empty(m, device=dev)
        ws = self._workspace(m, save, dev, slot)
        _native.refnerf_fwd(tree.flat, x, d, rays, ts, n, T, save, ws, dens, rgb, aux_mse, aux_neg)
        return dens, rgb, aux_mse, aux_neg, ws

    def apply(self, variables, x: torch.Tensor, d: torch.Tensor):
        tree = self.flatten_params(variables["params"])
        x = _native._f32c(x.contiguous(), "x")
        d = _native._f32c(d.contiguous(), "d")
        dens, rgb, a1, a2, _ = self._forward(tree, x, d, None, None, x.shape[0], 1, save=False)
        return dens[:, None], rgb, dict(normal_mse=a1, neg_normal=a2)

    def apply_rays(self, params, rays, ts, save: bool = False, slot=None):
        tree = self.flatten_params(params)
        n, T = ts.shape
        rays, ts = _native._f32c(rays, "rays"), _native._f32c(ts, "ts")
        dens, rgb, a1, a2, ws = self._forward(tree, None, None, rays, ts, n, T, save, slot)
        ctx = dict(tree=tree, ws=ws, rays=rays, ts=ts, n=n, T=T) if save else None
        return dens.view(n, T), rgb.view(n, T, 3), dict(normal_mse=a1.view(n, T), neg_normal=a2.view(n, T)), ctx

    def backward_rays(self, ctx, d_dens, d_rgb, d_flat, d_aux=None):
        m = ctx["n"] * ctx["T"]
        zeros = None
        def aux(name):
            nonlocal zeros
            if d_aux is not None and name in d_aux:
                return d_aux[name].reshape(-1)
            if zeros is None:
                zeros = torch.zeros(m, device=d_flat.device)
            return zeros
        _native.refnerf_bwd(ctx["tree"].flat, None, None, ctx["rays"], ctx["ts"], ctx["n"], ctx["T"], ctx["ws"],
                            d_dens.reshape(-1), d_rgb.reshape(-1, 3), aux("normal_mse"), aux("neg_normal"),
                            d_flat)
