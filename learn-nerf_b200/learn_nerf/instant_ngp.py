"""Host-side mirror of learn_nerf/instant_ngp.py: InstantNGPModel (+ the encodings).

Same constructor fields as the reference (instant_ngp.py:16-31); parameters keep the Flax
names (``Dense_0..4`` and ``MultiresHashTableEncoding_0/HashTableEncoding_l/table``) as
views into one flat fp32 buffer ``[MLP | table_0 | table_1 | ...]``.  The arithmetic runs
in liblnrf.so: lnrf_hashgrid_fwd/_bwd (K7/K8) and lnrf_ngp_mlp_fwd/_bwd.
``InstantNGPRefNERFModel`` (instant_ngp.py:57-89) is the Ref-NeRF variant on a smooth hash grid
(lnrf_ngpref_fwd/_bwd).
"""
import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import torch

from . import _native
from .model import ModelBase, ParamTree, _rng_key, init_flat_


def hash_table_lookup(table: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """instant_ngp.py:211-224 (host helper for tools/tests; the kernels hash in-line)."""
    c = coords.to(torch.int64) & 0xFFFFFFFF
    idx = (c[:, 0] ^ ((19_349_663 * c[:, 1]) & 0xFFFFFFFF) ^ ((83_492_791 * c[:, 2]) & 0xFFFFFFFF))
    return table[idx % table.shape[0]]


@dataclass
class InstantNGPModel(ModelBase):
    """instant_ngp.py:16-54.  ``precision``: "fp32" (fp32-accurate heads: split-fp16 tcgen05 GEMMs when training, fused FFMA kernel for the workspace-free forward; 1e-5) or "bf16" (heads on the tcgen05
    tensor cores, 2e-2 abs on density / rgb); the hash grid itself is fp32 on both paths."""

    table_sizes: List[int]
    grid_sizes: List[int]
    bbox_min: Any
    bbox_max: Any
    table_feature_dim: int = 2
    table_smooth: bool = False
    d_freqs: int = 4
    hidden_dim: int = 64
    density_dim: int = 16
    density_layers: int = 1
    color_layers: int = 2
    precision: str = "fp32"
    _spec: Optional[_native.GridSpec] = field(default=None, repr=False, compare=False)

    def _check_arch(self):
        if (self.table_feature_dim, self.d_freqs, self.hidden_dim, self.density_dim,
                self.density_layers, self.color_layers) != (2, 4, 64, 16, 1, 2):
            raise _native.LnrfError("liblnrf implements the default InstantNGPModel head sizes only "
                                    "(F=2, hidden 64, density_dim 16, 1 density + 2 colour layers)")
        if len(self.grid_sizes) != len(self.table_sizes) or not 1 <= len(self.grid_sizes) <= 16:
            raise _native.LnrfError("1..16 levels supported")
        if self.precision not in _native.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_native.PRECISIONS)}")

    @property
    def L(self) -> int:
        return len(self.grid_sizes)

    def spec(self) -> _native.GridSpec:
        if self._spec is None:
            self._check_arch()
            def v3(v):
                return [float(x) for x in (v.tolist() if hasattr(v, "tolist") else v)]
            self._spec = _native.GridSpec(self.table_sizes, self.grid_sizes, v3(self.bbox_min),
                                          v3(self.bbox_max), self.table_smooth,
                                          base_offset=_native.ngp_mlp_param_floats(self.L))
        return self._spec

    def layer_dims(self):
        return [(2 * self.L, 64), (64, 16), (24 + 16, 64), (64, 64), (64, 3)]

    def param_floats(self) -> int:
        return self.spec().end

    def param_count(self) -> int:
        return sum(a * b + b for a, b in self.layer_dims()) + sum(r * 2 for r in self.spec().rows)

    def bind(self, flat: torch.Tensor) -> ParamTree:
        spec = self.spec()
        offs = _native.ngp_mlp_param_offsets(self.L)
        tree = ParamTree()
        for i, (a, b) in enumerate(self.layer_dims()):
            tree[f"Dense_{i}"] = dict(kernel=flat[offs[2 * i]: offs[2 * i] + a * b].view(a, b),
                                      bias=flat[offs[2 * i + 1]: offs[2 * i + 1] + b])
        tree["MultiresHashTableEncoding_0"] = {
            f"HashTableEncoding_{l}": dict(table=flat[o: o + r * 2].view(r, 2))
            for l, (o, r) in enumerate(zip(spec.offsets, spec.rows))}
        tree.flat = flat
        return tree

    def flatten_params(self, params: Dict[str, Any], device=None) -> ParamTree:
        if isinstance(params, ParamTree) and params.flat is not None:
            return params
        first = params["Dense_0"]["kernel"]
        device = device or (first.device if isinstance(first, torch.Tensor) else "cuda")
        tree = self.bind(torch.zeros(self.param_floats(), device=device))
        def put(dst, src):
            for k, v in dst.items():
                if isinstance(v, dict):
                    put(v, src[k])
                else:
                    v.copy_(torch.as_tensor(src[k], dtype=torch.float32))
        put(tree, params)
        return tree

    def init(self, rngs, x=None, d=None, device=None, flat: Optional[torch.Tensor] = None):
        device = torch.device(device or (x.device if isinstance(x, torch.Tensor) else "cuda"))
        if flat is None:
            flat = torch.zeros(self.param_floats(), device=device)
        else:
            flat.zero_()
        tree = self.bind(flat)
        spec = self.spec()
        offs = (_native.ngpref_param_offsets if isinstance(self, InstantNGPRefNERFModel)
                else _native.ngp_mlp_param_offsets)(self.L)
        init_flat_(flat, [(offs[2 * i], a, a * b) for i, (a, b) in enumerate(self.layer_dims())],
                   [(o, r * 2) for o, r in zip(spec.offsets, spec.rows)], _rng_key(rngs))  # tables: :181-186
        return {"params": tree}

    # ------------------------------------------------------------------ native calls
    def _packed(self, tree: ParamTree) -> Optional[torch.Tensor]:
        """bf16 operand images of the head weights (bf16 path), rebuilt when the parameters changed."""
        if self.precision != "bf16":
            return None
        key = tree.version_key()
        if tree._packed is None or tree._packed_key != key:
            if tree._packed is None:
                nbytes = _native.ngp_packed_bytes()
                raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=tree.flat.device)
                shift = (-raw.data_ptr()) % 1024
                tree._packed = raw[shift: shift + nbytes]
            _native.ngp_pack_weights(tree.flat, self.L, tree._packed)
            tree._packed_key = key
        return tree._packed

    def _workspace(self, m: int, device, slot, save):
        cache = self.__dict__.setdefault("_ws_cache", {})
        bf16 = self.precision == "bf16"
        nbytes = 0
        if save:
            nbytes = _native.ngp_mlp_tc_workspace_bytes(m) if bf16 else _native.ngp_mlp_workspace_bytes(m, self.L)
        ws = cache.get((str(device), slot, save))
        if ws is None or ws[0].numel() < nbytes or ws[1].shape[0] < m:
            raw = torch.empty(max(nbytes, 256) + 1024, dtype=torch.uint8, device=device)
            shift = (-raw.data_ptr()) % 1024  # the bf16 stash holds 1024-byte aligned tile images
            ws = (raw[shift: shift + max(nbytes, 256)], torch.empty(m, 2 * self.L, device=device))
            cache[(str(device), slot, save)] = ws
        return ws

    def _forward(self, tree, x, d, rays, ts, n, T, slot=None, save=False):
        spec = self.spec()
        dev = tree.flat.device
        m = n * T
        ws, enc = self._workspace(m, dev, slot, save)
        enc = enc[:m]
        _native.hashgrid_fwd(tree.flat, spec, x, rays, ts, n, T, enc)
        dens = torch.empty(m, device=dev)
        rgb = torch.empty(m, 3, device=dev)
        if self.precision == "bf16":
            _native.ngp_mlp_fwd_tc(self._packed(tree), self.L, enc, d, rays, n, T, save, ws if save else None, dens, rgb)
        else:
            _native.ngp_mlp_fwd(tree.flat, self.L, enc, d, rays, n, T, save, ws if save else None, dens, rgb)
        return dens, rgb, ws, enc

    def encode(self, params, x: torch.Tensor) -> torch.Tensor:
        """MultiresHashTableEncoding(x) -> [N, 2L] (instant_ngp.py:92-118)."""
        tree = self.flatten_params(params)
        x = _native._f32c(x.contiguous(), "x")
        enc = torch.empty(x.shape[0], 2 * self.L, device=x.device)
        _native.hashgrid_fwd(tree.flat, self.spec(), x, None, None, x.shape[0], 1, enc)
        return enc

    def apply(self, variables, x: torch.Tensor, d: torch.Tensor):
        tree = self.flatten_params(variables["params"])
        x = _native._f32c(x.contiguous(), "x")
        d = _native._f32c(d.contiguous(), "d")
        dens, rgb, _, _ = self._forward(tree, x, d, None, None, x.shape[0], 1)
        return dens[:, None], rgb, {}

    def apply_rays(self, params, rays, ts, save: bool = False, slot=None):
        tree = self.flatten_params(params)
        n, T = ts.shape
        rays, ts = _native._f32c(rays, "rays"), _native._f32c(ts, "ts")
        dens, rgb, ws, enc = self._forward(tree, None, None, rays, ts, n, T, slot if save else None, save)
        ctx = dict(tree=tree, ws=ws, enc=enc, rays=rays, ts=ts, n=n, T=T, dens=dens, rgb=rgb) if save else None
        return dens.view(n, T), rgb.view(n, T, 3), {}, ctx

    def backward_rays(self, ctx, d_dens, d_rgb, d_flat, d_aux=None):
        tree = ctx["tree"]
        m = ctx["n"] * ctx["T"]
        d_enc = torch.empty_like(ctx["enc"])
        if self.precision == "bf16":
            _native.ngp_mlp_bwd_tc(self._packed(tree), self.L, m, ctx["ws"], ctx["dens"], ctx["rgb"],
                                   d_dens.reshape(-1), d_rgb.reshape(-1, 3), d_flat, d_enc)
        else:
            _native.ngp_mlp_bwd(tree.flat, self.L, ctx["enc"], m, ctx["ws"], ctx["dens"], ctx["rgb"],
                                d_dens.reshape(-1), d_rgb.reshape(-1, 3), d_flat, d_enc)
        _native.hashgrid_bwd(self.spec(), None, ctx["rays"], ctx["ts"], ctx["n"], ctx["T"], d_enc, d_flat)


@dataclass
class InstantNGPRefNERFModel(ModelBase):
    """instant_ngp.py:57-89 on RefNERFBase (ref_nerf.py:19-77, sh_degree = 4): smooth hash grid ->
    Dense(64) ReLU -> Dense(16) as the spatial block, Dense(64) ReLU x2 -> Dense(3) as the
    directional block.  Same model contract as RefNERFModel:
    ``(x[N,3], d[N,3]) -> (density[N,1], rgb[N,3], {"normal_mse": [N], "neg_normal": [N]})``."""

    table_sizes: List[int]
    grid_sizes: List[int]
    bbox_min: Any
    bbox_max: Any
    sh_degree: int = 4
    table_feature_dim: int = 2
    d_freqs: int = 4
    hidden_dim: int = 64
    density_dim: int = 16
    density_layers: int = 1
    color_layers: int = 2
    precision: str = "fp32"
    _spec: Optional[_native.GridSpec] = field(default=None, repr=False, compare=False)

    aux_names = ("normal_mse", "neg_normal")

    def _check_arch(self):
        assert 1 <= self.sh_degree <= 8  # ref_nerf.py:154
        if (self.sh_degree, self.table_feature_dim, self.hidden_dim, self.density_dim, self.density_layers,
                self.color_layers) != (4, 2, 64, 16, 1, 2):
            raise _native.LnrfError("liblnrf implements InstantNGPRefNERFModel(sh_degree=4) with the default "
                                    "head sizes only (F=2, hidden 64, density_dim 16, 1 + 2 layers)")
        if len(self.grid_sizes) != len(self.table_sizes) or not 1 <= len(self.grid_sizes) <= 16:
            raise _native.LnrfError("1..16 levels supported")
        if self.precision != "fp32":
            raise _native.LnrfError("InstantNGPRefNERFModel runs on the fp32 path only")

    @property
    def L(self) -> int:
        return len(self.grid_sizes)

    def spec(self) -> _native.GridSpec:
        if self._spec is None:
            self._check_arch()
            def v3(v):
                return [float(x) for x in (v.tolist() if hasattr(v, "tolist") else v)]
            self._spec = _native.GridSpec(self.table_sizes, self.grid_sizes, v3(self.bbox_min), v3(self.bbox_max),
                                          True, base_offset=_native.ngpref_mlp_param_floats(self.L))  # smooth=True :79
        return self._spec

    def layer_dims(self):
        return [(2 * self.L, 64), (64, 16), (16 + 16 + 1, 64), (64, 64), (64, 3)]

    def param_floats(self) -> int:
        return self.spec().end

    def param_count(self) -> int:
        return sum(a * b + b for a, b in self.layer_dims()) + sum(r * 2 for r in self.spec().rows)

    def bind(self, flat: torch.Tensor) -> ParamTree:
        spec = self.spec()
        offs = _native.ngpref_param_offsets(self.L)
        tree = ParamTree()
        for i, (a, b) in enumerate(self.layer_dims()):  # Dense_2 owns 36 rows in the buffer: 3 zero pads follow
            tree[f"Dense_{i}"] = dict(kernel=flat[offs[2 * i]: offs[2 * i] + a * b].view(a, b),
                                      bias=flat[offs[2 * i + 1]: offs[2 * i + 1] + b])
        tree["MultiresHashTableEncoding_0"] = {
            f"HashTableEncoding_{l}": dict(table=flat[o: o + r * 2].view(r, 2))
            for l, (o, r) in enumerate(zip(spec.offsets, spec.rows))}
        tree.flat = flat
        return tree

    flatten_params = InstantNGPModel.flatten_params
    init = InstantNGPModel.init

    # ------------------------------------------------------------------ native calls
    def _workspace(self, m: int, save: bool, device, slot=None) -> torch.Tensor:
        nbytes = _native.ngpref_workspace_bytes(m, self.L, save)
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
        _native.ngpref_fwd(tree.flat, self.spec(), x, d, rays, ts, n, T, save, ws, dens, rgb, aux_mse, aux_neg)
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
                return d_aux[name].reshape(-1).contiguous()
            if zeros is None:
                zeros = torch.zeros(m, device=d_flat.device)
            return zeros
        _native.ngpref_bwd(ctx["tree"].flat, self.spec(), None, None, ctx["rays"], ctx["ts"], ctx["n"], ctx["T"],
                           ctx["ws"], d_dens.reshape(-1).contiguous(), d_rgb.reshape(-1, 3).contiguous(),
                           aux("normal_mse"), aux("neg_normal"), d_flat)
