"""torch-CPU restatement of the reference field models (differentiable).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (no JAX here).

Follows learn_nerf/model.py:42-77, learn_nerf/instant_ngp.py:33-54,92-224 and
learn_nerf/ref_nerf.py:34-143,146-195,314-326.  Parameters are plain nested
dicts of tensors with Flax's auto-generated names (``Dense_i/kernel[in,out]``,
``Dense_i/bias[out]``, ``MultiresHashTableEncoding_0/HashTableEncoding_l/table``),
so the same tree drives the oracle and the CUDA path.  Works in fp32 or fp64.
"""
import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

Params = Dict[str, Dict[str, torch.Tensor]]


# --------------------------------------------------------------------------- init
def _trunc_normal(gen: torch.Generator, shape, std: float) -> torch.Tensor:
    """Unit normal truncated to [-2, 2], times ``std`` (the draw of flax's lecun_normal; the
    PRNG stream is torch's, not JAX's: the benchmark only needs "that architecture, random init")."""
    out = torch.empty(shape, dtype=torch.float64)
    flat = out.view(-1)
    filled = 0
    while filled < flat.numel():
        cand = torch.randn(flat.numel() - filled + 64, generator=gen, dtype=torch.float64)
        cand = cand[cand.abs() <= 2.0][: flat.numel() - filled]
        flat[filled: filled + cand.numel()] = cand
        filled += cand.numel()
    return (out * std).float()


def dense_init(gen: torch.Generator, fan_in: int, fan_out: int) -> Dict[str, torch.Tensor]:
    """flax nn.Dense defaults: kernel lecun_normal = variance_scaling(1, "fan_in", "truncated_normal"),
    i.e. stddev sqrt(1 / fan_in) / 0.87962566103423978 on a +-2 truncated unit normal; zero bias."""
    return dict(kernel=_trunc_normal(gen, (fan_in, fan_out), math.sqrt(1.0 / fan_in) / 0.87962566103423978),
                bias=torch.zeros(fan_out))


def dense(p: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """flax nn.Dense: x @ kernel + bias."""
    return x @ p["kernel"] + p["bias"]


# --------------------------------------------------------------------------- model.py
def sinusoidal_emb(coords: torch.Tensor, freqs: int) -> torch.Tensor:
    """model.py:65-77.  Per coordinate: [sin 2^0..2^{F-1}, cos 2^0..2^{F-1}]."""
    coeffs = 2.0 ** torch.arange(freqs, dtype=coords.dtype)
    inputs = coords[..., None] * coeffs
    combined = torch.cat([torch.sin(inputs), torch.cos(inputs)], dim=-1)
    return combined.reshape(combined.shape[:-2] + (-1,))


class NeRFModel:
    """model.py:30-62."""

    def __init__(self, input_layers=5, mid_layers=4, hidden_dim=256, color_layer_dim=128,
                 x_freqs=10, d_freqs=4):
        self.input_layers, self.mid_layers = input_layers, mid_layers
        self.hidden_dim, self.color_layer_dim = hidden_dim, color_layer_dim
        self.x_freqs, self.d_freqs = x_freqs, d_freqs

    def layer_dims(self) -> List[Tuple[int, int]]:
        xe, de, h = 6 * self.x_freqs, 6 * self.d_freqs, self.hidden_dim
        dims = [(xe, h)] + [(h, h)] * (self.input_layers - 1)
        dims += [(h + xe, h)] + [(h, h)] * (self.mid_layers - 1)
        dims += [(h, 1), (h + de, self.color_layer_dim), (self.color_layer_dim, 3)]
        return dims

    def init(self, gen: torch.Generator) -> Params:
        return {f"Dense_{i}": dense_init(gen, a, b) for i, (a, b) in enumerate(self.layer_dims())}

    def apply(self, params: Params, x: torch.Tensor, d: torch.Tensor):
        x_emb = sinusoidal_emb(x, self.x_freqs)  # :46
        d_emb = sinusoidal_emb(d, self.d_freqs)  # :47
        li = 0
        z = x_emb
        for _ in range(self.input_layers):  # :50-51
            z = torch.relu(dense(params[f"Dense_{li}"], z)); li += 1
        z = torch.cat([z, x_emb], dim=-1)  # :52
        for i in range(self.mid_layers):  # :53-56
            if i > 0:
                z = torch.relu(z)
            z = dense(params[f"Dense_{li}"], z); li += 1
        density = torch.nn.functional.softplus(dense(params[f"Dense_{li}"], z)); li += 1  # :57
        z = torch.cat([z, d_emb], dim=-1)  # :58
        z = torch.relu(dense(params[f"Dense_{li}"], z)); li += 1  # :59
        rgb = torch.tanh(dense(params[f"Dense_{li}"], z))  # :60
        return density, rgb, {}


# --------------------------------------------------------------------------- instant_ngp.py
def hash_table_lookup_indices(coords: torch.Tensor, table_size: int) -> torch.Tensor:
    """instant_ngp.py:211-224: (x ^ 19349663*y ^ 83492791*z) % T in uint32 wrap-around.
    coords int64 holding uint32 values."""
    m = 0xFFFFFFFF
    x, y, z = coords[:, 0] & m, coords[:, 1] & m, coords[:, 2] & m
    return (x ^ ((19_349_663 * y) & m) ^ ((83_492_791 * z) & m)) % table_size


def hash_level_rows(table_size: int, grid_size: int) -> int:
    """instant_ngp.py:178,191: hashed table if G^3 > T else a dense G^3 table."""
    return table_size if grid_size ** 3 > table_size else grid_size ** 3


def hash_table_encoding(table: torch.Tensor, x: torch.Tensor, table_size: int, grid_size: int,
                        bbox_min: torch.Tensor, bbox_max: torch.Tensor, smooth: bool = False,
                        return_debug: bool = False):
    """HashTableEncoding.__call__, instant_ngp.py:134-208."""
    frac = torch.clamp((x - bbox_min) / (bbox_max - bbox_min), 0, 1)  # :138-140
    if smooth:
        fi = 0.5 + (grid_size - 2) * frac  # :144
    else:
        fi = (grid_size - 1) * frac  # :146
    floored = torch.clamp(torch.floor(fi), max=grid_size - 2)  # :147-150
    cf = fi - floored  # :152
    if smooth:
        cf = (cf ** 2) * (3 - 2 * cf)  # :154
    base = floored.to(torch.int64)  # :156
    feats = None
    dbg = []
    for xo in (0, 1):  # :160-176
        for yo in (0, 1):
            for zo in (0, 1):
                off = torch.tensor([xo, yo, zo])
                cc = base + off
                w = torch.prod(1 + (2 * cf - 1) * off.to(cf.dtype) - cf, dim=-1, keepdim=True)
                if grid_size ** 3 > table_size:  # :178-190
                    idx = hash_table_lookup_indices(cc, table.shape[0])
                else:  # :191-204
                    idx = cc[:, 0] + grid_size * (cc[:, 1] + grid_size * cc[:, 2])
                term = w * table[idx]
                feats = term if feats is None else feats + term  # :206-208 (sum over 8)
                dbg.append((idx, w))
    if return_debug:
        return feats, dbg
    return feats


class InstantNGPModel:
    """instant_ngp.py:16-54 (+ MultiresHashTableEncoding :92-118)."""

    def __init__(self, table_sizes: Sequence[int], grid_sizes: Sequence[int], bbox_min, bbox_max,
                 table_feature_dim=2, table_smooth=False, d_freqs=4, hidden_dim=64,
                 density_dim=16, density_layers=1, color_layers=2):
        self.table_sizes, self.grid_sizes = list(table_sizes), list(grid_sizes)
        self.bbox_min = torch.as_tensor(bbox_min, dtype=torch.float32)
        self.bbox_max = torch.as_tensor(bbox_max, dtype=torch.float32)
        self.F, self.smooth, self.d_freqs = table_feature_dim, table_smooth, d_freqs
        self.hidden_dim, self.density_dim = hidden_dim, density_dim
        self.density_layers, self.color_layers = density_layers, color_layers

    def layer_dims(self):
        enc = self.F * len(self.grid_sizes)
        dims, cur = [], enc
        for _ in range(self.density_layers):
            dims.append((cur, self.hidden_dim)); cur = self.hidden_dim
        dims.append((cur, self.density_dim))
        cur = 6 * self.d_freqs + self.density_dim
        for _ in range(self.color_layers):
            dims.append((cur, self.hidden_dim)); cur = self.hidden_dim
        dims.append((cur, 3))
        return dims

    def init(self, gen: torch.Generator) -> Params:
        p = {f"Dense_{i}": dense_init(gen, a, b) for i, (a, b) in enumerate(self.layer_dims())}
        enc = {}
        for l, (t, g) in enumerate(zip(self.table_sizes, self.grid_sizes)):
            rows = hash_level_rows(t, g)
            u = torch.rand((rows, self.F), generator=gen)
            enc[f"HashTableEncoding_{l}"] = dict(table=(1e-4 * (u * 2 - 1)).float())  # :181-186
        p["MultiresHashTableEncoding_0"] = enc
        return p

    def encode(self, params: Params, x: torch.Tensor) -> torch.Tensor:
        enc = params["MultiresHashTableEncoding_0"]
        outs = []
        for l, (t, g) in enumerate(zip(self.table_sizes, self.grid_sizes)):
            table = enc[f"HashTableEncoding_{l}"]["table"]
            outs.append(hash_table_encoding(table, x, t, g, self.bbox_min.to(x.dtype),
                                            self.bbox_max.to(x.dtype), self.smooth))
        return torch.cat(outs, dim=1)

    def apply(self, params: Params, x: torch.Tensor, d: torch.Tensor):
        d_emb = sinusoidal_emb(d, self.d_freqs)  # :37
        out = self.encode(params, x)  # :38-45
        li = 0
        for _ in range(self.density_layers):  # :46-47
            out = torch.relu(dense(params[f"Dense_{li}"], out)); li += 1
        out = dense(params[f"Dense_{li}"], out); li += 1  # :48
        density = torch.exp(out[:, :1])  # :49
        out = torch.cat([d_emb, out], dim=1)  # :50
        for _ in range(self.color_layers):  # :51-52
            out = torch.relu(dense(params[f"Dense_{li}"], out)); li += 1
        color = torch.tanh(dense(params[f"Dense_{li}"], out))  # :53
        return density, color, {}


# --------------------------------------------------------------------------- ref_nerf.py
# Real spherical harmonics, degree <= 4 (16 terms), as polynomials in (x,y,z);
# ref_nerf.py:174-195 (constants are the standard real-SH normalisations).
_SH_C0 = 0.28209479177387814
_SH_C1 = 0.48860251190291987
_SH_C2 = (1.0925484305920792, 0.94617469575755997, 0.31539156525251999, 0.54627421529603959)
_SH_C3 = (0.59004358992664352, 2.8906114426405538, 0.45704579946446572, 0.3731763325901154,
          1.4453057213202769)
HARMONIC_COUNTS = [1, 3, 5, 7]


def spherical_harmonic(sh_degree: int, coords: torch.Tensor) -> torch.Tensor:
    assert 1 <= sh_degree <= 4, "oracle restates degrees 1..4 (config uses 4)"
    x, y, z = coords[:, 0], coords[:, 1], coords[:, 2]
    x2, y2, z2 = x * x, y * y, z * z
    out = [torch.full_like(x, _SH_C0)]
    if sh_degree > 1:
        out += [-_SH_C1 * y, _SH_C1 * z, -_SH_C1 * x]
    if sh_degree > 2:
        a, b, c, e = _SH_C2
        out += [a * (x * y), -a * (y * z), b * z2 - c, -a * (x * z), e * x2 - e * y2]
    if sh_degree > 3:
        a, b, c, e, f = _SH_C3
        out += [a * y * (-3.0 * x2 + y2), b * (x * y) * z, c * y * (1.0 - 5.0 * z2),
                e * z * (5.0 * z2 - 3.0), c * x * (1.0 - 5.0 * z2), f * z * (x2 - y2),
                a * x * (-x2 + 3.0 * y2)]
    return torch.stack(out, dim=1)


def integrated_directional_encoding(sh_degree: int, coords: torch.Tensor, roughness: torch.Tensor):
    """ref_nerf.py:121-143."""
    assert roughness.dim() == 2 and roughness.shape[1] == 1
    levels = torch.tensor([i for i, c in enumerate(HARMONIC_COUNTS[:sh_degree]) for _ in range(c)],
                          dtype=roughness.dtype)
    attenuation = torch.exp(-roughness * (levels * (levels + 1)) / 2)
    return spherical_harmonic(sh_degree, coords) * attenuation


def linear_rgb_to_srgb(c: torch.Tensor) -> torch.Tensor:
    """ref_nerf.py:110-118."""
    safe = torch.clamp(c, min=1e-5)
    return torch.where(c <= 0.0031308, 12.92 * c, 1.055 * (safe ** (1 / 2.4)) - 0.055)


def _safe_normalize(v: torch.Tensor, eps=1e-10) -> torch.Tensor:  # ref_nerf.py:314-317
    return v / torch.sqrt(torch.sum(v ** 2, dim=-1, keepdim=True) + eps)


def _leaky_clip(x: torch.Tensor) -> torch.Tensor:  # ref_nerf.py:320-326
    return x + (torch.clamp(x, 0, 1) - x).detach()


class RefNERFModel:
    """ref_nerf.py:80-107 on RefNERFBase.__call__ (:34-77)."""

    def __init__(self, sh_degree=4, input_layers=5, mid_layers=4, hidden_dim=256,
                 color_layer_dim=128, x_freqs=10, d_freqs=4):
        self.sh_degree = sh_degree
        self.input_layers, self.mid_layers = input_layers, mid_layers
        self.hidden_dim, self.color_layer_dim = hidden_dim, color_layer_dim
        self.x_freqs, self.d_freqs = x_freqs, d_freqs

    def layer_dims(self):
        xe, h = 6 * self.x_freqs, self.hidden_dim
        dims = [(xe, h)] + [(h, h)] * (self.input_layers - 1)
        dims += [(h + xe, h)] + [(h, h)] * (self.mid_layers - 1)
        enc = sum(HARMONIC_COUNTS[: self.sh_degree])
        dims += [(h + enc + 1, self.color_layer_dim), (self.color_layer_dim, 3)]
        return dims

    def init(self, gen: torch.Generator) -> Params:
        return {f"Dense_{i}": dense_init(gen, a, b) for i, (a, b) in enumerate(self.layer_dims())}

    def spatial_block(self, params: Params, x: torch.Tensor) -> torch.Tensor:
        x_emb = sinusoidal_emb(x, self.x_freqs)
        z, li = x_emb, 0
        for _ in range(self.input_layers):
            z = torch.relu(dense(params[f"Dense_{li}"], z)); li += 1
        z = torch.cat([z, x_emb], dim=-1)
        for i in range(self.mid_layers):
            if i > 0:
                z = torch.relu(z)
            z = dense(params[f"Dense_{li}"], z); li += 1
        return z

    def directional_block(self, params: Params, x: torch.Tensor) -> torch.Tensor:
        li = self.input_layers + self.mid_layers
        z = torch.relu(dense(params[f"Dense_{li}"], x))
        return dense(params[f"Dense_{li + 1}"], z)

    def apply(self, params: Params, x: torch.Tensor, d: torch.Tensor, create_graph=None):
        if create_graph is None:
            create_graph = torch.is_grad_enabled()
        with torch.enable_grad():
            xg = x.detach().requires_grad_(True) if not x.requires_grad else x
            spatial_out = self.spatial_block(params, xg)
            (real_normal,) = torch.autograd.grad(-spatial_out[:, 0].sum(), xg,
                                                 create_graph=create_graph, retain_graph=True)  # :38-42
        if not create_graph:
            spatial_out, real_normal = spatial_out.detach(), real_normal.detach()
        real_normal = _safe_normalize(real_normal)  # :43
        density, diffuse, spectral, roughness, normal = (
            spatial_out[:, 0:1], spatial_out[:, 1:4], spatial_out[:, 4:5],
            spatial_out[:, 5:6], spatial_out[:, 6:9])  # :45-47
        density = torch.exp(density)  # :48
        diffuse = torch.sigmoid(diffuse - math.log(3))  # :52
        spectral = torch.sigmoid(spectral)  # :54
        roughness = torch.nn.functional.softplus(roughness)  # :55
        normal = _safe_normalize(normal)  # :56
        reflection = d - 2 * normal * torch.sum(d * normal, dim=-1, keepdim=True)  # :58
        enc = integrated_directional_encoding(self.sh_degree, reflection, roughness)  # :59-61
        normal_dot = torch.sum(-d * normal, dim=-1, keepdim=True)  # :62
        dir_input = torch.cat([spatial_out, enc, normal_dot], dim=1)  # :63
        spectral_color = torch.sigmoid(self.directional_block(params, dir_input))  # :64-65
        full = linear_rgb_to_srgb(_leaky_clip(spectral_color * spectral + diffuse)) * 2 - 1  # :67-71
        aux = dict(normal_mse=torch.sum((normal - real_normal) ** 2, dim=-1),
                   neg_normal=torch.clamp(torch.sum(normal * d, dim=-1), min=0.0) ** 2)  # :72-75
        return density, full, aux


class InstantNGPRefNERFModel(RefNERFModel):
    """instant_ngp.py:57-89: RefNERFBase.__call__ (RefNERFModel.apply above) with a smooth hash-grid
    spatial block (:69-82) and a 2 x 64 directional block (:84-89).  Flax creation order inside
    __call__: MultiresHashTableEncoding_0, Dense_0, Dense_1 (spatial), Dense_2..4 (directional)."""

    def __init__(self, table_sizes: Sequence[int], grid_sizes: Sequence[int], bbox_min, bbox_max, sh_degree=4,
                 table_feature_dim=2, d_freqs=4, hidden_dim=64, density_dim=16, density_layers=1,
                 color_layers=2):
        self.sh_degree = sh_degree
        self.table_sizes, self.grid_sizes = list(table_sizes), list(grid_sizes)
        self.bbox_min = torch.as_tensor(bbox_min, dtype=torch.float32)
        self.bbox_max = torch.as_tensor(bbox_max, dtype=torch.float32)
        self.F, self.d_freqs = table_feature_dim, d_freqs
        self.hidden_dim, self.density_dim = hidden_dim, density_dim
        self.density_layers, self.color_layers = density_layers, color_layers

    def layer_dims(self):
        dims, cur = [], self.F * len(self.grid_sizes)
        for _ in range(self.density_layers):
            dims.append((cur, self.hidden_dim)); cur = self.hidden_dim
        dims.append((cur, self.density_dim))
        cur = self.density_dim + sum(HARMONIC_COUNTS[: self.sh_degree]) + 1  # ref_nerf.py:63
        for _ in range(self.color_layers):
            dims.append((cur, self.hidden_dim)); cur = self.hidden_dim
        dims.append((cur, 3))
        return dims

    def init(self, gen: torch.Generator) -> Params:
        p = {f"Dense_{i}": dense_init(gen, a, b) for i, (a, b) in enumerate(self.layer_dims())}
        enc = {}
        for l, (t, g) in enumerate(zip(self.table_sizes, self.grid_sizes)):
            u = torch.rand((hash_level_rows(t, g), self.F), generator=gen)
            enc[f"HashTableEncoding_{l}"] = dict(table=(1e-4 * (u * 2 - 1)).float())
        p["MultiresHashTableEncoding_0"] = enc
        return p

    def spatial_block(self, params: Params, x: torch.Tensor) -> torch.Tensor:  # :69-82
        enc = params["MultiresHashTableEncoding_0"]
        outs = []
        for l, (t, g) in enumerate(zip(self.table_sizes, self.grid_sizes)):
            outs.append(hash_table_encoding(enc[f"HashTableEncoding_{l}"]["table"], x, t, g,
                                            self.bbox_min.to(x.dtype), self.bbox_max.to(x.dtype), smooth=True))
        z, li = torch.cat(outs, dim=1), 0
        for _ in range(self.density_layers):
            z = torch.relu(dense(params[f"Dense_{li}"], z)); li += 1
        return dense(params[f"Dense_{li}"], z)

    def directional_block(self, params: Params, x: torch.Tensor) -> torch.Tensor:  # :84-89
        li = self.density_layers + 1
        for _ in range(self.color_layers):
            x = torch.relu(dense(params[f"Dense_{li}"], x)); li += 1
        return dense(params[f"Dense_{li}"], x)


# --------------------------------------------------------------------------- helpers
def tree_map(fn, tree):
    if isinstance(tree, dict):
        return {k: tree_map(fn, v) for k, v in tree.items()}
    return fn(tree)


def tree_leaves(tree, prefix=""):
    """Deterministic (sorted-key, like jax pytrees) list of (path, leaf)."""
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree.keys()):
            out += tree_leaves(tree[k], f"{prefix}/{k}" if prefix else k)
        return out
    return [(prefix, tree)]


def as_numpy_model_fn(model, params):
    """Adapter: numpy in / numpy out, fp32, no grad (for oracle.render_np)."""
    def fn(x: np.ndarray, d: np.ndarray):
        with torch.no_grad():
            de, rgb, aux = model.apply(params, torch.from_numpy(np.ascontiguousarray(x)),
                                       torch.from_numpy(np.ascontiguousarray(d)))
        return de.numpy(), rgb.numpy(), {k: v.numpy() for k, v in aux.items()}
    return fn
