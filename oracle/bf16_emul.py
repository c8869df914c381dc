"""bf16-emulating oracle of the NeRF MLP forward/backward (torch CPU, fp64 accumulation).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED (no JAX here).

The tensor-core path (learn-nerf_b200/csrc/mlp_tc*.cu) rounds to bf16 at fixed points:
weights of the tensor layers, the positional encodings, every hidden activation tile
and every backward gradient tile; accumulation is fp32.  Against the exact fp64 model
(oracle.models_torch) this only supports a loose tolerance, and ReLU masks taken from
bf16-rounded activations make the gradient that of a slightly different function.  This
module restates model.py:42-62 and its hand-derived backward with the SAME rounding
points (fp64 accumulation instead of fp32), so kernel bugs can be told apart from
precision effects: the CUDA path must match it tightly.
"""
from typing import Dict

import torch

from .models_torch import sinusoidal_emb


def bf(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 (round-to-nearest-even) and return as float64."""
    return t.float().bfloat16().double()


def forward_backward(params: Dict[str, Dict[str, torch.Tensor]], x: torch.Tensor, d: torch.Tensor,
                     d_dens: torch.Tensor = None, d_rgb: torch.Tensor = None):
    """x, d [M,3] fp32 points/directions.  Returns (dens[M], rgb[M,3]) and, if upstream
    gradients are given, the parameter-gradient tree (fp64)."""
    W = {i: params[f"Dense_{i}"]["kernel"].double() for i in range(12)}
    B = {i: params[f"Dense_{i}"]["bias"].double() for i in range(12)}
    Wb = {i: bf(W[i]) for i in range(12)}
    xe = bf(sinusoidal_emb(x.float(), 10))
    de = bf(sinusoidal_emb(d.float(), 4))
    h = {}
    a = xe
    for l in range(5):  # model.py:50-51
        a = bf(torch.relu(a @ Wb[l] + B[l]))
        h[l] = a
    a = bf(torch.relu(torch.cat([a, xe], 1) @ Wb[5] + B[5]))  # :52-56 (Dense_5..7 feed a ReLU)
    h[5] = a
    for l in (6, 7):
        a = bf(torch.relu(a @ Wb[l] + B[l]))
        h[l] = a
    z8 = bf(h[7] @ Wb[8] + B[8])  # raw (model.py:56-58)
    pre_d = z8 @ Wb[9][:, 0] + B[9][0]
    dens = torch.nn.functional.softplus(pre_d)  # :57
    c = torch.relu(torch.cat([z8, de], 1) @ Wb[10] + B[10])  # :59, kept fp32 in the kernel
    rgb = torch.tanh(c @ W[11] + B[11])  # :60, fp32 head weights
    if d_dens is None:
        return dens, rgb, None

    g = {}
    spre = d_dens.double() * (1 - torch.exp(-dens))
    dpre = d_rgb.double() * (1 - rgb * rgb)
    cb = bf(c)
    dc = bf((dpre @ W[11].T) * (c > 0))
    grads = {i: {} for i in range(12)}
    grads[11]["kernel"], grads[11]["bias"] = cb.T @ dpre, dpre.sum(0)
    grads[9]["kernel"], grads[9]["bias"] = (z8.T @ spre)[:, None], spre.sum()[None]
    grads[10]["kernel"], grads[10]["bias"] = torch.cat([z8, de], 1).T @ dc, dc.sum(0)
    g[8] = bf(dc @ Wb[10][:256].T + spre[:, None] * W[9][:, 0][None])
    for l in range(8, 0, -1):
        inp = torch.cat([h[4], xe], 1) if l == 5 else h[l - 1]
        grads[l]["kernel"], grads[l]["bias"] = inp.T @ g[l], g[l].sum(0)
        g[l - 1] = bf((g[l] @ Wb[l][:256].T) * (h[l - 1] > 0))
    grads[0]["kernel"], grads[0]["bias"] = xe.T @ g[0], g[0].sum(0)
    tree = {f"Dense_{i}": grads[i] for i in range(12)}
    return dens, rgb, tree
