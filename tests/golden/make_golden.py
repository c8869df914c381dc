"""Generates the committed fixtures in tests/golden/ FROM THE ORACLE ITSELF.

The reference (JAX) cannot run in this image, so these are regression fixtures of the
oracle restatement, not reference outputs ("parity unpinned", see oracle/__init__.py).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import BBOX_MAX, BBOX_MIN, F, make_rays, make_uniforms  # noqa: E402
from oracle import models_torch as M  # noqa: E402
from oracle import render_np  # noqa: E402
from oracle import train_torch as T  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    n = 24
    batch = make_rays(n, seed=11, miss_frac=0.25)
    uc, uf = make_uniforms(n, 64, 12), make_uniforms(n, 128, 13)
    nerf = M.NeRFModel()
    params = T.init_params(nerf, nerf, 2)
    r = render_np.NeRFRenderer(M.as_numpy_model_fn(nerf, params["coarse"]),
                               M.as_numpy_model_fn(nerf, params["fine"]),
                               params["background"].numpy(), BBOX_MIN, BBOX_MAX, 64, 128)
    smp = {}
    out = r.render_rays(uc, uf, batch[:, :2], smp)
    np.savez_compressed(
        os.path.join(HERE, "nerf_render_small.npz"), batch=batch, coarse_ts=smp["coarse"].ts,
        mask=smp["coarse"].mask, t_min=smp["coarse"].t_min, t_max=smp["coarse"].t_max,
        fine_ts=smp["fine"].ts, coarse_outputs=out["coarse"]["outputs"],
        fine_outputs=out["fine"]["outputs"], fine_alphas=out["fine"]["alphas"],
        fine_coords=out["fine"]["coords"], coarse_densities=out["coarse"]["densities"])

    # fine sampling from stored inputs, incl. hard cases: zero density, one huge density
    # (flat CDF bins -> dx == 0 branch), masked rays with t in [0, 1e-3]
    rs = np.random.RandomState(21)
    n2 = 48
    rays = make_rays(n2, seed=22, miss_frac=0.2, with_targets=False)
    t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
    cs = render_np.RaySamples.stratified_sampling(t_min, t_max, mask, 64, make_uniforms(n2, 64, 23))
    dens = rs.gamma(0.5, 4.0, (n2, 64)).astype(F)
    dens[0] = 0.0
    dens[1] = 0.0
    dens[1, 10] = 1e6
    dens[2, :5] = 1e4
    dens[3] = 1e-12
    u = make_uniforms(n2, 128, 24)
    u[4, 127] = np.float32(1.0 - 2.0 ** -23)  # largest representable uniform
    fs, idx = cs.fine_sampling(128, u, dens, return_indices=True)
    new_only, _ = cs.fine_sampling(128, u, dens, combine=False, return_indices=True)
    np.savez_compressed(os.path.join(HERE, "fine_sampling_small.npz"), ts=cs.ts, t_min=t_min,
                        t_max=t_max, mask=mask, densities=dens, u=u, fine_ts=fs.ts, idx=idx,
                        new_ts=new_only.ts, rays=rays)
    models_fixture()
    print("wrote fixtures to", HERE)


def models_case():
    """Inputs and seeded oracle models shared by make_golden.py and the tests."""
    import torch
    rs = np.random.RandomState(31)
    x = rs.uniform(-1.1, 1.1, (64, 3)).astype(F)
    d = rs.randn(64, 3).astype(F)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ngp = M.InstantNGPModel([2 ** 18] * 16, [2 ** (4 + i // 2) for i in range(16)], BBOX_MIN, BBOX_MAX)
    p_ngp = ngp.init(torch.Generator().manual_seed(41))
    for leaf in p_ngp["MultiresHashTableEncoding_0"].values():
        leaf["table"] *= 1e4  # O(1) tables so that the encoding matters
    ref = M.RefNERFModel()
    p_ref = ref.init(torch.Generator().manual_seed(42))
    for i, leaf in enumerate(p_ref.values()):
        leaf["bias"] = torch.from_numpy((0.1 * np.random.RandomState(50 + i).randn(*leaf["bias"].shape)).astype(F))
    cam = dict(camera_direction=(0.1, -0.2, -0.97), camera_origin=(0.5, 1.0, 4.0), x_axis=(0.99, 0.05, 0.09),
               y_axis=(0.04, -0.98, 0.2), x_fov=1.0471975511965976, y_fov=0.7853981633974483)
    return x, d, ngp, p_ngp, ref, p_ref, cam


def models_fixture():
    import torch
    from oracle import prng_np
    x, d, ngp, p_ngp, ref, p_ref, cam = models_case()
    with torch.no_grad():
        nd, nrgb, _ = ngp.apply(p_ngp, torch.from_numpy(x), torch.from_numpy(d))
        nenc = ngp.encode(p_ngp, torch.from_numpy(x))
    rd, rrgb, raux = ref.apply(p_ref, torch.from_numpy(x), torch.from_numpy(d), create_graph=False)
    rays = render_np.bare_rays(width=7, height=5, **cam)
    key = prng_np.split(prng_np.prng_key(123))[1]
    np.savez_compressed(
        os.path.join(HERE, "models_small.npz"), x=x, d=d, ngp_enc=nenc.numpy(), ngp_dens=nd.numpy(),
        ngp_rgb=nrgb.numpy(), ref_dens=rd.numpy(), ref_rgb=rrgb.numpy(),
        ref_normal_mse=raux["normal_mse"].numpy(), ref_neg_normal=raux["neg_normal"].numpy(), rays_7x5=rays,
        key=key, uniforms_5x7=prng_np.uniform(key, (5, 7)))


def ngpref_case():
    """Seeded InstantNGPRefNERFModel (instant_ngp.py:57-89) and a density-penalty case (train.py:153-184)."""
    import torch
    rs = np.random.RandomState(61)
    x = rs.uniform(-1.05, 1.05, (48, 3)).astype(F)
    d = rs.randn(48, 3).astype(F)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = M.InstantNGPRefNERFModel([2 ** 14] * 16, [2 ** (4 + i // 2) for i in range(16)], BBOX_MIN, BBOX_MAX)
    p = o.init(torch.Generator().manual_seed(62))
    for leaf in p["MultiresHashTableEncoding_0"].values():
        leaf["table"] *= 0.5e4
    for i in range(5):
        leaf = p[f"Dense_{i}"]
        leaf["bias"] = torch.from_numpy((0.1 * np.random.RandomState(70 + i).randn(*leaf["bias"].shape)).astype(F))
    return x, d, o, p


def ngpref_fixture():
    """Separate file so that the older fixtures stay byte-identical: python -c
    'import make_golden as m; m.ngpref_fixture()' from tests/golden/."""
    import torch
    x, d, o, p = ngpref_case()
    de, rgb, aux = o.apply(p, torch.from_numpy(x), torch.from_numpy(d), create_graph=False)
    nerf = M.NeRFModel()
    params = T.init_params(nerf, nerf, 7)
    batch, uc, uf = make_rays(8, seed=63), make_uniforms(8, 64, 64), make_uniforms(8, 128, 65)
    total, ld, _ = T.losses(nerf, nerf, params, BBOX_MIN, BBOX_MAX, batch, uc, uf, 64, 128, density_penalty=0.1,
                            density_points=(x, d))
    np.savez_compressed(os.path.join(HERE, "ngpref_small.npz"), x=x, d=d, dens=de.numpy(), rgb=rgb.numpy(),
                        normal_mse=aux["normal_mse"].numpy(), neg_normal=aux["neg_normal"].numpy(),
                        penalty_total=np.float32(total), penalty_fine=np.float32(ld["fine_density"]),
                        penalty_coarse=np.float32(ld["coarse_density"]))


if __name__ == "__main__":
    main()
