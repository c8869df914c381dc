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
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
