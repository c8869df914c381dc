"""CPU tests of the oracle itself: self-derived known-answer tests (the
reference holds no golden vectors for this path -> parity unpinned) + regression
against the committed fixtures in tests/golden/."""
import os

import numpy as np
import pytest
import torch

from oracle import models_torch as M
from oracle import render_np
from oracle import train_torch as T
from oracle.expf import expf

from helpers import BBOX_MAX, BBOX_MIN, F, make_rays, make_uniforms

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_expf_accuracy():
    x = np.concatenate([np.linspace(-87, 0, 200001), -np.logspace(-30, 1.9, 20000)]).astype(F)
    y = expf(x)
    ref = np.exp(x.astype(np.float64))
    ulp = np.spacing(ref.astype(F)).astype(np.float64)
    assert np.max(np.abs(y.astype(np.float64) - ref) / ulp) <= 1.0
    assert expf(F(0.0)) == F(1.0)
    assert expf(F(-100.0)) == F(0.0)
    assert np.isinf(expf(F(100.0)))


def test_ray_t_range_kat():
    rays = np.array([[[0, 0, -3], [0, 0, 1]], [[0, 5, -3], [0, 0, 1]]], F)
    t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
    assert mask.tolist() == [True, False]
    np.testing.assert_array_equal(t_min, np.array([2, 0], F))
    np.testing.assert_array_equal(t_max, np.array([4, 1e-3], F))


def test_termination_probs_sum_to_one():
    rays = make_rays(32, with_targets=False)
    t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
    s = render_np.RaySamples.stratified_sampling(t_min, t_max, mask, 64, make_uniforms(32, 64))
    dens = np.random.RandomState(3).gamma(1.0, 2.0, (32, 64)).astype(F)
    p = s.termination_probs(dens)
    assert p.shape == (32, 65)
    np.testing.assert_allclose(p.sum(1), 1.0, atol=2e-6)


def test_zero_density_renders_background_and_uniform_cdf():
    n = 16
    rays = make_rays(n, with_targets=False)
    t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
    s = render_np.RaySamples.stratified_sampling(t_min, t_max, mask, 64, make_uniforms(n, 64))
    bg = np.array([0.25, -0.5, 1.0], F)
    zeros = np.zeros((n, 64), F)
    out = s.render_rays(zeros, np.ones((n, 64, 3), F), bg)
    np.testing.assert_array_equal(out, np.tile(bg, (n, 1)))
    np.testing.assert_array_equal(s.render_alpha(zeros), np.zeros((n, 1), F))
    fs, idx = s.fine_sampling(128, make_uniforms(n, 128, 5), zeros, combine=False, return_indices=True)
    # uniform CDF over [t_min, t_max]: new_ts ~ t_min + u' * (t_max - t_min)
    u = render_np.RaySamples.stratified_sampling(np.zeros(n, F), np.ones(n, F), mask, 128,
                                                 make_uniforms(n, 128, 5)).ts
    expect = t_min[:, None] + u * (t_max - t_min)[:, None]
    np.testing.assert_allclose(fs.ts, expect, atol=2e-2)  # bins are mid-point based, not exact
    assert idx.min() >= 1 and idx.max() <= 64


def test_opaque_first_sample_returns_its_colour():
    rays = make_rays(4, with_targets=False)
    t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
    s = render_np.RaySamples.stratified_sampling(t_min, t_max, mask, 8, make_uniforms(4, 8))
    dens = np.zeros((4, 8), F)
    dens[:, 0] = 1e9
    rgb = np.random.RandomState(0).uniform(-1, 1, (4, 8, 3)).astype(F)
    out = s.render_rays(dens, rgb, np.zeros(3, F))
    np.testing.assert_allclose(out, rgb[:, 0], atol=1e-6)


def test_sinusoidal_emb_layout():
    e = M.sinusoidal_emb(torch.tensor([[0.1, 0.2, 0.3]]), 10)[0]
    np.testing.assert_allclose(e[:3].numpy(), np.sin([0.1, 0.2, 0.4]), atol=1e-7)
    np.testing.assert_allclose(float(e[10]), np.cos(0.1), atol=1e-7)
    np.testing.assert_allclose(float(e[20]), np.sin(0.2), atol=1e-7)


def test_param_counts():
    nerf = M.NeRFModel()
    p = T.init_params(nerf, nerf, 0)
    assert sum(t.numel() for _, t in M.tree_leaves(p)) == 1_187_851
    assert sum(a * b + b for a, b in nerf.layer_dims()) == 593_924
    ref = M.RefNERFModel()
    assert sum(a * b + b for a, b in ref.layer_dims()) == 592_771


def test_hash_kats():
    assert int(M.hash_table_lookup_indices(torch.tensor([[0, 0, 0]]), 2 ** 18)) == 0
    c = torch.tensor([[1, 2, 3]])
    expect = (1 ^ ((19_349_663 * 2) & 0xFFFFFFFF) ^ ((83_492_791 * 3) & 0xFFFFFFFF)) % (2 ** 18)
    assert int(M.hash_table_lookup_indices(c, 2 ** 18)) == expect
    assert M.hash_level_rows(2 ** 18, 64) == 64 ** 3  # 64^3 == 2^18 is NOT > table_size -> dense
    assert M.hash_level_rows(2 ** 18, 128) == 2 ** 18
    # at an exact grid vertex the encoding equals that row; weights sum to one
    g = 16
    table = torch.rand(g ** 3, 2)
    x = torch.tensor([[-1 + 2 * 3 / 15, -1 + 2 * 5 / 15, -1 + 2 * 7 / 15]], dtype=torch.float64)
    f, dbg = M.hash_table_encoding(table.double(), x, 2 ** 18, g, torch.tensor([-1.0] * 3).double(),
                                   torch.tensor([1.0] * 3).double(), return_debug=True)
    np.testing.assert_allclose(f[0].numpy(), table[3 + g * (5 + g * 7)].numpy(), atol=1e-6)
    np.testing.assert_allclose(sum(float(w) for _, w in dbg), 1.0, atol=1e-12)


def test_sh_and_ide_kats():
    v = torch.tensor([[0.0, 0.0, 1.0], [0.6, 0.0, 0.8]])
    sh = M.spherical_harmonic(4, v)
    assert sh.shape == (2, 16)
    np.testing.assert_allclose(sh[:, 0].numpy(), 0.28209479177387814, atol=1e-7)
    ide = M.integrated_directional_encoding(4, v, torch.zeros(2, 1))
    np.testing.assert_allclose(ide.numpy(), sh.numpy(), atol=0)
    a, b = torch.tensor([0.0031308]), torch.tensor([0.0031308 + 1e-7])
    assert abs(float(M.linear_rgb_to_srgb(a)) - float(M.linear_rgb_to_srgb(b))) < 1e-5


def test_composite_gradient_matches_finite_differences():
    """fp64 finite differences pin the autograd oracle used for K5/K6 checks."""
    from oracle.render_torch import composite
    torch.manual_seed(0)
    n, t = 3, 9
    ts = torch.sort(torch.rand(n, t, dtype=torch.float64) * 2 + 2, dim=1).values
    t_min, t_max = torch.full((n,), 2.0, dtype=torch.float64), torch.full((n,), 4.0, dtype=torch.float64)
    mask = torch.tensor([True, True, False])
    dens = torch.rand(n, t, dtype=torch.float64, requires_grad=True)
    rgb = torch.rand(n, t, 3, dtype=torch.float64, requires_grad=True)
    bg = torch.rand(3, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda d, c, b: composite(ts, t_min, t_max, mask, d, c, b),
                                    (dens, rgb, bg), eps=1e-6, atol=1e-6)


def test_adam_matches_torch_optim():
    torch.manual_seed(0)
    p = dict(a=torch.randn(5, 3), b=dict(c=torch.randn(7)))
    ref = [p["a"].clone().requires_grad_(True), p["b"]["c"].clone().requires_grad_(True)]
    opt = torch.optim.Adam(ref, lr=1e-2, betas=(0.9, 0.999), eps=1e-7)
    st = T.AdamState(p)
    for _ in range(3):
        g = dict(a=torch.randn(5, 3), b=dict(c=torch.randn(7)))
        ref[0].grad, ref[1].grad = g["a"].clone(), g["b"]["c"].clone()
        opt.step()
        p = T.adam_update(p, g, st, 1e-2, eps=1e-7)
    # torch adds eps to sqrt(v_hat) exactly as optax (eps_root = 0)
    np.testing.assert_allclose(p["a"].numpy(), ref[0].detach().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(p["b"]["c"].numpy(), ref[1].detach().numpy(), rtol=1e-5, atol=1e-7)


def _golden_case():
    n = 24
    batch = make_rays(n, seed=11, miss_frac=0.25)
    return batch, make_uniforms(n, 64, 12), make_uniforms(n, 128, 13)


def test_golden_render_fixture():
    """Regression: the committed fixture was produced by tests/golden/make_golden.py from this
    oracle (self-generated; it guards the oracle against drift, it does not pin the reference)."""
    path = os.path.join(GOLDEN, "nerf_render_small.npz")
    g = np.load(path)
    batch, uc, uf = _golden_case()
    np.testing.assert_array_equal(g["batch"], batch)
    nerf = M.NeRFModel()
    params = T.init_params(nerf, nerf, 2)
    r = render_np.NeRFRenderer(M.as_numpy_model_fn(nerf, params["coarse"]),
                               M.as_numpy_model_fn(nerf, params["fine"]),
                               params["background"].numpy(), BBOX_MIN, BBOX_MAX, 64, 128)
    smp = {}
    out = r.render_rays(uc, uf, batch[:, :2], smp)
    np.testing.assert_array_equal(smp["coarse"].ts, g["coarse_ts"])  # bit-exact stage
    np.testing.assert_array_equal(smp["coarse"].mask, g["mask"])
    # fine positions depend on MLP densities (BLAS summation order): tolerance, not bits
    np.testing.assert_allclose(smp["fine"].ts, g["fine_ts"], atol=1e-4)
    np.testing.assert_allclose(out["fine"]["outputs"], g["fine_outputs"], atol=1e-5)
    np.testing.assert_allclose(out["coarse"]["outputs"], g["coarse_outputs"], atol=1e-5)
    np.testing.assert_allclose(out["fine"]["alphas"], g["fine_alphas"], atol=1e-5)
    np.testing.assert_allclose(out["fine"]["coords"], g["fine_coords"], atol=1e-5)


def test_golden_fine_sampling_fixture_bit_exact():
    """Given stored (ts, densities, u) the fine sampler is pure fp32 elementwise work: bit-exact."""
    g = np.load(os.path.join(GOLDEN, "fine_sampling_small.npz"))
    s = render_np.RaySamples(t_min=g["t_min"], t_max=g["t_max"], mask=g["mask"], ts=g["ts"])
    out, idx = s.fine_sampling(128, g["u"], g["densities"], return_indices=True)
    np.testing.assert_array_equal(out.ts, g["fine_ts"])
    np.testing.assert_array_equal(idx, g["idx"])


def test_threefry_known_answers():
    """Random123 Threefry-2x32 vectors (also used by JAX's own tests) and JAX's documented outputs."""
    from oracle import prng_np as P
    h = lambda a: [int(x) for x in a]
    assert h(P.threefry_2x32([0, 0], [0, 0])) == [0x6B200159, 0x99BA4EFE]
    assert h(P.threefry_2x32([0xFFFFFFFF] * 2, [0xFFFFFFFF] * 2)) == [0x1CB996FC, 0xBB002BE7]
    assert h(P.threefry_2x32([0x13198A2E, 0x03707344], [0x243F6A88, 0x85A308D3])) == [0xC4923A9C, 0x483DF7A0]
    assert P.split(P.prng_key(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert abs(float(P.uniform(P.prng_key(0), ())) - 0.41845703) < 1e-8
    u = P.uniform(P.prng_key(3), (5, 7))
    assert u.dtype == np.float32 and (u >= 0).all() and (u < 1).all()
    assert np.all(u * 2 ** 23 == np.round(u * 2 ** 23))


def test_golden_models_fixture():
    """Regression of the NGP / Ref-NeRF / ray-generation / PRNG restatements against the committed
    fixture (generated by tests/golden/make_golden.py from this oracle; the GPU tests compare the
    CUDA path with the same file)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    from oracle import prng_np
    g = np.load(os.path.join(GOLDEN, "models_small.npz"))
    x, d, ngp, p_ngp, ref, p_ref, cam = mg.models_case()
    np.testing.assert_array_equal(x, g["x"])
    with torch.no_grad():
        nd, nrgb, _ = ngp.apply(p_ngp, torch.from_numpy(x), torch.from_numpy(d))
    np.testing.assert_allclose(nd.numpy(), g["ngp_dens"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(nrgb.numpy(), g["ngp_rgb"], atol=1e-5)
    rd, rrgb, raux = ref.apply(p_ref, torch.from_numpy(x), torch.from_numpy(d), create_graph=False)
    np.testing.assert_allclose(rd.numpy(), g["ref_dens"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(rrgb.numpy(), g["ref_rgb"], atol=1e-5)
    np.testing.assert_allclose(raux["neg_normal"].numpy(), g["ref_neg_normal"], atol=1e-5)
    np.testing.assert_array_equal(render_np.bare_rays(width=7, height=5, **cam), g["rays_7x5"])
    np.testing.assert_array_equal(prng_np.uniform(g["key"], (5, 7)), g["uniforms_5x7"])


def test_density_penalty_branch():
    """train.py:153-184 restated in oracle.train_torch.losses: with weight w the total grows by
    w * (mean density of the fine model + of the coarse model) at the given points, and the
    gradient of that term only reaches the density branch (the rgb head gets none)."""
    from oracle import models_torch as M
    from oracle import train_torch as T
    nerf = M.NeRFModel()
    params = T.init_params(nerf, nerf, 3)
    rs = np.random.RandomState(0)
    batch = make_rays(8, seed=2)
    uc, uf = make_uniforms(8, 64, 3), make_uniforms(8, 128, 4)
    coords = rs.uniform(-1, 1, (16, 3)).astype(np.float32)
    dirs = rs.randn(16, 3).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    t0, ld0, _ = T.losses(nerf, nerf, params, BBOX_MIN, BBOX_MAX, batch, uc, uf, 64, 128)
    t1, ld1, _ = T.losses(nerf, nerf, params, BBOX_MIN, BBOX_MAX, batch, uc, uf, 64, 128,
                          density_penalty=0.25, density_points=(coords, dirs))
    with torch.no_grad():
        df = nerf.apply(params["fine"], torch.from_numpy(coords), torch.from_numpy(dirs))[0].mean()
        dc = nerf.apply(params["coarse"], torch.from_numpy(coords), torch.from_numpy(dirs))[0].mean()
    assert list(ld1)[-2:] == ["fine_density", "coarse_density"]  # fine first (:154)
    np.testing.assert_allclose(float(ld1["fine_density"]), float(df), rtol=1e-6)
    np.testing.assert_allclose(float(t1 - t0), 0.25 * float(df + dc), rtol=1e-4)
    # the directions do not matter for the penalty
    t2, _, _ = T.losses(nerf, nerf, params, BBOX_MIN, BBOX_MAX, batch, uc, uf, 64, 128,
                        density_penalty=0.25, density_points=(coords, -dirs))
    assert float(t2) == float(t1)


def test_ngp_refnerf_oracle_normals_match_finite_differences():
    """InstantNGPRefNERFModel (instant_ngp.py:57-89): the spatial block's input gradient (what
    RefNERFBase takes with jax.grad, ref_nerf.py:38-43) equals a central finite difference of
    -spatial_out[:, 0] in fp64 through the SMOOTH hash grid, and the aux loss normal_mse is
    |normalize(out[6:9]) - normalize(grad)|^2."""
    from oracle import models_torch as M
    L = 6
    o = M.InstantNGPRefNERFModel([2 ** 12] * L, [2 ** (2 + i // 2) for i in range(L)], BBOX_MIN, BBOX_MAX)
    assert o.layer_dims() == [(12, 64), (64, 16), (33, 64), (64, 64), (64, 3)]
    p = o.init(torch.Generator().manual_seed(1))
    for leaf in p["MultiresHashTableEncoding_0"].values():
        leaf["table"] *= 1e4
    p64 = M.tree_map(lambda t: t.double(), p)
    rs = np.random.RandomState(2)
    x = torch.from_numpy(rs.uniform(-0.9, 0.9, (40, 3)))
    d = torch.from_numpy(rs.randn(40, 3))
    d = d / d.norm(dim=1, keepdim=True)
    xg = x.clone().requires_grad_(True)
    out = o.spatial_block(p64, xg)
    (grad,) = torch.autograd.grad(-out[:, 0].sum(), xg)
    h = 1e-6
    fd = torch.zeros_like(x)
    for a in range(3):
        e = torch.zeros(3, dtype=torch.float64)
        e[a] = h
        with torch.no_grad():
            fd[:, a] = -(o.spatial_block(p64, x + e)[:, 0] - o.spatial_block(p64, x - e)[:, 0]) / (2 * h)
    ok = (grad - fd).abs().max(dim=1).values < 1e-5 * (1 + grad.abs().max())
    assert ok.float().mean() > 0.9  # the rest straddle a cell face or a ReLU kink within +-h
    _, _, aux = o.apply(p64, x, d, create_graph=False)
    n = out[:, 6:9].detach()
    n = n / torch.sqrt((n ** 2).sum(1, keepdim=True) + 1e-10)
    rn = grad / torch.sqrt((grad ** 2).sum(1, keepdim=True) + 1e-10)
    np.testing.assert_allclose(aux["normal_mse"].numpy(), ((n - rn) ** 2).sum(1).numpy(), rtol=1e-9, atol=1e-12)


def test_golden_ngpref_fixture():
    """Regression of the InstantNGPRefNERFModel restatement and of the density-penalty branch
    against tests/golden/ngpref_small.npz (generated by make_golden.ngpref_fixture from this oracle)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(os.path.join(GOLDEN, "ngpref_small.npz"))
    x, d, o, p = mg.ngpref_case()
    np.testing.assert_array_equal(x, g["x"])
    de, rgb, aux = o.apply(p, torch.from_numpy(x), torch.from_numpy(d), create_graph=False)
    np.testing.assert_allclose(de.numpy(), g["dens"], rtol=1e-5)
    np.testing.assert_allclose(rgb.numpy(), g["rgb"], atol=1e-6)
    np.testing.assert_allclose(aux["normal_mse"].numpy(), g["normal_mse"], atol=1e-5)
    np.testing.assert_allclose(aux["neg_normal"].numpy(), g["neg_normal"], atol=1e-6)
    nerf = M.NeRFModel()
    params = T.init_params(nerf, nerf, 7)
    batch, uc, uf = make_rays(8, seed=63), make_uniforms(8, 64, 64), make_uniforms(8, 128, 65)
    total, ld, _ = T.losses(nerf, nerf, params, BBOX_MIN, BBOX_MAX, batch, uc, uf, 64, 128, density_penalty=0.1,
                            density_points=(x, d))
    np.testing.assert_allclose(float(total), float(g["penalty_total"]), rtol=1e-5)
    np.testing.assert_allclose(float(ld["fine_density"]), float(g["penalty_fine"]), rtol=1e-5)


# ------------------------------------------------------------------ property tests
def test_interp_rows_matches_np_interp():
    """render.py:251 calls jnp.interp, documented as NumPy's algorithm: cross-check the fp32
    restatement (_interp_rows) against np.interp (float64 arithmetic) on monotone CDFs with flat
    bins, queries on bin edges and at xp[0].  One documented difference is excluded: for x == xp[-1]
    behind a flat last bin jnp.interp's formula (clip the index, dx == 0 -> fp[i-1]) returns fp[-2]
    where NumPy returns fp[-1]; the restatement follows jnp (checked separately below)."""
    rs = np.random.RandomState(0)
    n, k, m = 64, 65, 128
    w = rs.gamma(0.3, 1.0, (n, k - 1)) * (rs.uniform(size=(n, k - 1)) < 0.6)  # many empty (flat) bins
    w[0] = 0.0
    w[0, 7] = 1.0
    xp = np.concatenate([np.zeros((n, 1)), np.cumsum(w + 1e-8, axis=1)], axis=1)
    xp = (xp / xp[:, -1:]).astype(F)
    fp = np.sort(rs.uniform(2.0, 6.0, (n, k)), axis=1).astype(F)
    x = rs.uniform(0, 1, (n, m)).astype(F)
    x[:, 0] = 0.0
    x[:, 1] = 1.0
    x[:, 2] = xp[:, 10]  # exactly on a bin edge: searchsorted side="right"
    got, idx = render_np._interp_rows(x, xp, fp)
    for r in range(n):
        want = np.interp(x[r].astype(np.float64), xp[r].astype(np.float64), fp[r].astype(np.float64))
        below_end = x[r] < xp[r, -1]
        np.testing.assert_allclose(got[r][below_end], want[below_end], rtol=0, atol=4e-6 * 6.0, err_msg=f"row {r}")
        i = idx[r]
        at_end = ~below_end
        flat_last = xp[r, -1] == xp[r, -2]
        np.testing.assert_array_equal(got[r][at_end], np.full(at_end.sum(), fp[r, -2] if flat_last else fp[r, -1]))
        assert ((1 <= i) & (i <= k - 1)).all()
        # the bin the restatement picked brackets the query (left-closed, right-open, clamped at the ends)
        inside = (x[r] > xp[r, 0]) & (x[r] < xp[r, -1])
        assert (xp[r, i - 1][inside] <= x[r][inside]).all() and (x[r][inside] < xp[r, i][inside]).all()


try:
    from hypothesis import given, settings, strategies as st
    HAVE_HYPOTHESIS = True
except Exception:  # noqa: BLE001
    HAVE_HYPOTHESIS = False

if HAVE_HYPOTHESIS:
    _coord = st.floats(-6.0, 6.0, width=32, allow_nan=False)
    _dirc = st.sampled_from([0.0, 1.0, -1.0, 1e-8, -1e-8, 0.25, -0.7, 3e-5, 0.5])

    @settings(max_examples=150, deadline=None)
    @given(o=st.tuples(_coord, _coord, _coord), d=st.tuples(_dirc, _dirc, _dirc))
    def test_ray_t_range_properties(o, d):
        """render.py:346-389 on arbitrary origins / degenerate directions: t_max - t_min >= ~1e-3,
        0 <= t_min, masked rays get exactly (0, 1e-3), and for hits both endpoints lie on or inside
        the (slightly inflated) box."""
        rays = np.array([[o, d]], F)
        t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
        degenerate = any(np.float32(c) + np.float32(1e-8) == 0 for c in d)  # d + eps == 0: the reference divides by 0
        assert t_min[0] >= 0 and np.isfinite(t_min[0]) and (np.isfinite(t_max[0]) or degenerate)
        assert t_max[0] - t_min[0] >= np.float32(1e-3) * np.float32(0.999)
        if not mask[0]:
            assert t_min[0] == 0 and t_max[0] == np.float32(1e-3)
        elif np.linalg.norm(d) > 0.2:
            p0 = np.asarray(o, np.float64) + np.asarray(d, np.float64) * float(t_min[0])
            assert (np.abs(p0) <= 1.0 + 1e-3 * max(1.0, np.abs(o).max())).all() or t_min[0] == 0

    @settings(max_examples=60, deadline=None)
    @given(seed=st.integers(0, 2 ** 31 - 1), kind=st.sampled_from(["gamma", "zero", "spike", "tiny", "huge"]),
           tc=st.sampled_from([8, 33, 64]), tf=st.sampled_from([5, 128]))
    def test_fine_sampling_properties(seed, kind, tc, tf):
        """render.py:211-257: the combined fine set is sorted, has Tc + Tf entries, contains every
        coarse position, stays inside [t_min, t_max]; with zero density the CDF is linear in the bin
        index, so the new samples are the piecewise-linear map of the stratified inputs through the bin
        ends: monotone, and sample j lies in the bins that [j, j+1) / Tf covers."""
        rs = np.random.RandomState(seed)
        n = 6
        rays = make_rays(n, seed=seed % 1000, miss_frac=0.3, with_targets=False)
        t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
        u_c = (rs.randint(0, 2 ** 23, (n, tc)) * 2.0 ** -23).astype(F)
        cs = render_np.RaySamples.stratified_sampling(t_min, t_max, mask, tc, u_c)
        dens = {"gamma": lambda: rs.gamma(0.5, 4.0, (n, tc)), "zero": lambda: np.zeros((n, tc)),
                "spike": lambda: np.eye(tc)[rs.randint(0, tc, n)] * 1e6, "tiny": lambda: np.full((n, tc), 1e-12),
                "huge": lambda: np.full((n, tc), 1e5)}[kind]().astype(F)
        u = (rs.randint(0, 2 ** 23, (n, tf)) * 2.0 ** -23).astype(F)
        fs, idx = cs.fine_sampling(tf, u, dens, return_indices=True)
        assert fs.ts.shape == (n, tc + tf)
        assert (np.diff(fs.ts, axis=1) >= 0).all()
        assert ((1 <= idx) & (idx <= tc)).all()
        lo, hi = t_min[:, None], t_max[:, None]
        slack = np.float32(1e-5) * np.maximum(np.float32(1.0), np.abs(hi))
        assert (fs.ts >= lo - slack).all() and (fs.ts <= hi + slack).all()
        for r in range(n):  # union: every coarse position survives the sort
            assert np.isin(cs.ts[r], fs.ts[r]).all()
        if kind == "zero":
            new_only = cs.fine_sampling(tf, u, dens, combine=False).ts
            assert (np.diff(new_only, axis=1) >= 0).all()
            ys = np.concatenate([t_min[:, None], cs.ends()], axis=1)  # bin edges, [n, tc + 1]
            j = np.arange(tf)
            b_lo = np.floor(j * tc / tf).astype(int)
            b_hi = np.minimum(np.ceil((j + 1) * tc / tf).astype(int), tc)
            assert (new_only >= ys[:, b_lo] - slack).all() and (new_only <= ys[:, b_hi] + slack).all()
