"""Parity at BASELINE.json's FULL sizes.  Two kinds of checks:
(1) against the ORACLE at 4096 rays (configs[1]: train step, fp32 and bf16, every gradient tensor vs
fp64 autograd in ray chunks; a 4096-ray render vs oracle.render_np) and on a 2,048-ray subsample of one
32,768-ray Instant-NGP launch (configs[2]);
(2) size-independent properties where the oracle would take minutes: sortedness and range of the
sample positions, the compositing identity recomputed with a plain torch fp32 reference from the
kernels' own densities / colours, chunk independence of render_rays at 65,536 rays (configs[4]),
linearity of the backward in the upstream gradient, and the hash-grid partition-of-unity."""
import numpy as np
import pytest
import torch

from helpers import BBOX_MAX, BBOX_MIN, F, make_rays, make_uniforms, reproducible

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _renderer(precision, seed=0):
    from learn_nerf.model import NeRFModel
    from learn_nerf.render import NeRFRenderer
    coarse, fine = NeRFModel(precision=precision), NeRFModel(precision=precision)
    pc = coarse.init(seed, device="cuda")["params"]
    pf = fine.init(seed + 1, device="cuda")["params"]
    return NeRFRenderer(coarse=coarse, fine=fine, coarse_params=pc, fine_params=pf,
                        background=torch.tensor([-1.0, 0.25, 0.5], device="cuda"), bbox_min=BBOX_MIN,
                        bbox_max=BBOX_MAX, coarse_ts=64, fine_ts=128)


def _torch_composite(ts, t_min, t_max, mask, dens, rgb, bg):
    """render.py:155-190, 259-287 restated with torch ops on the GPU (fp32 reference of K3)."""
    mid = (ts[:, 1:] + ts[:, :-1]) * 0.5
    starts = torch.cat([t_min[:, None], mid], dim=1)
    ends = torch.cat([mid, t_max[:, None]], dim=1)
    a = dens * (ends - starts)
    acc = torch.cumsum(a, dim=1)
    prev = torch.cat([torch.zeros_like(acc[:, :1]), acc[:, :-1]], dim=1)
    p = torch.exp(-prev) * (1 - torch.exp(-a))
    p_esc = torch.exp(-acc[:, -1:])
    out = (p[..., None] * rgb).sum(1) + p_esc * bg
    out = torch.where(mask[:, None], out, bg.expand_as(out))
    alpha = torch.where(mask[:, None], 1 - p_esc, torch.zeros_like(p_esc))
    return out, alpha, p, p_esc


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_render_4096_rays_properties(precision):
    n = 4096
    r = _renderer(precision)
    batch = make_rays(n, seed=21, miss_frac=0.3, with_targets=False)
    uc, uf = make_uniforms(n, 64, 22), make_uniforms(n, 128, 23)
    out = r.render_rays((dev(uc), dev(uf)), dev(batch), _save=True)
    bg = r.background
    for level, T in (("coarse", 64), ("fine", 192)):
        lv = out[level]
        s = lv["_ts"]
        ts, mask = s.ts, s.mask.bool()
        assert ts.shape == (n, T)
        # sample positions: sorted, inside [t_min, t_max] for rays that hit the box
        assert bool((ts[:, 1:] >= ts[:, :-1]).all()), f"{level}: positions not sorted"
        hit = mask
        assert bool((ts[hit] >= s.t_min[hit][:, None]).all()) and bool((ts[hit] <= s.t_max[hit][:, None]).all())
        assert 0.6 < float(hit.float().mean()) < 0.8  # ~30 % of the rays were aimed away from the box
        # compositing identity on the kernels' own densities / colours (torch fp32 reference)
        ref_out, ref_alpha, p, p_esc = _torch_composite(ts, s.t_min, s.t_max, mask, lv["densities"], lv["rgbs"], bg)
        torch.testing.assert_close(lv["outputs"], ref_out, atol=2e-5, rtol=0)
        torch.testing.assert_close(lv["alphas"], ref_alpha, atol=2e-5, rtol=0)
        # termination probabilities sum to one, alphas in [0, 1], missed rays show the background
        torch.testing.assert_close(p.sum(1, keepdim=True)[hit] + p_esc[hit], torch.ones_like(p_esc[hit]), atol=1e-5,
                                   rtol=0)
        assert bool((lv["alphas"] >= 0).all()) and bool((lv["alphas"] <= 1).all())
        assert bool((lv["outputs"][~hit] == bg).all()) and bool((lv["alphas"][~hit] == 0).all())
        assert bool((lv["densities"] >= 0).all()) and bool(lv["rgbs"].abs().max() <= 1)
    # the first 64 fine... the fine set contains every coarse position (render.py:253-255: union, then sort)
    cts, fts = out["coarse"]["_ts"].ts, out["fine"]["_ts"].ts
    pos = torch.searchsorted(fts.contiguous(), cts.contiguous())
    assert bool((torch.gather(fts, 1, pos.clamp(max=191)) == cts).all())


def test_render_chunk_independence_65536_rays():
    """render_nerf.py:88-92 renders an image in chunks: with the uniforms given, a 65,536-ray call
    equals the concatenation of four 16,384-ray calls bit for bit (no cross-ray coupling, no
    dependence on tile / CTA assignment), on the bf16 tcgen05 path."""
    n = 65536
    r = _renderer("bf16", seed=4)
    batch = dev(make_rays(n, seed=31, miss_frac=0.1, with_targets=False))
    uc, uf = dev(make_uniforms(n, 64, 32)), dev(make_uniforms(n, 128, 33))
    whole = r.render_rays((uc, uf), batch)["fine"]
    parts = [r.render_rays((uc[a:a + 16384].contiguous(), uf[a:a + 16384].contiguous()),
                           batch[a:a + 16384].contiguous())["fine"] for a in range(0, n, 16384)]
    for k in ("outputs", "alphas", "coords", "densities"):
        got = torch.cat([p[k] for p in parts], dim=0)
        assert torch.equal(got, whole[k]), f"{k} depends on the chunking"
    assert bool(torch.isfinite(whole["outputs"]).all())


@pytest.mark.parametrize("precision,tol", [("fp32", 5e-5), ("bf16", 2e-2)])
def test_train_step_4096_rays_properties(precision, tol):
    """configs[1] size.  (a) the logged losses equal the MSE of a forward-only render of the same
    parameters; (b) the gradient norm the fused Adam kernel reports equals the norm of the gradient
    buffer; (c) MLP backward is linear in the upstream gradient: scaling d_dens, d_rgb by 2 doubles
    every parameter gradient (rel-L2 1e-5 fp32: the split-K dW sums are atomic, so two runs differ by
    fp32 reordering noise; the bf16 path rounds g tiles, so 1e-2)."""
    from learn_nerf.model import NeRFModel
    from learn_nerf.train import TrainLoop
    n = 4096
    loop = TrainLoop(NeRFModel(precision=precision), NeRFModel(precision=precision), init_rng=5, lr=1e-4,
                     coarse_ts=64, fine_ts=128)
    batch = dev(make_rays(n, seed=41))
    uc, uf = dev(make_uniforms(n, 64, 42)), dev(make_uniforms(n, 128, 43))
    rend = loop._renderer(list(BBOX_MIN), list(BBOX_MAX), loop.state.params)
    fwd = rend.render_rays((uc, uf), batch[:, :2].contiguous())
    want = {lv: float(((fwd[lv]["outputs"] - batch[:, 2]) ** 2).mean()) for lv in ("coarse", "fine")}
    logs = loop.step_fn(BBOX_MIN, BBOX_MAX)((uc, uf), batch)
    for lv in ("coarse", "fine"):
        assert abs(float(logs[lv]) - want[lv]) <= tol * max(1.0, want[lv]), (lv, float(logs[lv]), want[lv])
    gnorm = float(loop._grads.double().norm())
    assert abs(float(logs["grad_norm"]) - gnorm) <= 1e-4 * gnorm
    assert np.isfinite(gnorm) and gnorm > 0
    # (c) linearity of lnrf_nerf_mlp_bwd at 4096 x 192 samples
    model = loop.fine
    tree = loop.state.params["fine"]
    rays = batch[:, :2].contiguous()
    ts = torch.rand(n, 192, device="cuda").sort(dim=1).values * 2 + 3
    gen = torch.Generator(device="cuda").manual_seed(3)
    d_dens = torch.randn(n, 192, device="cuda", generator=gen) * 1e-3
    d_rgb = torch.randn(n, 192, 3, device="cuda", generator=gen) * 1e-3
    grads = []
    for scale in (1.0, 2.0):
        _, _, _, ctx = model.apply_rays(tree, rays, ts, save=True, slot="lin")
        g = torch.zeros_like(tree.flat)
        model.backward_rays(ctx, d_dens * scale, d_rgb * scale, g)
        grads.append(g)
    rel = float((grads[1] - 2 * grads[0]).norm() / (2 * grads[0]).norm())
    assert rel < (1e-5 if precision == "fp32" else 1e-2), rel


def test_ngp_32768_rays_properties():
    """configs[2] size: (a) a hash grid whose tables are constant per level encodes every point to
    that constant (the eight trilinear weights sum to one, instant_ngp.py:165-176), for the dense,
    the hashed and the smooth variant; (b) the scatter-add backward conserves mass: the sum over
    all table rows of dL/dtable equals the sum over points of dL/denc, level by level; (c) one
    train step at 32,768 rays logs finite losses and a gradient norm that matches the buffer."""
    from learn_nerf import _native
    from learn_nerf.instant_ngp import InstantNGPModel
    from learn_nerf.train import TrainLoop
    L = 16
    grids = [2 ** (4 + i // 2) for i in range(L)]
    n, T = 32768, 16
    rays = dev(make_rays(n, seed=51, with_targets=False))
    ts = torch.rand(n, T, device="cuda").sort(dim=1).values * 2 + 3
    for smooth in (False, True):
        m = InstantNGPModel(table_sizes=[2 ** 18] * L, grid_sizes=grids, bbox_min=BBOX_MIN, bbox_max=BBOX_MAX,
                            table_smooth=smooth)
        tree = m.init(0, device="cuda")["params"]
        for l in range(L):
            tab = tree["MultiresHashTableEncoding_0"][f"HashTableEncoding_{l}"]["table"]
            tab[:, 0] = 0.5 + l
            tab[:, 1] = -(0.25 + l)
        enc = torch.empty(n * T, 2 * L, device="cuda")
        _native.hashgrid_fwd(tree.flat, m.spec(), None, rays, ts, n, T, enc)
        want = torch.tensor([[0.5 + l, -(0.25 + l)] for l in range(L)], device="cuda").reshape(-1)
        torch.testing.assert_close(enc, want.expand_as(enc), atol=2e-5, rtol=2e-6)
        # (b) mass conservation of the scatter
        gen = torch.Generator(device="cuda").manual_seed(7)
        d_enc = torch.randn(n * T, 2 * L, device="cuda", generator=gen)
        g = torch.zeros_like(tree.flat)
        _native.hashgrid_bwd(m.spec(), None, rays, ts, n, T, d_enc, g)
        gt = m.bind(g)
        for l in (0, 5, 6, 15):
            got = gt["MultiresHashTableEncoding_0"][f"HashTableEncoding_{l}"]["table"].double().sum(0)
            ref = d_enc[:, 2 * l:2 * l + 2].double().sum(0)
            assert float((got - ref).abs().max()) <= 1e-3 * float(d_enc.abs().sum(0).max()), (smooth, l)
    mk = lambda levels: InstantNGPModel(table_sizes=[2 ** 18] * levels, grid_sizes=grids[:levels],
                                        bbox_min=BBOX_MIN, bbox_max=BBOX_MAX)
    loop = TrainLoop(mk(6), mk(16), init_rng=1, lr=1e-3, coarse_ts=64, fine_ts=128, adam_eps=1e-15,
                     adam_b1=0.9, adam_b2=0.99)
    batch = dev(make_rays(n, seed=52))
    logs = loop.step_fn(BBOX_MIN, BBOX_MAX)(3, batch)
    assert all(np.isfinite(float(v)) for v in logs.values())
    gnorm = float(loop._grads.double().norm())
    assert abs(float(logs["grad_norm"]) - gnorm) <= 1e-4 * gnorm and gnorm > 0


@pytest.mark.parametrize("family", ["nerf-bf16", "nerf-fp32", "ngp", "refnerf", "ngpref"])
def test_empty_and_tiny_batches(family):
    """The other end of the size range: 0, 1 and 3 rays through render_rays and (n > 0) a train step,
    for every model family -- empty launches are skipped inside the library, a single ray fills one
    partial 128-sample tile pair."""
    from learn_nerf.instant_ngp import InstantNGPModel, InstantNGPRefNERFModel
    from learn_nerf.model import NeRFModel
    from learn_nerf.ref_nerf import RefNERFModel
    from learn_nerf.train import TrainLoop
    grids = lambda L: [2 ** (4 + i // 2) for i in range(L)]
    kw = lambda L: dict(table_sizes=[2 ** 14] * L, grid_sizes=grids(L), bbox_min=BBOX_MIN, bbox_max=BBOX_MAX)
    mk = {"nerf-bf16": lambda: (NeRFModel(precision="bf16"), NeRFModel(precision="bf16")),
          "nerf-fp32": lambda: (NeRFModel(), NeRFModel()),
          "ngp": lambda: (InstantNGPModel(**kw(6)), InstantNGPModel(**kw(16))),
          "refnerf": lambda: (RefNERFModel(), RefNERFModel()),
          "ngpref": lambda: (InstantNGPRefNERFModel(**kw(6)), InstantNGPRefNERFModel(**kw(16)))}[family]
    for n in (0, 1, 3):
        c, f = mk()
        loop = TrainLoop(c, f, init_rng=1, lr=1e-4, coarse_ts=64, fine_ts=128)
        batch = dev(make_rays(max(n, 1), seed=1)[:n])
        r = loop._renderer(list(BBOX_MIN), list(BBOX_MAX), loop.state.params)
        out = r.render_rays(3, batch[:, :2].contiguous())
        assert out["fine"]["outputs"].shape == (n, 3) and out["coarse"]["densities"].shape == (n, 64)
        assert bool(torch.isfinite(out["fine"]["outputs"]).all())
        if n > 0:
            logs = loop.step_fn(BBOX_MIN, BBOX_MAX)(4, batch)
            assert all(np.isfinite(float(v)) for v in logs.values()), logs


# ------------------------------------------------------------------------------ vs the ORACLE
# at BASELINE sizes.  The fp64 oracle runs in ray chunks (rays are independent in the loss, so the
# gradient of the 4096-ray mean is the chunk-size-weighted sum of the chunk gradients): ~15 s of CPU.
def _oracle_chunked_grads(T, nerf, params, batch, uc, uf, fine_ts, chunk=512, dtype=torch.float64):
    """-> (grad tree, {coarse, fine} losses, per-level outputs) of the full batch, accumulated over
    ``chunk``-ray slices of oracle.train_torch.grads with the fine positions held fixed."""
    n = batch.shape[0]
    total, losses, outs = None, dict(coarse=0.0, fine=0.0), dict(coarse=[], fine=[])
    for a in range(0, n, chunk):
        b = min(a + chunk, n)
        g, ld, ro = T.grads(nerf, nerf, params, BBOX_MIN, BBOX_MAX, batch[a:b], uc[a:b], uf[a:b], 64, 128,
                            fixed_fine_ts=fine_ts[a:b], dtype=dtype)
        w = (b - a) / n
        for lv in ("coarse", "fine"):
            losses[lv] += ld[lv] * w
            outs[lv].append(ro[lv]["outputs"].detach().double().numpy())
        scaled = _tree_scale(g, w)
        total = scaled if total is None else _tree_add(total, scaled)
    return total, losses, {lv: np.concatenate(v) for lv, v in outs.items()}


def _tree_scale(t, w):
    return {k: _tree_scale(v, w) for k, v in t.items()} if isinstance(t, dict) else t.double() * w


def _tree_add(a, b):
    return {k: _tree_add(a[k], b[k]) for k in a} if isinstance(a, dict) else a + b


def _rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("precision,loss_tol,grad_tol", [("fp32", 1e-4, 5e-3), ("bf16", 2e-2, 5e-2)])
def test_train_step_4096_rays_vs_oracle(precision, loss_tol, grad_tol):
    """configs[1] at its full size (4096 rays = 1,048,576 MLP evaluations: every CTA of the persistent
    kernels, all dW waves and the stash ring wrap are exercised): logged losses, the gradient norm and
    EVERY parameter-gradient tensor of the CUDA train step against the fp64 oracle
    (oracle.train_torch.grads, train.py:85-106).  Tolerances: losses 1e-4 rel (fp32) / 2e-2 abs
    (bf16); gradients rel-L2 per tensor 5e-3 (fp32) / 5e-2 (bf16), the same as the 256-ray tests."""
    from oracle import models_torch as M
    from oracle import train_torch as T
    from learn_nerf.model import NeRFModel
    from learn_nerf.train import TrainLoop
    n = 4096
    nerf = M.NeRFModel()
    params = T.init_params(nerf, nerf, 12)
    batch, uc, uf = make_rays(n, seed=61, miss_frac=0.1), make_uniforms(n, 64, 62), make_uniforms(n, 128, 63)
    loop = TrainLoop(NeRFModel(precision=precision), NeRFModel(precision=precision), init_rng=0, lr=1e-4,
                     coarse_ts=64, fine_ts=128)
    for name in ("coarse", "fine"):
        for lname, leaf in params[name].items():
            for k in ("kernel", "bias"):
                loop.state.params[name][lname][k].copy_(leaf[k])
        loop.state.params[name].mark_updated()
    loop.state.params["background"].copy_(params["background"])
    rend = loop._renderer(list(BBOX_MIN), list(BBOX_MAX), loop.state.params)
    fwd = rend.render_rays((dev(uc), dev(uf)), dev(batch[:, :2]), _save=True)
    fine_ts = fwd["fine"]["_ts"].ts.cpu().numpy()
    g, ld, o_out = _oracle_chunked_grads(T, nerf, params, batch, uc, uf, fine_ts)
    # forward outputs at the full size (fine positions are the CUDA path's own: see
    # test_render_4096_rays_vs_oracle for the positions themselves)
    out_tol = 1e-5 if precision == "fp32" else 2e-2
    for lv in ("coarse", "fine"):
        err = float(np.abs(fwd[lv]["outputs"].cpu().numpy() - o_out[lv]).max())
        assert err <= out_tol, (lv, err)
    logs = loop.step_fn(BBOX_MIN, BBOX_MAX)((dev(uc), dev(uf)), dev(batch))
    for lv in ("coarse", "fine"):
        if precision == "fp32":
            np.testing.assert_allclose(float(logs[lv]), ld[lv], rtol=loss_tol)
        else:
            np.testing.assert_allclose(float(logs[lv]), ld[lv], atol=loss_tol)
    np.testing.assert_allclose(float(logs["grad_norm"]), T.tree_norm(g), rtol=grad_tol)
    worst = []
    for name in ("coarse", "fine"):
        a, b = loop._slices[name]
        gt = getattr(loop, name).bind(loop._grads[a:b])
        for lname, leaf in g[name].items():
            for k in ("kernel", "bias"):
                worst.append((_rel_l2(gt[lname][k].cpu().numpy(), leaf[k].numpy()), name, lname, k))
    worst.sort(reverse=True)
    print(f"4096-ray {precision} worst grad rel-L2 vs fp64 oracle:", worst[:4])
    assert worst[0][0] < grad_tol, worst[:4]
    sb = loop._slices["background"]
    assert _rel_l2(loop._grads[sb[0]:sb[1]].cpu().numpy(), g["background"].numpy()) < (1e-4 if precision == "fp32" else 2e-2)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_render_4096_rays_vs_oracle(precision, tol):
    """A 4096-ray render_rays call (render.py:39-91) against oracle.render_np.NeRFRenderer with the
    fp32 torch-CPU MLP: coarse positions bit-exact, rendered RGB / alpha / coords within the north
    star's tolerance (1e-5 abs fp32, 2e-2 abs bf16) on both levels, 30 % of the rays missing the box."""
    from oracle import models_torch as M
    from oracle import render_np
    from oracle import train_torch as T
    from learn_nerf.model import NeRFModel
    from learn_nerf.render import NeRFRenderer
    n = 4096
    nerf = M.NeRFModel()
    params = T.init_params(nerf, nerf, 13)
    batch = make_rays(n, seed=71, miss_frac=0.3, with_targets=False)
    uc, uf = make_uniforms(n, 64, 72), make_uniforms(n, 128, 73)
    smp = {}
    o = render_np.NeRFRenderer(M.as_numpy_model_fn(nerf, params["coarse"]), M.as_numpy_model_fn(nerf, params["fine"]),
                               params["background"].numpy(), BBOX_MIN, BBOX_MAX, 64, 128
                               ).render_rays(uc, uf, batch, smp)
    coarse, fine = NeRFModel(precision=precision), NeRFModel(precision=precision)
    put = lambda m, t: m.flatten_params({k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in t.items()})
    r = NeRFRenderer(coarse=coarse, fine=fine, coarse_params=put(coarse, params["coarse"]),
                     fine_params=put(fine, params["fine"]), background=params["background"].cuda(),
                     bbox_min=BBOX_MIN, bbox_max=BBOX_MAX, coarse_ts=64, fine_ts=128)
    out = r.render_rays((dev(uc), dev(uf)), dev(batch), _save=True)
    np.testing.assert_array_equal(out["coarse"]["_ts"].ts.cpu().numpy().view(np.uint32), smp["coarse"].ts.view(np.uint32))
    for lv in ("coarse", "fine"):
        for k in ("outputs", "alphas", "coords"):
            err = float(np.abs(out[lv][k].cpu().numpy() - o[lv][k]).max())
            assert err <= tol * (3 if k == "coords" and precision == "bf16" else 1), (lv, k, err)
    if precision == "fp32":  # fine positions follow the coarse densities: equal up to the density noise
        d = np.abs(out["fine"]["_ts"].ts.cpu().numpy() - smp["fine"].ts)
        assert float(np.quantile(d, 0.999)) < 1e-4, float(np.quantile(d, 0.999))


def test_ngp_32768_rays_forward_vs_oracle_subsample():
    """configs[2] size: ONE 32,768-ray x 16-sample launch of the hash grid + NGP heads; a 2,048-ray
    subsample of that launch's outputs is compared with the torch-CPU oracle (instant_ngp.py:33-54,
    134-224): densities rel 2e-5, colours 2e-5 abs."""
    from oracle import models_torch as M
    from learn_nerf.instant_ngp import InstantNGPModel
    L = 16
    grids = [2 ** (4 + i // 2) for i in range(L)]
    n, T = 32768, 16
    o_ngp = M.InstantNGPModel([2 ** 18] * L, grids, BBOX_MIN, BBOX_MAX)
    p = o_ngp.init(torch.Generator().manual_seed(81))
    for leaf in p["MultiresHashTableEncoding_0"].values():
        leaf["table"] *= 1e4
    rays = make_rays(n, seed=82, with_targets=False)
    ts = np.sort(np.random.RandomState(83).uniform(3.0, 5.0, (n, T)).astype(F), axis=1)
    ngp = InstantNGPModel(table_sizes=[2 ** 18] * L, grid_sizes=grids, bbox_min=BBOX_MIN, bbox_max=BBOX_MAX)
    cu = lambda t: {k: cu(v) if isinstance(v, dict) else v.cuda() for k, v in t.items()}
    dens, rgb, _, _ = ngp.apply_rays(ngp.flatten_params(cu(p)), dev(rays), dev(ts))
    sel = np.random.RandomState(84).choice(n, 2048, replace=False)
    pts = (rays[sel, :1] + (rays[sel, 1:2] * ts[sel][:, :, None]).astype(F)).astype(F).reshape(-1, 3)
    dirs = np.ascontiguousarray(np.broadcast_to(rays[sel, 1:2], (2048, T, 3))).reshape(-1, 3)
    with torch.no_grad():
        o_d, o_rgb, _ = reproducible(lambda: o_ngp.apply(p, torch.from_numpy(pts), torch.from_numpy(dirs)))
    got_d = dens.cpu().numpy()[sel].reshape(-1)
    got_rgb = rgb.cpu().numpy()[sel].reshape(-1, 3)
    np.testing.assert_allclose(got_d, o_d.numpy().reshape(-1), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(got_rgb, o_rgb.numpy(), atol=2e-5)
