"""GPU parity tests of the Ref-NeRF path (K9: lnrf_refnerf_fwd / _bwd) against the CPU oracle
(oracle.models_torch.RefNERFModel restating learn_nerf/ref_nerf.py, incl. the input-gradient
normals and the second-order term that training differentiates through)."""
import numpy as np
import pytest
import torch

from helpers import BBOX_MAX, BBOX_MIN, F, make_rays, make_uniforms, reproducible

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def setup(seed=5, bias_scale=0.1):
    from learn_nerf.ref_nerf import RefNERFModel
    from oracle import models_torch as M
    o = M.RefNERFModel()
    p = o.init(torch.Generator().manual_seed(seed))
    rs = np.random.RandomState(seed)
    for leaf in p.values():  # non-zero biases so that every path is exercised
        leaf["bias"] = torch.from_numpy((bias_scale * rs.randn(*leaf["bias"].shape)).astype(F))
    n = RefNERFModel()
    tree = n.flatten_params({k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in p.items()})
    return o, n, p, tree


def test_refnerf_layout():
    from learn_nerf import _native
    _, n, p, tree = setup()
    assert n.param_count() == 592_771 == _native.refnerf_param_count()
    assert tree["Dense_9"]["kernel"].shape == (273, 128) and tree["Dense_10"]["kernel"].shape == (128, 3)
    offs = _native.refnerf_param_offsets()
    assert all(o % 4 == 0 for o in offs)
    # 3 zero pad rows after Dense_9's 273 kernel rows
    assert offs[19] - offs[18] == 276 * 128


@pytest.mark.parametrize("m", [1, 257, 3000])
def test_refnerf_apply_vs_oracle(m):
    o, n, p, tree = setup()
    rs = np.random.RandomState(m)
    x = rs.uniform(-1, 1, (m, 3)).astype(F)
    d = rs.randn(m, 3).astype(F)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o_d, o_rgb, o_aux = reproducible(lambda: o.apply(p, torch.from_numpy(x), torch.from_numpy(d), create_graph=False))
    dens, rgb, aux = n.apply(dict(params=tree), dev(x), dev(d))
    assert dens.shape == (m, 1) and rgb.shape == (m, 3) and set(aux) == {"normal_mse", "neg_normal"}
    np.testing.assert_allclose(dens.cpu().numpy(), o_d.numpy(), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(rgb.cpu().numpy(), o_rgb.numpy(), atol=1e-5)
    # normal_mse compares two unit vectors, one of them the normalised input gradient of a
    # 9-layer MLP through 2^9-frequency sinusoids: fp32 noise on it is ~1e-4 relative
    # ... and the input gradient is DISCONTINUOUS at ReLU kinks: a unit whose pre-activation is
    # within fp32 noise of 0 may get a different mask on the two sides, which moves that sample's
    # normal by O(1e-2).  Allow that for at most 0.2 % of the samples.
    err = np.abs(aux["normal_mse"].cpu().numpy() - o_aux["normal_mse"].numpy())
    assert (err > 2e-4).mean() <= 2e-3, (err > 2e-4).sum()
    assert np.median(err) < 1e-5
    np.testing.assert_allclose(aux["neg_normal"].cpu().numpy(), o_aux["neg_normal"].numpy(), atol=1e-5)


def test_refnerf_backward_vs_autograd():
    """d params of a random linear functional of (density, rgb, normal_mse, neg_normal),
    against fp64 autograd with create_graph (double backward through the normals)."""
    from learn_nerf import _native
    from oracle import models_torch as M
    o, n, p, tree = setup(seed=9)
    m = 1500
    rs = np.random.RandomState(3)
    rays = make_rays(m // 4 + 1, seed=2)[:, :2]
    x = rs.uniform(-1, 1, (m, 3)).astype(F)
    d = rs.randn(m, 3).astype(F)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    cd, cr = rs.randn(m).astype(F) * 1e-2, rs.randn(m, 3).astype(F)
    cm, cn = rs.randn(m).astype(F), rs.randn(m).astype(F)

    def oracle_grads(dtype):
        pp = M.tree_map(lambda t: t.to(dtype).clone().requires_grad_(True), p)
        de, rgb, aux = o.apply(pp, torch.from_numpy(x).to(dtype), torch.from_numpy(d).to(dtype), create_graph=True)
        loss = ((de[:, 0] * torch.from_numpy(cd).to(dtype)).sum() + (rgb * torch.from_numpy(cr).to(dtype)).sum()
                + (aux["normal_mse"] * torch.from_numpy(cm).to(dtype)).sum()
                + (aux["neg_normal"] * torch.from_numpy(cn).to(dtype)).sum())
        leaves = M.tree_leaves(pp)
        gs = torch.autograd.grad(loss, [t for _, t in leaves])
        return {path: g for (path, _), g in zip(leaves, gs)}

    g64, g32 = oracle_grads(torch.float64), oracle_grads(torch.float32)
    # native: forward with save on explicit points (T = 1 "rays" of one sample each are not needed:
    # the C ABI takes x/d directly)
    ws = torch.empty(_native.refnerf_workspace_bytes(m, True), dtype=torch.uint8, device="cuda")
    dens, rgb = torch.empty(m, device="cuda"), torch.empty(m, 3, device="cuda")
    a1, a2 = torch.empty(m, device="cuda"), torch.empty(m, device="cuda")
    xd, dd = dev(x), dev(d)
    _native.refnerf_fwd(tree.flat, xd, dd, None, None, m, 1, True, ws, dens, rgb, a1, a2)
    g = torch.zeros_like(tree.flat)
    _native.refnerf_bwd(tree.flat, xd, dd, None, None, m, 1, ws, dev(cd), dev(cr), dev(cm), dev(cn), g)
    gt = n.bind(g)
    worst = []
    for path, ref in g64.items():
        lname, k = path.split("/")
        e_gpu = rel_l2(gt[lname][k].cpu().numpy(), ref.numpy())
        e_cpu = rel_l2(g32[path].numpy(), ref.numpy())
        worst.append((e_gpu, e_cpu, path))
    worst.sort(reverse=True)
    print("worst Ref-NeRF grad rel-L2 (gpu-vs-fp64, cpu32-vs-fp64):", worst[:5])
    # stated tolerance: rel-L2 <= 1e-3 per tensor and no worse than 3x the CPU fp32 autograd error
    assert worst[0][0] < 1e-3, worst[:5]
    assert worst[0][0] < 3 * max(w[1] for w in worst) + 1e-6, worst[:5]
    # pad rows of Dense_9 never receive gradient
    offs = _native.refnerf_param_offsets()
    assert float(g[offs[18] + 273 * 128: offs[18] + 276 * 128].abs().max()) == 0.0


def test_refnerf_train_step_vs_oracle():
    """TrainLoop with RefNERFModel coarse + fine (train_nerf.py:162-168): logged losses incl. the
    aux terms, and gradients of total = mse + 3e-4 normal_mse + 0.1 neg_normal per level."""
    from learn_nerf.ref_nerf import RefNERFModel
    from learn_nerf.train import TrainLoop
    from oracle import models_torch as M
    from oracle import train_torch as T
    o = M.RefNERFModel()
    params = T.init_params(o, o, 4)
    n = 96
    batch = make_rays(n, seed=31, miss_frac=0.2)
    uc, uf = make_uniforms(n, 64, 32), make_uniforms(n, 128, 33)
    nc, nf = RefNERFModel(), RefNERFModel()
    loop = TrainLoop(nc, nf, init_rng=0, lr=1e-4, coarse_ts=64, fine_ts=128, ray_chunk=40)
    for name in ("coarse", "fine"):
        for lname, leaf in params[name].items():
            for k in ("kernel", "bias"):
                loop.state.params[name][lname][k].copy_(leaf[k])
    loop.state.params["background"].copy_(params["background"])
    step = loop.step_fn(BBOX_MIN, BBOX_MAX)
    fine_ts = loop._renderer(list(BBOX_MIN), list(BBOX_MAX), loop.state.params).render_rays(
        (dev(uc), dev(uf)), dev(batch[:, :2]), _save=True)["fine"]["_ts"].ts.cpu().numpy()
    g, ld, _ = T.grads(o, o, params, BBOX_MIN, BBOX_MAX, batch, uc, uf, 64, 128,
                       fixed_fine_ts=fine_ts, dtype=torch.float64)
    g32, _, _ = T.grads(o, o, params, BBOX_MIN, BBOX_MAX, batch, uc, uf, 64, 128, fixed_fine_ts=fine_ts)
    logs = step((dev(uc), dev(uf)), dev(batch))
    assert set(logs) == {"coarse", "fine", "coarse_normal_mse", "coarse_neg_normal", "fine_normal_mse",
                         "fine_neg_normal", "grad_norm", "param_norm"}
    for k in ("coarse", "fine", "coarse_normal_mse", "coarse_neg_normal", "fine_normal_mse", "fine_neg_normal"):
        np.testing.assert_allclose(float(logs[k]), ld[k], rtol=2e-3, atol=1e-7, err_msg=k)
    np.testing.assert_allclose(float(logs["grad_norm"]), T.tree_norm(g), rtol=2e-3)
    grads = loop._grads
    worst = []
    for name, model in (("coarse", nc), ("fine", nf)):
        gt = model.bind(grads[loop._slices[name][0]:loop._slices[name][1]])
        for lname, leaf in g[name].items():
            for k in ("kernel", "bias"):
                e_gpu = rel_l2(gt[lname][k].cpu().numpy(), leaf[k].numpy())
                e_cpu = rel_l2(g32[name][lname][k].numpy(), leaf[k].numpy())
                worst.append((e_gpu, e_cpu, name, lname, k))
    worst.sort(reverse=True)
    print("worst Ref-NeRF train grad rel-L2 (gpu, cpu-fp32):", worst[:4])
    # stated tolerance: per tensor rel-L2 vs fp64 <= 2e-2 and no worse than 2x what the CPU fp32
    # autograd of the same graph shows (measured: GPU 6.8e-3, CPU fp32 6.3e-3 on coarse Dense_0)
    assert worst[0][0] < 2e-2, worst[:4]
    assert worst[0][0] < 2 * max(w[1] for w in worst) + 1e-5, worst[:4]
