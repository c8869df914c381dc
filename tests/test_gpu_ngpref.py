"""GPU parity tests of InstantNGPRefNERFModel (lnrf_ngpref_fwd / _bwd; instant_ngp.py:57-89 on
RefNERFBase, ref_nerf.py:34-77) against the CPU oracle: Ref-NeRF heads on a SMOOTH hash grid, whose
real_normal is the input gradient THROUGH the grid (d enc / d x) and whose training gradient has a
second-order term in the tables and in Dense_0 / Dense_1."""
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


def models(L):
    from learn_nerf.instant_ngp import InstantNGPRefNERFModel
    from oracle import models_torch as M
    grids = [2 ** (4 + i // 2) for i in range(L)]
    kw = dict(table_sizes=[2 ** 14] * L, grid_sizes=grids, bbox_min=BBOX_MIN, bbox_max=BBOX_MAX)
    return M.InstantNGPRefNERFModel(**kw), InstantNGPRefNERFModel(**kw)


def oracle_params(o, seed, table_scale=0.5, bias_scale=0.1):
    p = o.init(torch.Generator().manual_seed(seed))
    rs = np.random.RandomState(seed)
    for name, leaf in p.items():
        if name.startswith("Dense_"):  # non-zero biases so that every path is exercised
            leaf["bias"] = torch.from_numpy((bias_scale * rs.randn(*leaf["bias"].shape)).astype(F))
    for leaf in p["MultiresHashTableEncoding_0"].values():
        leaf["table"] *= table_scale / 1e-4  # O(1) tables: the encoding and its x-gradient matter
    return p


def to_native(n, p):
    def cu(t):
        return {k: cu(v) if isinstance(v, dict) else v.cuda() for k, v in t.items()}
    return n.flatten_params(cu(p))


def test_ngpref_layout():
    from learn_nerf import _native
    o, n = models(16)
    p = oracle_params(o, 0)
    tree = to_native(n, p)
    from oracle.models_torch import tree_leaves
    assert n.param_count() == sum(t.numel() for _, t in tree_leaves(p))
    assert tree["Dense_2"]["kernel"].shape == (33, 64) and tree["Dense_4"]["kernel"].shape == (64, 3)
    offs = _native.ngpref_param_offsets(16)
    assert all(v % 4 == 0 for v in offs) and offs[5] - offs[4] == 36 * 64  # 3 zero pad rows after Dense_2
    # levels 0-3 dense (16^3, 16^3 <= 2^14? no: 4096 <= 16384 dense; 32^3 = 32768 > 16384 hashed)
    assert n.spec().rows[:2] == [4096, 4096] and n.spec().rows[2] == 2 ** 14 and n.spec().smooth == 1


@pytest.mark.parametrize("L,m", [(16, 1), (16, 3000), (6, 777)])
def test_ngpref_apply_vs_oracle(L, m):
    o, n = models(L)
    p = oracle_params(o, 3 + L)
    tree = to_native(n, p)
    rs = np.random.RandomState(m)
    x = rs.uniform(-1.1, 1.1, (m, 3)).astype(F)  # a few points outside the box (clipped: zero x-gradient)
    d = rs.randn(m, 3).astype(F)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o_d, o_rgb, o_aux = reproducible(lambda: o.apply(p, torch.from_numpy(x), torch.from_numpy(d), create_graph=False))
    dens, rgb, aux = n.apply(dict(params=tree), dev(x), dev(d))
    assert dens.shape == (m, 1) and rgb.shape == (m, 3) and set(aux) == {"normal_mse", "neg_normal"}
    np.testing.assert_allclose(dens.cpu().numpy(), o_d.numpy(), rtol=5e-5, atol=1e-6)
    np.testing.assert_allclose(rgb.cpu().numpy(), o_rgb.numpy(), atol=2e-5)
    np.testing.assert_allclose(aux["neg_normal"].cpu().numpy(), o_aux["neg_normal"].numpy(), atol=2e-5)
    # normal_mse compares two unit vectors, one the normalised input gradient through the grid: it is
    # discontinuous at ReLU kinks and at cell faces (fi - floor(fi)), so allow a few outliers
    err = np.abs(aux["normal_mse"].cpu().numpy() - o_aux["normal_mse"].numpy())
    assert (err > 5e-4).mean() <= 5e-3, (err > 5e-4).sum()
    assert np.median(err) < 2e-5


@pytest.mark.parametrize("L", [16, 6])
def test_ngpref_backward_vs_autograd(L):
    """d params of a random linear functional of (density, rgb, normal_mse, neg_normal) against fp64
    autograd with create_graph (double backward through the hash grid's input Jacobian)."""
    from learn_nerf import _native
    from oracle import models_torch as M
    o, n = models(L)
    p = oracle_params(o, 20 + L)
    tree = to_native(n, p)
    m = 2000
    rs = np.random.RandomState(3)
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
    ws = torch.empty(_native.ngpref_workspace_bytes(m, L, True), dtype=torch.uint8, device="cuda")
    dens, rgb = torch.empty(m, device="cuda"), torch.empty(m, 3, device="cuda")
    a1, a2 = torch.empty(m, device="cuda"), torch.empty(m, device="cuda")
    xd, dd = dev(x), dev(d)
    _native.ngpref_fwd(tree.flat, n.spec(), xd, dd, None, None, m, 1, True, ws, dens, rgb, a1, a2)
    g = torch.zeros_like(tree.flat)
    _native.ngpref_bwd(tree.flat, n.spec(), xd, dd, None, None, m, 1, ws, dev(cd), dev(cr), dev(cm), dev(cn), g)
    gt = n.bind(g)

    def leaf_of(tree_, path):
        node = tree_
        for part in path.split("/"):
            node = node[part]
        return node

    worst = []
    for path, ref in g64.items():
        e_gpu = rel_l2(leaf_of(gt, path).cpu().numpy(), ref.numpy())
        e_cpu = rel_l2(g32[path].numpy(), ref.numpy())
        worst.append((e_gpu, e_cpu, path))
    worst.sort(reverse=True)
    print("worst NGP-Ref grad rel-L2 (gpu-vs-fp64, cpu32-vs-fp64):", worst[:5])
    # stated tolerance: rel-L2 <= 5e-3 per tensor and no worse than 3x the CPU fp32 autograd error of the
    # same graph (measured 2.1e-3 on one table for BOTH: a few samples sit on a ReLU kink / cell face
    # where fp32 and fp64 pick different branches of the piecewise-constant second derivative)
    assert worst[0][0] < 5e-3, worst[:5]
    assert worst[0][0] < 3 * max(w[1] for w in worst) + 1e-5, worst[:5]
    offs = _native.ngpref_param_offsets(L)
    assert float(g[offs[4] + 33 * 64: offs[4] + 36 * 64].abs().max()) == 0.0  # pad rows of Dense_2


def test_ngpref_train_step_vs_oracle():
    """TrainLoop with InstantNGPRefNERFModel coarse (L=6) + fine (L=16) as create_model builds it
    (train_nerf.py:141-170): logged losses incl. the aux terms; gradients vs fp64 autograd."""
    from learn_nerf.train import TrainLoop
    from oracle import train_torch as T
    oc, nc = models(6)
    of, nf = models(16)
    params = dict(coarse=oracle_params(oc, 1), fine=oracle_params(of, 2), background=torch.tensor([-1.0, -1.0, -1.0]))
    n = 64
    batch = make_rays(n, seed=31, miss_frac=0.2)
    uc, uf = make_uniforms(n, 64, 32), make_uniforms(n, 128, 33)
    loop = TrainLoop(nc, nf, init_rng=0, lr=1e-4, coarse_ts=64, fine_ts=128)

    def put(dst, src):
        for k, v in dst.items():
            if isinstance(v, dict):
                put(v, src[k])
            else:
                v.copy_(src[k])
    for name in ("coarse", "fine"):
        put(loop.state.params[name], params[name])
    loop.state.params["background"].copy_(params["background"])
    step = loop.step_fn(BBOX_MIN, BBOX_MAX)
    fine_ts = loop._renderer(list(BBOX_MIN), list(BBOX_MAX), loop.state.params).render_rays(
        (dev(uc), dev(uf)), dev(batch[:, :2]), _save=True)["fine"]["_ts"].ts.cpu().numpy()
    g, ld, _ = T.grads(oc, of, params, BBOX_MIN, BBOX_MAX, batch, uc, uf, 64, 128, fixed_fine_ts=fine_ts,
                       dtype=torch.float64)
    g32, _, _ = T.grads(oc, of, params, BBOX_MIN, BBOX_MAX, batch, uc, uf, 64, 128, fixed_fine_ts=fine_ts)
    logs = step((dev(uc), dev(uf)), dev(batch))
    assert set(logs) == {"coarse", "fine", "coarse_normal_mse", "coarse_neg_normal", "fine_normal_mse",
                         "fine_neg_normal", "grad_norm", "param_norm"}
    for k in ("coarse", "fine", "coarse_normal_mse", "coarse_neg_normal", "fine_normal_mse", "fine_neg_normal"):
        np.testing.assert_allclose(float(logs[k]), ld[k], rtol=5e-3, atol=1e-6, err_msg=k)
    grads = loop._grads
    from oracle.models_torch import tree_leaves
    worst = []
    for name, model in (("coarse", nc), ("fine", nf)):
        gt = model.bind(grads[loop._slices[name][0]:loop._slices[name][1]])
        for path, ref in tree_leaves(g[name]):
            node, node32 = gt, g32[name]
            for part in path.split("/"):
                node, node32 = node[part], node32[part]
            if float(ref.abs().max()) == 0.0:
                continue
            worst.append((rel_l2(node.cpu().numpy(), ref.numpy()), rel_l2(node32.numpy(), ref.numpy()), name, path))
    worst.sort(reverse=True)
    print("worst NGP-Ref train grad rel-L2 (gpu, cpu-fp32):", worst[:4])
    assert worst[0][0] < 2e-2, worst[:4]
    assert worst[0][0] < 3 * max(w[1] for w in worst) + 1e-5, worst[:4]


def test_golden_ngpref_fixture_gpu():
    """The CUDA InstantNGPRefNERFModel path against the committed fixture tests/golden/ngpref_small.npz."""
    import importlib.util
    import os
    from learn_nerf.instant_ngp import InstantNGPRefNERFModel
    golden = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(os.path.join(golden, "ngpref_small.npz"))
    x, d, o, p = mg.ngpref_case()
    n = InstantNGPRefNERFModel(table_sizes=[2 ** 14] * 16, grid_sizes=[2 ** (4 + i // 2) for i in range(16)],
                               bbox_min=BBOX_MIN, bbox_max=BBOX_MAX)
    dens, rgb, aux = n.apply(dict(params=to_native(n, p)), dev(x), dev(d))
    np.testing.assert_allclose(dens.cpu().numpy(), g["dens"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(rgb.cpu().numpy(), g["rgb"], atol=2e-5)
    np.testing.assert_allclose(aux["neg_normal"].cpu().numpy(), g["neg_normal"], atol=2e-5)
    assert np.median(np.abs(aux["normal_mse"].cpu().numpy() - g["normal_mse"])) < 2e-5
