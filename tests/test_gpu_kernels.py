"""GPU parity tests, stage by stage, CUDA path (through the C ABI) vs the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import BBOX_MAX, BBOX_MIN, F, make_rays, make_uniforms

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def nat():
    from learn_nerf import _native
    _native.ensure_init(torch.device("cuda", 0))
    return _native


# ------------------------------------------------------------------ tcgen05 building block
@pytest.mark.parametrize("N,K", [(256, 64), (256, 256), (144, 128), (16, 64), (128, 192)])
def test_umma_debug_gemm(nat, N, K):
    """Pins the UMMA smem/instruction descriptor encodings and the TMEM read-back layout."""
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + K)
    a = torch.randn(128, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g)
    out = nat.debug_umma_gemm(a, b)
    torch.cuda.synchronize()
    ref = a.bfloat16().float() @ b.bfloat16().float().T
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("M,N", [(128, 64), (128, 256), (256, 256), (256, 192), (256, 128)])
def test_umma_debug_gemm_tn(nat, M, N):
    """MN-major (transposed) operand descriptors used by the dW = act^T @ grad GEMMs."""
    g = torch.Generator(device="cuda").manual_seed(M * 1000 + N)
    at = torch.randn(128, M, device="cuda", generator=g)
    bt = torch.randn(128, N, device="cuda", generator=g)
    out = nat.debug_umma_gemm_tn(at, bt)
    torch.cuda.synchronize()
    ref = at.bfloat16().float().T @ bt.bfloat16().float()
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


# ------------------------------------------------------------------ K1
def _edge_rays():
    r = make_rays(512, seed=3, miss_frac=0.3, with_targets=False)
    r[0] = [[0, 0, -3], [0, 0, 1]]           # axis aligned: d == 0 on two axes
    r[1] = [[0, 5, -3], [0, 0, 1]]           # misses the box
    r[2] = [[0.5, 0.5, 0.5], [1, 0, 0]]      # starts inside
    r[3] = [[-1, -1, -3], [0, 0, 1]]         # grazes an edge
    r[4] = [[0, 0, -3], [0, 0, -1]]          # points away
    r[5] = [[0, 0, -3], [-1e-8, 0, 1]]       # d + eps == 0 on x
    return r


def test_sample_coarse_bit_exact(nat):
    from oracle import render_np
    rays = _edge_rays()
    n = len(rays)
    u = make_uniforms(n, 64, 7)
    t_min, t_max, mask, ts = nat.sample_coarse(dev(rays), BBOX_MIN, BBOX_MAX, dev(u))
    o_min, o_max, o_mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
    o = render_np.RaySamples.stratified_sampling(o_min, o_max, o_mask, 64, u)
    np.testing.assert_array_equal(mask.cpu().numpy().astype(bool), o_mask)
    np.testing.assert_array_equal(t_min.cpu().numpy().view(np.uint32), o_min.view(np.uint32))
    np.testing.assert_array_equal(t_max.cpu().numpy().view(np.uint32), o_max.view(np.uint32))
    np.testing.assert_array_equal(ts.cpu().numpy().view(np.uint32), o.ts.view(np.uint32))
    # standalone stratified kernel, odd T
    u2 = make_uniforms(n, 37, 8)
    ts2 = nat.stratified(t_min, t_max, dev(u2))
    o2 = render_np.RaySamples.stratified_sampling(o_min, o_max, o_mask, 37, u2)
    np.testing.assert_array_equal(ts2.cpu().numpy().view(np.uint32), o2.ts.view(np.uint32))


def test_sample_coarse_empty_and_errors(nat):
    e = torch.empty(0, 2, 3, device="cuda")
    t_min, t_max, mask, ts = nat.sample_coarse(e, BBOX_MIN, BBOX_MAX, torch.empty(0, 64, device="cuda"))
    assert ts.shape == (0, 64)
    with pytest.raises(nat.LnrfError):
        nat.sample_coarse(torch.zeros(2, 2, 3), BBOX_MIN, BBOX_MAX, torch.zeros(2, 4))  # CPU tensors


# ------------------------------------------------------------------ K4
def test_sample_fine_golden_bit_exact(nat):
    g = np.load(os.path.join(GOLDEN, "fine_sampling_small.npz"))
    out, idx, new_ts = nat.sample_fine(dev(g["ts"]), dev(g["densities"]), dev(g["t_min"]),
                                       dev(g["t_max"]), dev(g["u"]), debug=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), g["idx"])
    np.testing.assert_array_equal(new_ts.cpu().numpy().view(np.uint32), g["new_ts"].view(np.uint32))
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), g["fine_ts"].view(np.uint32))


@pytest.mark.parametrize("Tc,Tf", [(64, 128), (16, 48), (33, 17)])
def test_sample_fine_random_bit_exact(nat, Tc, Tf):
    from oracle import render_np
    n = 300
    rays = make_rays(n, seed=5, miss_frac=0.2, with_targets=False)
    t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
    cs = render_np.RaySamples.stratified_sampling(t_min, t_max, mask, Tc, make_uniforms(n, Tc, 9))
    rs = np.random.RandomState(4)
    dens = (rs.gamma(0.3, 10.0, (n, Tc)) * (rs.uniform(size=(n, Tc)) < 0.5)).astype(F)
    u = make_uniforms(n, Tf, 10)
    o, o_idx = cs.fine_sampling(Tf, u, dens, return_indices=True)
    out, idx, _ = nat.sample_fine(dev(cs.ts), dev(dens), dev(t_min), dev(t_max), dev(u), debug=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), o_idx)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), o.ts.view(np.uint32))
    assert bool((out[:, 1:] >= out[:, :-1]).all())


def test_sample_fine_unsorted_inputs_take_the_sort_path(nat):
    """The 64+128 kernel merges two sorted lists; rays whose coarse samples are NOT sorted (possible
    through fp32 rounding, forced here) must fall back to the full sort and still equal np.sort."""
    from oracle import render_np
    n = 64
    rays = make_rays(n, seed=15, with_targets=False)
    t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
    cs = render_np.RaySamples.stratified_sampling(t_min, t_max, mask, 64, make_uniforms(n, 64, 19))
    rs = np.random.RandomState(14)
    ts = cs.ts.copy()
    for r in range(0, n, 2):  # every other ray: one or three 1-ulp inversions, as rounding could produce
        for i in rs.choice(63, 3 if r % 4 == 0 else 1, replace=False):
            ts[r, i + 1] = np.nextafter(ts[r, i], np.float32(-np.inf))
    assert (np.diff(ts, axis=1) < 0).any()
    cs2 = render_np.RaySamples(t_min=t_min, t_max=t_max, mask=mask, ts=ts)
    dens = rs.gamma(0.5, 4.0, (n, 64)).astype(F)
    u = make_uniforms(n, 128, 20)
    o, o_idx = cs2.fine_sampling(128, u, dens, return_indices=True)
    out, idx, _ = nat.sample_fine(dev(ts), dev(dens), dev(t_min), dev(t_max), dev(u), debug=True)
    np.testing.assert_array_equal(idx.cpu().numpy(), o_idx)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), o.ts.view(np.uint32))


# ------------------------------------------------------------------ K3 / K5
def _composite_case(n, T, seed, miss=0.25):
    from oracle import render_np
    rays = make_rays(n, seed=seed, miss_frac=miss, with_targets=False)
    t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
    s = render_np.RaySamples.stratified_sampling(t_min, t_max, mask, T, make_uniforms(n, T, seed + 1))
    rs = np.random.RandomState(seed + 2)
    dens = rs.gamma(0.7, 3.0, (n, T)).astype(F)
    rgb = rs.uniform(-1, 1, (n, T, 3)).astype(F)
    bg = np.array([-1.0, 0.3, 0.8], F)
    return rays, s, dens, rgb, bg


@pytest.mark.parametrize("T", [64, 192, 7, 100])
def test_composite_fwd(nat, T):
    rays, s, dens, rgb, bg = _composite_case(257, T, 20 + T)
    out, alphas, coords = nat.composite_fwd(dev(rays), dev(s.ts), dev(s.t_min), dev(s.t_max),
                                            dev(s.mask.astype(np.uint8)), dev(dens), dev(rgb), dev(bg))
    np.testing.assert_allclose(out.cpu().numpy(), s.render_rays(dens, rgb, bg), atol=1e-5)
    np.testing.assert_allclose(alphas.cpu().numpy(), s.render_alpha(dens), atol=1e-5)
    np.testing.assert_allclose(coords.cpu().numpy(),
                               s.render_rays(dens, s.points(rays), np.zeros(3, F)), atol=2e-5)


@pytest.mark.parametrize("T", [64, 192, 5])
def test_composite_bwd_vs_autograd(nat, T):
    from oracle.render_torch import composite
    rays, s, dens, rgb, bg = _composite_case(130, T, 40 + T)
    rs = np.random.RandomState(1)
    d_out = rs.randn(130, 3).astype(F)
    td = torch.from_numpy(dens).double().requires_grad_(True)
    tc = torch.from_numpy(rgb).double().requires_grad_(True)
    tb = torch.from_numpy(bg).double().requires_grad_(True)
    out = composite(torch.from_numpy(s.ts).double(), torch.from_numpy(s.t_min).double(),
                    torch.from_numpy(s.t_max).double(), torch.from_numpy(s.mask), td, tc, tb)
    (out * torch.from_numpy(d_out).double()).sum().backward()
    d_bg = torch.zeros(3, device="cuda")
    d_dens, d_rgb = nat.composite_bwd(dev(s.ts), dev(s.t_min), dev(s.t_max),
                                      dev(s.mask.astype(np.uint8)), dev(dens), dev(rgb), dev(bg),
                                      dev(d_out), d_bg)
    def rel(a, b):
        return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
    assert rel(d_dens.cpu().numpy(), td.grad.numpy()) < 1e-4   # stated tolerance: rel-L2 1e-4 (fp32)
    assert rel(d_rgb.cpu().numpy(), tc.grad.numpy()) < 1e-4
    assert rel(d_bg.cpu().numpy(), tb.grad.numpy()) < 1e-4


def test_mse_loss(nat):
    rs = np.random.RandomState(0)
    batch = rs.randn(100, 3, 3).astype(F)
    out = rs.randn(100, 3).astype(F)
    tb = dev(batch)
    loss = torch.zeros(1, device="cuda")
    d_out = torch.empty(100, 3, device="cuda")
    nat.mse_loss(dev(out), tb[:, 2], 9, 100, 1.0 / 300, loss, d_out)
    ref = ((out - batch[:, 2]) ** 2).sum()
    np.testing.assert_allclose(loss.item(), ref, rtol=1e-5)
    np.testing.assert_allclose(d_out.cpu().numpy(), 2 * (out - batch[:, 2]) / 300, rtol=1e-6, atol=1e-9)


# ------------------------------------------------------------------ K10
def test_adam_step(nat):
    from oracle import train_torch as T
    rs = np.random.RandomState(0)
    n = 10007
    p0 = rs.randn(n).astype(F)
    params = dict(w=torch.from_numpy(p0.copy()))
    st = T.AdamState(params)
    p, m, v = dev(p0.copy()), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        g = rs.randn(n).astype(F)
        norms = torch.zeros(2, device="cuda")
        p_before = p.cpu().numpy().copy()
        nat.adam_step(p, dev(g * 2.0), m, v, 1e-3, 0.9, 0.999, 1e-7, step, 0.5, norms)
        params = T.adam_update(params, dict(w=torch.from_numpy(g)), st, 1e-3, eps=1e-7)
        np.testing.assert_allclose(p.cpu().numpy(), params["w"].numpy(), rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(norms.cpu().numpy(),
                                   [(g.astype(np.float64) ** 2).sum(), (p_before.astype(np.float64) ** 2).sum()],
                                   rtol=1e-5)


def test_adam_step_peers_single_process(nat):
    """lnrf_adam_step_peers (fused peer all-reduce + Adam) with the 'peers' being three gradient
    buffers of this one GPU: rank-order sum, 1/world scaling, the trailing loss sums, norms and the
    update against the oracle's optax.adam restatement fed with the mean gradient.  (The real
    multi-GPU exchange is covered by tests/mgpu_check.py under torchrun.)"""
    from oracle import train_torch as T
    rs = np.random.RandomState(1)
    n, extra, world = 10007, 2, 3
    p0 = rs.randn(n).astype(F)
    params = dict(w=torch.from_numpy(p0.copy()))
    st = T.AdamState(params)
    p, m, v = dev(p0.copy()), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    # 16-byte aligned per-rank buffers of n + extra (+ pad) floats
    bufs = [torch.zeros(n + extra + 3, device="cuda") for _ in range(world)]
    for step in range(1, 4):
        gs = [rs.randn(n + extra).astype(F) for _ in range(world)]
        for b, g in zip(bufs, gs):
            b[: n + extra].copy_(dev(g))
        norms, sums = torch.zeros(2, device="cuda"), torch.zeros(extra, device="cuda")
        p_before = p.cpu().numpy().copy()
        nat.adam_step_peers(p, [b.data_ptr() for b in bufs], m, v, n, extra, 1e-3, 0.9, 0.999, 1e-7, step,
                            1.0 / world, norms, sums)
        gsum = gs[0].astype(np.float32).copy()
        for g in gs[1:]:
            gsum = gsum + g  # rank order, fp32
        gmean = gsum[:n] * np.float32(1.0 / world)
        params = T.adam_update(params, dict(w=torch.from_numpy(gmean.copy())), st, 1e-3, eps=1e-7)
        np.testing.assert_allclose(p.cpu().numpy(), params["w"].numpy(), rtol=2e-6, atol=1e-7)
        np.testing.assert_array_equal(sums.cpu().numpy(), gsum[n:])
        np.testing.assert_allclose(norms.cpu().numpy(),
                                   [(gmean.astype(np.float64) ** 2).sum(), (p_before.astype(np.float64) ** 2).sum()],
                                   rtol=1e-5)


def _camera():
    import math
    return dict(camera_direction=(0.1, -0.2, -0.97), camera_origin=(0.5, 1.0, 4.0), x_axis=(0.99, 0.05, 0.09),
                y_axis=(0.04, -0.98, 0.2), x_fov=math.radians(60.0), y_fov=math.radians(45.0))


@pytest.mark.parametrize("w,h", [(1, 1), (7, 5), (128, 128), (801, 333)])
def test_bare_rays_bit_exact(w, h):
    """CameraView.bare_rays (dataset.py:52-78) on the device vs the oracle restatement."""
    from learn_nerf.dataset import CameraView
    from oracle import render_np
    cam = _camera()
    ref = render_np.bare_rays(width=w, height=h, **cam)
    view = CameraView(**cam)
    got = view.bare_rays(w, h).cpu().numpy()
    np.testing.assert_array_equal(got, ref)
    # a row block equals the corresponding slice (how a view is sharded over GPUs)
    if h >= 3:
        part = view.bare_rays(w, h, row0=1, rows=h - 2).cpu().numpy()
        np.testing.assert_array_equal(part, ref[w:(h - 1) * w])


def test_rgb_to_u8_and_render_view():
    from learn_nerf import _native
    from learn_nerf.dataset import CameraView
    from learn_nerf.model import NeRFModel
    from learn_nerf.render import NeRFRenderer
    from learn_nerf.scripts.render_nerf import render_view
    from oracle import render_np
    rs = np.random.RandomState(0)
    c = np.concatenate([rs.uniform(-1.2, 1.2, 1000), [-1.0, 1.0, 0.0, -0.0, 0.999999]]).astype(F)
    np.testing.assert_array_equal(_native.rgb_to_u8(dev(c)).cpu().numpy(), render_np.rgb_to_u8(c))
    model = NeRFModel(precision="bf16")
    params = model.init(3, device="cuda")["params"]
    r = NeRFRenderer(coarse=model, fine=model, coarse_params=params, fine_params=params,
                     background=torch.tensor([-1.0, -1.0, -1.0], device="cuda"), bbox_min=BBOX_MIN,
                     bbox_max=BBOX_MAX, coarse_ts=64, fine_ts=128)
    view = CameraView(camera_direction=(0.0, 0.0, -1.0), camera_origin=(0.0, 0.0, 4.0), x_axis=(1.0, 0.0, 0.0),
                      y_axis=(0.0, -1.0, 0.0), x_fov=1.0, y_fov=1.0)
    img = render_view(r, view, 40, 30, batch_size=512, key=1)
    assert img.shape == (30, 40, 3) and img.dtype == torch.uint8
    corner = img[0, 0].cpu().numpy()  # fov 1 rad at distance 4: the corner ray misses the unit bbox -> background
    np.testing.assert_array_equal(corner, [0, 0, 0])
    assert int(img[15, 20].sum()) > 0


def test_render_view_cuda_graph_matches_eager():
    """render_view(cuda_graph=True): chunks replayed from one captured render_rays give the same
    image as the eager loop (same keys -> same uniforms -> identical uint8 pixels), including a
    ragged last chunk, on the bf16 path."""
    from learn_nerf.dataset import CameraView
    from learn_nerf.model import NeRFModel
    from learn_nerf.render import NeRFRenderer
    from learn_nerf.scripts.render_nerf import render_view
    coarse, fine = NeRFModel(precision="bf16"), NeRFModel(precision="bf16")
    r = NeRFRenderer(coarse=coarse, fine=fine, coarse_params=coarse.init(1, device="cuda")["params"],
                     fine_params=fine.init(2, device="cuda")["params"], background=torch.tensor([-1.0, 0.0, 1.0]).cuda(),
                     bbox_min=BBOX_MIN, bbox_max=BBOX_MAX, coarse_ts=64, fine_ts=128)
    view = CameraView(**_camera())
    eager = render_view(r, view, 50, 41, batch_size=512, key=9)          # 2050 rays: 4 chunks + 2 rays
    graph = render_view(r, view, 50, 41, batch_size=512, key=9, cuda_graph=True)
    again = render_view(r, view, 50, 41, batch_size=512, key=9, cuda_graph=True)  # cached graph
    assert eager.shape == (41, 50, 3) and eager.dtype == torch.uint8
    assert torch.equal(eager, graph) and torch.equal(graph, again)


@pytest.mark.parametrize("shape", [(1,), (5, 7), (4096, 64), (333, 128)])
def test_threefry_uniform_bit_exact(shape):
    """Device uniforms from a key equal the numpy threefry oracle bit for bit."""
    from learn_nerf import prng
    from oracle import prng_np
    key = prng.split(11)[1]
    got = prng.uniform(key, shape, "cuda").cpu().numpy()
    ref = prng_np.uniform(np.array([key.k0, key.k1], np.uint32), shape)
    np.testing.assert_array_equal(got, ref)


def test_dataset_feeder_gpu_prefetch(tmp_path):
    """load_dataset (PNG + JSON views, dataset.py:266-288) -> ShuffledDataset -> batches delivered
    to the GPU through pinned staging buffers on a side stream: same batches, same order as the
    host path; colours are the premultiplied-alpha image mapped to [-1, 1]."""
    import json
    import math
    from PIL import Image
    from learn_nerf.dataset import load_dataset
    rs = np.random.RandomState(0)
    d = tmp_path / "scene"
    d.mkdir()
    (d / "metadata.json").write_text(json.dumps({"min": [-1, -1, -1], "max": [1, 1, 1]}))
    imgs = {}
    for i in range(3):
        rgba = rs.randint(0, 256, (12, 16, 4)).astype(np.uint8)
        Image.fromarray(rgba, "RGBA").save(d / f"v{i}.png")
        imgs[f"v{i}"] = rgba
        ang = 2 * math.pi * i / 3
        (d / f"v{i}.json").write_text(json.dumps(dict(
            origin=[4 * math.cos(ang), 4 * math.sin(ang), 0.5], z=[-math.cos(ang), -math.sin(ang), 0.0],
            x=[-math.sin(ang), math.cos(ang), 0.0], y=[0.0, 0.0, -1.0], x_fov=1.0, y_fov=0.8)))
    ds = load_dataset(str(d))
    assert len(ds.views) == 3 and ds.metadata.bbox_max == (1, 1, 1)
    v0 = ds.views[0]
    rays = v0.rays().cpu().numpy()
    assert rays.shape == (12 * 16, 3, 3)
    rgba = imgs["v0"].astype(np.float64)
    want = np.round(rgba[:, :, :3] * (rgba[:, :, 3:] / 255)).astype(np.uint8).reshape(-1, 3)
    np.testing.assert_array_equal(rays[:, 2], want.astype(F) / F(127.5) - F(1.0))
    np.testing.assert_array_equal(rays[:, :2], v0.bare_rays(16, 12).cpu().numpy())
    host = list(ds.iterate_batches(str(tmp_path / "s"), 7, batch_size=100, repeat=False))
    gpu = list(ds.iterate_batches(str(tmp_path / "s"), 7, batch_size=100, repeat=False, device="cuda"))
    assert len(host) == len(gpu) == 6 and gpu[0].is_cuda and gpu[-1].shape[0] == 3 * 192 - 500
    for a, b in zip(host, gpu):
        assert torch.equal(a, b.cpu())


def test_train_script_mirror_end_to_end(tmp_path):
    """scripts/train_nerf.train (train_nerf.py:72-133): dataset directory -> shuffled feeder with GPU
    prefetch -> TrainLoop.step_fn -> checkpoint, for two of create_model's four families; the loss
    on a constant-colour scene falls."""
    import json
    import math
    from PIL import Image
    from learn_nerf.dataset import load_dataset
    from learn_nerf.scripts.train_nerf import create_model, train
    d = tmp_path / "scene"
    d.mkdir()
    (d / "metadata.json").write_text(json.dumps({"min": [-1, -1, -1], "max": [1, 1, 1]}))
    for i in range(4):
        Image.fromarray(np.full((24, 24, 4), 255, np.uint8) * np.array([1, 0, 0, 1], np.uint8), "RGBA").save(d / f"v{i}.png")
        ang = 2 * math.pi * i / 4
        (d / f"v{i}.json").write_text(json.dumps(dict(
            origin=[4 * math.cos(ang), 4 * math.sin(ang), 0.0], z=[-math.cos(ang), -math.sin(ang), 0.0],
            x=[-math.sin(ang), math.cos(ang), 0.0], y=[0.0, 0.0, -1.0], x_fov=0.5, y_fov=0.5)))
    data = load_dataset(str(d))
    kinds = [type(m).__name__ for kw in (dict(), dict(ref_nerf=True), dict(instant_ngp=True),
                                        dict(instant_ngp=True, ref_nerf=True))
             for m in create_model(data.metadata, **kw)[:1]]
    assert kinds == ["NeRFModel", "RefNERFModel", "InstantNGPModel", "InstantNGPRefNERFModel"]
    for kw in (dict(precision="bf16"), dict(instant_ngp=True)):
        logs = []
        ck = str(tmp_path / ("ck_" + "_".join(kw) + ".pkl"))
        loop = train(data, ck, key=3, lr=2e-3, batch_size=512, max_steps=12, save_interval=5, log=logs.append, **kw)
        fine = [float(l.split(" fine=")[1].split()[0]) for l in logs if l.startswith("step")]
        assert len(fine) == 12 and fine[-1] < fine[0], fine
        assert os.path.exists(ck) and loop.state.step == 12


# ------------------------------------------------------------------ RaySamples method surface
@pytest.mark.parametrize("T", [64, 192, 7, 1])
def test_ray_samples_methods_bit_exact(nat, T):
    """RaySamples.starts / ends / deltas / termination_probs / average_aux_losses (render.py:192-209,
    259-287) through the host mirror against oracle.render_np: bit-exact (aux means: 1e-6)."""
    from learn_nerf.render import RaySamples
    rays, s, dens, rgb, bg = _composite_case(130, T, 300 + T)
    dens[0] = 0.0
    dens[1] = 1e6
    rs_ = RaySamples(t_min=dev(s.t_min), t_max=dev(s.t_max), mask=dev(s.mask), ts=dev(s.ts))
    for name in ("starts", "ends", "deltas"):
        got = getattr(rs_, name)().cpu().numpy()
        np.testing.assert_array_equal(got.view(np.uint32), getattr(s, name)().view(np.uint32), err_msg=name)
    probs = rs_.termination_probs(dev(dens)).cpu().numpy()
    want = s.termination_probs(dens)
    assert probs.shape == (130, T + 1)
    np.testing.assert_array_equal(probs.view(np.uint32), want.view(np.uint32))
    np.testing.assert_allclose(probs.sum(1), 1.0, atol=3e-6)
    aux = {"a": np.abs(rgb[..., 0]), "b": rgb[..., 1] ** 2}
    got = rs_.average_aux_losses(dev(dens), {k: dev(v) for k, v in aux.items()})
    ref = s.average_aux_losses(dens, aux)
    for k in aux:
        np.testing.assert_allclose(float(got[k]), float(ref[k]), atol=1e-6, rtol=1e-5)


def test_z_depth_bit_exact(nat):
    """lnrf_z_depth (render_new_dataset.py:100-117) against the oracle restatement."""
    from oracle import render_np
    rs = np.random.RandomState(5)
    n = 1000
    coords = rs.uniform(-2, 2, (n, 3)).astype(F)
    alphas = rs.uniform(0, 1, (n, 1)).astype(F)
    alphas[:5, 0] = [0.9, 0.90000004, 1.0, 0.0, 0.95]
    origin, direction = (0.5, -3.0, 1.0), (0.0, 0.8, -0.6)
    z, d32 = nat.z_depth(dev(coords), dev(alphas), origin, direction, 4.0)
    o_z, o_d = render_np.z_depth(coords, alphas, origin, direction, 4.0)
    np.testing.assert_array_equal(z.cpu().numpy().view(np.uint32), o_z.view(np.uint32))
    np.testing.assert_array_equal(d32.cpu().numpy().view(np.uint32), o_d)


# ------------------------------------------------------------------ property tests (hypothesis)
try:
    from hypothesis import given, settings, strategies as st
    HAVE_HYPOTHESIS = True
except Exception:  # noqa: BLE001
    HAVE_HYPOTHESIS = False

if HAVE_HYPOTHESIS:
    _coord = st.floats(-6.0, 6.0, width=32, allow_nan=False)
    _dirc = st.sampled_from([0.0, 1.0, -1.0, 1e-8, -1e-8, 0.25, -0.7, 3e-5, 0.5, -2.0])
    _ray = st.tuples(st.tuples(_coord, _coord, _coord), st.tuples(_dirc, _dirc, _dirc))

    @settings(max_examples=40, deadline=None)
    @given(rays=st.lists(_ray, min_size=1, max_size=40), seed=st.integers(0, 2 ** 31 - 1), T=st.sampled_from([64, 37, 4]))
    def test_k1_bit_exact_property(rays, seed, T):
        """K1 on arbitrary origins and degenerate directions (zero components, +-1e-8 that cancel the
        epsilon, rays inside / grazing / missing the box): masks, bounds and positions bit-exact."""
        from learn_nerf import _native
        from oracle import render_np
        r = np.array(rays, F)
        n = len(r)
        u = (np.random.RandomState(seed).randint(0, 2 ** 23, (n, T)) * 2.0 ** -23).astype(F)
        t_min, t_max, mask, ts = _native.sample_coarse(dev(r), BBOX_MIN, BBOX_MAX, dev(u))
        o_min, o_max, o_mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, r)
        o = render_np.RaySamples.stratified_sampling(o_min, o_max, o_mask, T, u)
        np.testing.assert_array_equal(mask.cpu().numpy().astype(bool), o_mask)
        np.testing.assert_array_equal(t_min.cpu().numpy().view(np.uint32), o_min.view(np.uint32))
        np.testing.assert_array_equal(t_max.cpu().numpy().view(np.uint32), o_max.view(np.uint32))
        np.testing.assert_array_equal(ts.cpu().numpy().view(np.uint32), o.ts.view(np.uint32))

    @settings(max_examples=40, deadline=None)
    @given(seed=st.integers(0, 2 ** 31 - 1), kind=st.sampled_from(["gamma", "zero", "spike", "tiny", "huge", "sparse"]),
           shape=st.sampled_from([(64, 128), (16, 48), (33, 17)]), top=st.booleans())
    def test_k4_bit_exact_property(seed, kind, shape, top):
        """K4 over random densities incl. the jnp.interp edge cases (SURVEY 7 hard-part 3): flat CDF bins
        (zero / spike / sparse densities), the largest representable uniform, masked rays; bin indices
        and sorted positions bit-exact against the oracle."""
        from learn_nerf import _native
        from oracle import render_np
        Tc, Tf = shape
        rs = np.random.RandomState(seed)
        n = 24
        rays = make_rays(n, seed=seed % 997, miss_frac=0.25, with_targets=False)
        t_min, t_max, mask = render_np.ray_t_range(BBOX_MIN, BBOX_MAX, rays)
        u_c = (rs.randint(0, 2 ** 23, (n, Tc)) * 2.0 ** -23).astype(F)
        cs = render_np.RaySamples.stratified_sampling(t_min, t_max, mask, Tc, u_c)
        dens = {"gamma": lambda: rs.gamma(0.5, 4.0, (n, Tc)), "zero": lambda: np.zeros((n, Tc)),
                "spike": lambda: np.eye(Tc)[rs.randint(0, Tc, n)] * 1e6, "tiny": lambda: np.full((n, Tc), 1e-12),
                "huge": lambda: np.full((n, Tc), 1e5),
                "sparse": lambda: rs.gamma(0.3, 10.0, (n, Tc)) * (rs.uniform(size=(n, Tc)) < 0.2)}[kind]().astype(F)
        u = (rs.randint(0, 2 ** 23, (n, Tf)) * 2.0 ** -23).astype(F)
        if top:
            u[:, -1] = np.float32(1.0 - 2.0 ** -23)
        o, o_idx = cs.fine_sampling(Tf, u, dens, return_indices=True)
        out, idx, _ = _native.sample_fine(dev(cs.ts), dev(dens), dev(t_min), dev(t_max), dev(u), debug=True)
        np.testing.assert_array_equal(idx.cpu().numpy(), o_idx)
        np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), o.ts.view(np.uint32))
