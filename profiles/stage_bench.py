"""Achieved GB/s of the HBM-bound stages (K1, K3, K4, K5, K7, K8, K10) against the measured HBM copy
peak, each kernel timed alone with CUDA events at a size that does not fit in L2.

  python profiles/stage_bench.py [N_RAYS]      -> one JSON line per stage
Algorithmic bytes per unit are SURVEY.md 8(d)'s.
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from learn_nerf import _native

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def report(name, ms, nbytes, note=""):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"stage": name, "units": n, "ms": round(ms, 4), "algorithmic_bytes": int(nbytes),
                      "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3), "note": note}))


g = torch.Generator(device=dev).manual_seed(0)
o = torch.randn(n, 3, device=dev, generator=g); o = 4 * o / o.norm(dim=1, keepdim=True)
tgt = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
d = tgt - o; d = d / d.norm(dim=1, keepdim=True)
rays = torch.stack([o, d], dim=1).contiguous()
uc = torch.rand(n, 64, device=dev, generator=g)
uf = torch.rand(n, 128, device=dev, generator=g)
lo, hi = [-1.0] * 3, [1.0] * 3

t_min, t_max, mask, ts_c = _native.sample_coarse(rays, lo, hi, uc)
report("K1 sample_coarse (t_range + stratified)", timeit(lambda: _native.sample_coarse(rays, lo, hi, uc)), 545 * n)
dens_c = torch.rand(n, 64, device=dev, generator=g) * 3
rgb_c = torch.rand(n, 64, 3, device=dev, generator=g) * 2 - 1
bg = torch.tensor([-1.0, -1.0, -1.0], device=dev)
report("K3 composite_fwd T=64", timeit(lambda: _native.composite_fwd(rays, ts_c, t_min, t_max, mask, dens_c, rgb_c, bg)), 1332 * n)
ts_f = _native.sample_fine(ts_c, dens_c, t_min, t_max, uf)
report("K4 sample_fine 64+128", timeit(lambda: _native.sample_fine(ts_c, dens_c, t_min, t_max, uf)), 1792 * n)
dens_f = torch.rand(n, 192, device=dev, generator=g) * 3
rgb_f = torch.rand(n, 192, 3, device=dev, generator=g) * 2 - 1
report("K3 composite_fwd T=192", timeit(lambda: _native.composite_fwd(rays, ts_f, t_min, t_max, mask, dens_f, rgb_f, bg)), 3892 * n)
d_out = torch.randn(n, 3, device=dev, generator=g)
d_bg = torch.zeros(3, device=dev)
report("K5 composite_bwd T=192", timeit(lambda: _native.composite_bwd(ts_f, t_min, t_max, mask, dens_f, rgb_f, bg, d_out, d_bg)), 6948 * n)
cnt = 7_653_929 // 4 * 4
p, gr, m, v = (torch.randn(cnt, device=dev, generator=g) for _ in range(4))
v.abs_()
norms = torch.zeros(2, device=dev)
ms = timeit(lambda: _native.adam_step(p, gr, m, v, 1e-4, 0.9, 0.999, 1e-7, 1, 1.0, norms))
gbs = 28 * cnt / (ms * 1e-3) / 1e9
print(json.dumps({"stage": "K10 adam_step (NGP-sized, 7.65 M params; fits L2)", "units": cnt, "ms": round(ms, 4),
                  "algorithmic_bytes": 28 * cnt, "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3)}))
# hash grid (fine model, L = 16) on n/8 rays x 192 samples
from learn_nerf.instant_ngp import InstantNGPModel
L = 16
ngp = InstantNGPModel(table_sizes=[2 ** 18] * L, grid_sizes=[2 ** (4 + i // 2) for i in range(L)], bbox_min=lo, bbox_max=hi)
tree = ngp.init(0, device=dev)["params"]
nr = n // 8
m_pts = nr * 192
enc = torch.empty(m_pts, 2 * L, device=dev)
ms = timeit(lambda: _native.hashgrid_fwd(tree.flat, ngp.spec(), None, rays[:nr].contiguous(), ts_f[:nr].contiguous(), nr, 192, enc))
nb = m_pts * (L * 72 + 12)
print(json.dumps({"stage": "K7 hashgrid_fwd L=16", "units": m_pts, "ms": round(ms, 4), "algorithmic_bytes": nb,
                  "achieved_gbs": round(nb / ms / 1e6, 1), "peak_gbs": peak, "frac": round(nb / ms / 1e6 / peak, 3),
                  "note": "tables (25.8 MB) are L2-resident: gathers are served by L2"}))
d_enc = torch.randn(m_pts, 2 * L, device=dev, generator=g)
gt = torch.zeros_like(tree.flat)
ms = timeit(lambda: _native.hashgrid_bwd(ngp.spec(), None, rays[:nr].contiguous(), ts_f[:nr].contiguous(), nr, 192, d_enc, gt))
nb = m_pts * (L * 136 + 12)
print(json.dumps({"stage": "K8 hashgrid_bwd L=16", "units": m_pts, "ms": round(ms, 4), "algorithmic_bytes": nb,
                  "achieved_gbs": round(nb / ms / 1e6, 1), "peak_gbs": peak, "frac": round(nb / ms / 1e6 / peak, 3),
                  "note": "float2 atomics into L2-resident tables (RMW counted twice)"}))
