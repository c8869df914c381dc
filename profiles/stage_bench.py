"""Achieved GB/s of the HBM-bound stages (K1, K3, K4, K5, K7, K8, K10) against the measured HBM copy
peak, each kernel timed alone with CUDA events (L2 flushed between repetitions) at a size that does
not fit in L2.  Algorithmic bytes per unit are SURVEY.md 8(d)'s.

  python profiles/stage_bench.py [N_RAYS]      -> one JSON line per stage
  bench.py imports run_stages() for the `hbm_stages` block of its JSON line.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run_stages(n=262144, dev=None, peak=None, reps=5):
    import torch
    from learn_nerf import _native
    from learn_nerf.instant_ngp import InstantNGPModel
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    if peak is None:
        path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak = json.load(open(path))["hbm_gbs"] if os.path.exists(path) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []

    def timeit(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    def report(name, kernel, units, ms, nbytes, note=None):
        gbs = nbytes / (ms * 1e-3) / 1e9
        row = {"stage": name, "kernel": kernel, "units": int(units), "ms": round(ms, 4), "algorithmic_bytes": int(nbytes),
               "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 3)}
        if note:
            row["note"] = note
        out.append(row)

    g = torch.Generator(device=dev).manual_seed(0)
    o = torch.randn(n, 3, device=dev, generator=g); o = 4 * o / o.norm(dim=1, keepdim=True)
    tgt = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
    d = tgt - o; d = d / d.norm(dim=1, keepdim=True)
    rays = torch.stack([o, d], dim=1).contiguous()
    uc = torch.rand(n, 64, device=dev, generator=g)
    uf = torch.rand(n, 128, device=dev, generator=g)
    lo, hi = [-1.0] * 3, [1.0] * 3
    t_min, t_max, mask, ts_c = _native.sample_coarse(rays, lo, hi, uc)
    report("K1 t_range + stratified sampling", "sample_coarse_kernel", n,
           timeit(lambda: _native.sample_coarse(rays, lo, hi, uc)), 545 * n)
    dens_c = torch.rand(n, 64, device=dev, generator=g) * 3
    rgb_c = torch.rand(n, 64, 3, device=dev, generator=g) * 2 - 1
    bg = torch.tensor([-1.0, -1.0, -1.0], device=dev)
    report("K3 composite fwd T=64", "composite_fwd_pf_kernel", n,
           timeit(lambda: _native.composite_fwd(rays, ts_c, t_min, t_max, mask, dens_c, rgb_c, bg)), 1332 * n)
    ts_f = _native.sample_fine(ts_c, dens_c, t_min, t_max, uf)
    report("K4 fine sampling 64+128", "sample_fine64_kernel", n,
           timeit(lambda: _native.sample_fine(ts_c, dens_c, t_min, t_max, uf)), 1792 * n,
           "bit-exact sequential cumsums + two 25-op exp per sample: issue-bound")
    dens_f = torch.rand(n, 192, device=dev, generator=g) * 3
    rgb_f = torch.rand(n, 192, 3, device=dev, generator=g) * 2 - 1
    report("K3 composite fwd T=192", "composite_fwd_pf_kernel", n,
           timeit(lambda: _native.composite_fwd(rays, ts_f, t_min, t_max, mask, dens_f, rgb_f, bg)), 3892 * n)
    d_out = torch.randn(n, 3, device=dev, generator=g)
    d_bg = torch.zeros(3, device=dev)
    report("K5 composite bwd T=192", "composite_bwd_pf_kernel", n,
           timeit(lambda: _native.composite_bwd(ts_f, t_min, t_max, mask, dens_f, rgb_f, bg, d_out, d_bg)), 6948 * n)
    cnt = 7_653_929 // 4 * 4
    p, gr, m, v = (torch.randn(cnt, device=dev, generator=g) for _ in range(4))
    v.abs_()
    norms = torch.zeros(2, device=dev)
    report("K10 Adam + norms (NGP-sized, 7.65 M params)", "adam_kernel", cnt,
           timeit(lambda: _native.adam_step(p, gr, m, v, 1e-4, 0.9, 0.999, 1e-7, 1, 1.0, norms)), 28 * cnt)
    # hash grid (fine model, L = 16) on n/8 rays x 192 samples
    L = 16
    ngp = InstantNGPModel(table_sizes=[2 ** 18] * L, grid_sizes=[2 ** (4 + i // 2) for i in range(L)], bbox_min=lo,
                          bbox_max=hi)
    tree = ngp.init(0, device=dev)["params"]
    nr = max(n // 8, 1)
    m_pts = nr * 192
    enc = torch.empty(m_pts, 2 * L, device=dev)
    r8, t8 = rays[:nr].contiguous(), ts_f[:nr].contiguous()
    report("K7 hash-grid gather L=16", "hashgrid_fwd_kernel", m_pts,
           timeit(lambda: _native.hashgrid_fwd(tree.flat, ngp.spec(), None, r8, t8, nr, 192, enc)), m_pts * (L * 72 + 12),
           "tables (25.8 MB) are L2-resident: gathers are served by L2")
    d_enc = torch.randn(m_pts, 2 * L, device=dev, generator=g)
    gt = torch.zeros_like(tree.flat)
    report("K8 hash-grid scatter-add L=16", "hashgrid_bwd_kernel", m_pts,
           timeit(lambda: _native.hashgrid_bwd(ngp.spec(), None, r8, t8, nr, 192, d_enc, gt)), m_pts * (L * 136 + 12),
           "float2 atomics into L2-resident tables (RMW counted twice)")
    return out


if __name__ == "__main__":
    import torch
    torch.cuda.set_device(0)
    for row in run_stages(int(sys.argv[1]) if len(sys.argv) > 1 else 262144):
        print(json.dumps(row))
