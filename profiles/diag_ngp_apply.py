"""Diagnostic: test_ngp_model_apply[6] inside a process that ran the GEMM-engine tests first (fresh-box flake hunt)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, pytest, torch
if os.environ.get("PRE", "1") == "1":
    pytest.main([os.path.join(ROOT, "tests/test_gpu_gemm_tc.py"), os.path.join(ROOT, "tests/test_gpu_ngp.py"), "-q", "-x",
                 "-k", "tcg or rows or tn or mask or layout or hashgrid or amax or strided"])
import test_gpu_ngp as T
F = np.float32
for levels in (6, 16):
    o, n = T.models(levels)
    p = T.oracle_params(o, 5)
    rs = np.random.RandomState(levels + 7)
    x = rs.uniform(-1, 1, (4097, 3)).astype(F)
    d = rs.randn(4097, 3).astype(F)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    outs = []
    with torch.no_grad():
        for _ in range(2):
            o_d, o_rgb, _ = o.apply(p, torch.from_numpy(x), torch.from_numpy(d))
            outs.append(o_rgb.numpy().copy())
    print(levels, "oracle run-to-run max diff", np.abs(outs[0] - outs[1]).max(), "threads", torch.get_num_threads())
    g = []
    for _ in range(3):
        dens, rgb, aux = n.apply(dict(params=T.to_native(n, p)), T.dev(x), T.dev(d))
        g.append(rgb.cpu().numpy().copy())
    print(levels, "gpu run-to-run max diff", np.abs(g[0] - g[1]).max(), np.abs(g[0] - g[2]).max())
    err = np.abs(g[0] - outs[0])
    bad = np.argwhere(err > 1e-5)
    print(levels, "gpu vs oracle max", err.max(), "violations", len(bad), "tiles of bad samples", np.unique(bad[:, 0] // 128)[:40])
    print(levels, "bad samples", bad[:30].tolist(), "per-channel max", err.max(0), "median err", np.median(err))
    print(levels, "dens err", np.abs(dens.cpu().numpy().reshape(-1) - o_d.numpy().reshape(-1)).max())
    with torch.no_grad():
        pd = T.M.tree_map(lambda t: t.double(), p) if hasattr(T, "M") else None
    from oracle import models_torch as M
    pd = M.tree_map(lambda t: t.double(), p)
    with torch.no_grad():
        _, r64, _ = o.apply(pd, torch.from_numpy(x).double(), torch.from_numpy(d).double())
    r64 = r64.numpy()
    print(levels, "vs fp64: gpu", np.abs(g[0] - r64).max(), "cpu32", np.abs(outs[0] - r64).max())
