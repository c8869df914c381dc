"""Times the bf16 MLP kernels one by one on the fine level (4096 rays x 192 samples) with CUDA events
(median of 5, L2 flushed): forward without / with the stash, dX + dW backward.
  LNRF_TC_KERNELS=pair|cta2 LNRF_VERBOSE=1 python profiles/c2_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from learn_nerf import _native
from learn_nerf.model import NeRFModel
torch.cuda.set_device(0)
m = NeRFModel(precision="bf16")
tree = m.init(0, device="cuda")["params"]
n, T = int(os.environ.get("N", "4096")), int(os.environ.get("T", "192"))
g = torch.Generator(device="cuda").manual_seed(0)
rays = torch.randn(n, 2, 3, device="cuda", generator=g)
ts = torch.rand(n, T, device="cuda", generator=g).sort(dim=1).values + 2
dd = torch.randn(n, T, device="cuda", generator=g) * 1e-3
dr = torch.randn(n, T, 3, device="cuda", generator=g) * 1e-3
gr = torch.zeros_like(tree.flat)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    out = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return sorted(out)[len(out) // 2]


kind = os.environ.get("LNRF_TC_KERNELS", "cta2")
print(kind, "fwd nosave ms", round(timeit(lambda: m.apply_rays(tree, rays, ts, save=False)), 4))
print(kind, "fwd save   ms", round(timeit(lambda: m.apply_rays(tree, rays, ts, save=True, slot="a")), 4))
ctx = m.apply_rays(tree, rays, ts, save=True, slot="a")[3]
orig = _native.load().lnrf_nerf_mlp_bwd
print(kind, "bwd dX+dW  ms", round(timeit(lambda: m.backward_rays(ctx, dd, dr, gr)), 4))
