"""Ablation timing of the fused bf16 forward (profiling helper, not part of the product)."""
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
n, T = 4096, int(os.environ.get("T", "192"))
rays = torch.randn(n, 2, 3, device="cuda")
ts = torch.rand(n, T, device="cuda").sort(dim=1).values + 2
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("fwd(nosave) ms", timeit(lambda: m.apply_rays(tree, rays, ts, save=False)))
_native.load().lnrf_set_debug_flags(64)
print("fwd(nosave, N-half schedule) ms", timeit(lambda: m.apply_rays(tree, rays, ts, save=False)))
_native.load().lnrf_set_debug_flags(0)
for flags in [int(x) for x in os.environ.get("FLAGS", "0,8,16").split(",")]:
    _native.load().lnrf_set_debug_flags(flags)
    print(f"fwd(save) flags={flags} ms", timeit(lambda: m.apply_rays(tree, rays, ts, save=True, slot="a")))
_native.load().lnrf_set_debug_flags(0)
