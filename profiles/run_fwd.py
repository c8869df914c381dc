"""Launch the fused bf16 forward a few times (profiling helper)."""
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
rays = torch.randn(n, 2, 3, device="cuda")
ts = torch.rand(n, T, device="cuda").sort(dim=1).values + 2
save = bool(int(os.environ.get("SAVE", "0")))
for _ in range(3):
    out = m.apply_rays(tree, rays, ts, save=save, slot="a")
torch.cuda.synchronize()
if os.environ.get("BWD"):
    dd = torch.randn(n, T, device="cuda") * 1e-3
    dr = torch.randn(n, T, 3, device="cuda") * 1e-3
    g = torch.zeros_like(tree.flat)
    for _ in range(2):
        m.backward_rays(out[3], dd, dr, g)
    torch.cuda.synchronize()
print("ok")
