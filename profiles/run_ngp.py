"""Launch the Instant-NGP forward/backward kernels a few times (profiling helper)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from learn_nerf.instant_ngp import InstantNGPModel
torch.cuda.set_device(0)
L = int(os.environ.get("L", "16"))
m = InstantNGPModel(table_sizes=[2 ** 18] * L, grid_sizes=[2 ** (4 + i // 2) for i in range(L)],
                    bbox_min=[-1.0] * 3, bbox_max=[1.0] * 3, precision=os.environ.get("PREC", "bf16"))
tree = m.init(0, device="cuda")["params"]
n, T = int(os.environ.get("N", "8192")), int(os.environ.get("T", "192"))
o = torch.randn(n, 3, device="cuda"); o = 4 * o / o.norm(dim=1, keepdim=True)
tgt = torch.rand(n, 3, device="cuda") * 2 - 1
d = tgt - o; d = d / d.norm(dim=1, keepdim=True)
rays = torch.stack([o, d], dim=1).contiguous()
ts = torch.rand(n, T, device="cuda").sort(dim=1).values * 2 + 3
g = torch.zeros_like(tree.flat)
for _ in range(3):
    dens, rgb, _, ctx = m.apply_rays(tree, rays, ts, save=True, slot="a")
    m.backward_rays(ctx, torch.randn_like(dens) * 1e-3, torch.randn_like(rgb) * 1e-3, g)
torch.cuda.synchronize()
print("ok")
