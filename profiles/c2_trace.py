"""Per-layer timeline of the CTA-pair forward (profiling build: LNRF_EXTRA_NVCC_FLAGS=-DLNRF_C2_TRACE
python learn-nerf_b200/build.py --force).  Prints, for cluster 0's leader CTA and tile iterations 0..2,
the MMA issuer's wait for a_ready / full barriers and each epilogue team's wait for the accumulator and
its epilogue duration, in SM clocks."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from learn_nerf import _native
from learn_nerf.model import NeRFModel
torch.cuda.set_device(0)
m = NeRFModel(precision="bf16")
tree = m.init(0, device="cuda")["params"]
n, T = 4096, 192
rays = torch.randn(n, 2, 3, device="cuda")
ts = torch.rand(n, T, device="cuda").sort(dim=1).values + 2
save = bool(int(os.environ.get("SAVE", "0")))
for _ in range(3):
    m.apply_rays(tree, rays, ts, save=save, slot="a")
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 8192)()
lib = _native.load()
rc = lib.lnrf_debug_c2_trace(buf, 8192)
a = np.frombuffer(buf, dtype=np.uint64).astype(np.int64)
t00 = a[0]
print("save", save, "rc", rc)
print("MMA issuer: (t,L,g) a_ready-wait  issue-span  full-wait   start(rel)")
for t in range(2):
    for L in range(10):
        for g in range(2):
            b = ((t * 10 + L) * 2 + g) * 4
            T0, T1, T2, fw = a[b:b + 4]
            print(f"  t{t} L{L} g{g}: a_wait {T1 - T0:6d}  issue {T2 - T1:6d}  full_wait {fw:6d}   @ {T0 - t00:8d}")
print("epilogue teams: (t,TL,team) acc-wait | stash-read wait+bar | tmem->smem | bar+bulk issue | arrive | start(rel)")
for t in range(2):
    for TL in range(9):
        for team in range(4):
            b = 1024 + ((t * 10 + TL) * 4 + team) * 6
            E0, E1, Ea, Eb, Ed, E2 = a[b:b + 6]
            print(f"  t{t} TL{TL} team{team}: acc_wait {E1 - E0:6d} | {Ea - E1:5d} | {Eb - Ea:5d} | {Ed - Eb:5d} | {E2 - Ed:4d} |  @ {E0 - t00:8d}")
