"""Launch the split-fp16 tcgen05 GEMMs of the fp32-accurate paths on a fine-level shape (profiling helper):
rows NN 256 -> 256 with bias + ReLU, rows NT with mask, TN 256 x 256; prints CUDA-event times and GB/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from learn_nerf import _native
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
_native.ensure_init(dev)
M = int(os.environ.get("M", str(4096 * 192)))
N = 256
K = int(os.environ.get("K", "256"))
A = torch.randn(M, K, device=dev)
H = torch.relu(torch.randn(M, N, device=dev))
W = torch.randn(max(K, N), N, device=dev) / 16
bias = torch.zeros(N, device=dev)
C = torch.empty(M, N, device=dev)
dW = torch.zeros(N, N, device=dev)
db = torch.zeros(N, device=dev)
amax = torch.ones(1, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
p, lib, st = _native._p, _native.load(), _native._stream


bits = torch.zeros(((M + 31) // 32) * (N // 32) * 32, dtype=torch.int32, device=dev)


def rows(epi, mode):
    rc = lib.lnrf_tcgemm(mode, epi, M, N, p(A), K, K, None, 0, 0, p(W), N, p(C), N, p(bias), p(H), N, None, None, None,
                         p(amax), None, None, p(bits), p(bits) if epi == 0 else None, st())
    assert rc == 0, lib.lnrf_last_error()


def tn():
    rc = lib.lnrf_tcgemm(2, 0, M, N, p(H), N, N, None, 0, 0, p(C), N, p(dW), N, None, None, 0, None, None, p(db), None,
                         p(amax), None, None, None, st())
    assert rc == 0, lib.lnrf_last_error()


def timed(fn, name, nbytes):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s algorithmic ({nbytes / 1e9:.2f} GB)")


timed(lambda: rows(0, 0), "rows NN bias+relu", (K + N) * M * 4)
timed(lambda: rows(2, 1), "rows NT fp32 mask", 3 * M * N * 4)
timed(lambda: rows(6, 1), "rows NT bit mask ", 2 * M * N * 4)
timed(tn, "TN 256x256 + db  ", 2 * M * N * 4)
M_full = M
M = 4096   # one tile per CTA pair: the launch + weight-conversion prologue
timed(lambda: rows(0, 0), "rows NN, M = 4096 (fixed cost)", 2 * M * N * 4)
timed(lambda: rows(6, 1), "rows NT, M = 4096 (fixed cost)", 2 * M * N * 4)
print("ok")
