"""Where the gap between the device-timed step and the end-to-end step comes from (profiling helper)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
from bench import synth_batch
from learn_nerf.model import NeRFModel
from learn_nerf.train import TrainLoop
torch.cuda.set_device(0)
loop = TrainLoop(NeRFModel(precision="bf16"), NeRFModel(precision="bf16"), init_rng=2, lr=1e-4, coarse_ts=64, fine_ts=128)
step = loop.step_fn([-1.0] * 3, [1.0] * 3)
host = synth_batch(4096, 0).pin_memory()
dev = host.cuda()
for i in range(5):
    step(i, dev)
torch.cuda.synchronize()
def run(name, fn, reps=30):
    enq, tot = [], []
    for i in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(i)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        enq.append(t1 - t0); tot.append(t2 - t0)
    print(f"{name:46s} host enqueue {np.mean(enq)*1e3:6.3f} ms   total {np.mean(tot)*1e3:6.3f} ms")
run("device batch, no readback", lambda i: step(100 + i, dev))
run("H2D + step, no readback", lambda i: step(100 + i, host.to("cuda", non_blocking=True)))
run("H2D + step + float() per scalar (bench e2e)", lambda i: [float(v) for v in step(100 + i, host.to("cuda", non_blocking=True)).values()])
run("H2D + step + one stacked readback", lambda i: torch.stack(list(step(100 + i, host.to("cuda", non_blocking=True)).values())).cpu())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(30):
    torch.cuda.synchronize(); e0.record(); step(200 + i, dev); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"CUDA-event time of the same step (first launch -> last kernel end): {np.mean(ts):6.3f} ms")
loop.cuda_graph = True
step(0, dev)  # capture
run("CUDA graph: device batch, no readback", lambda i: step(100 + i, dev))
run("CUDA graph: H2D + step + float() per scalar", lambda i: [float(v) for v in step(100 + i, host.to("cuda", non_blocking=True)).values()])
