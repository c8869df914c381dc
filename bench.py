"""bench.py -- rays/sec of the NeRF hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

Workload at N=1: BASELINE.json configs[1] -- NeRF coarse+fine TRAINING STEP (fwd + bwd +
Adam) on synthetic rays, 4096 rays/GPU, 64 coarse + 128 fine samples, random-init
NeRFModel, bbox [-1,1]^3.  Rays shard over ranks (weak scaling: 4096 rays per GPU), one
NCCL all-reduce of the flat gradient per step.  `--workload render` times config[4]-style
rendering instead.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

FLOP_FWD_PER_SAMPLE = 1_182_976          # 2 x 591,488 MAC
FLOP_TRAIN_PER_SAMPLE = 3_481_344        # fwd + dW + dX, no padding / recompute
SAMPLES_PER_RAY = 64 + 192


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf=p["bf16_tflops_sustained"], tf_burst=p["bf16_tflops"],
                    source="measured")
    return dict(hbm=6650.0, tf=1400.0, tf_burst=1590.0, source="fallback")


def synth_batch(n, seed):
    """origins on the radius-4 sphere, directions toward a uniform bbox point,
    targets U(-1,1)."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g)
    o = 4.0 * o / o.norm(dim=1, keepdim=True)
    tgt = torch.rand(n, 3, generator=g) * 2 - 1
    d = tgt - o
    d = d / d.norm(dim=1, keepdim=True)
    col = torch.rand(n, 3, generator=g) * 2 - 1
    return torch.stack([o, d, col], dim=1).contiguous()


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=smax,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------ CPU arm
def cpu_train_sample(rays_per_step, steps, warmup, threads):
    """Times the oracle's train step (torch-CPU fp32 port of the reference path) on a bounded
    sample of the workload.  Returns (rays/s, seconds per step)."""
    from oracle import models_torch as M
    from oracle import train_torch as T
    torch.set_num_threads(threads)
    nerf = M.NeRFModel()
    params = T.init_params(nerf, nerf, 2)
    state = T.AdamState(params)
    batch = synth_batch(rays_per_step, 0).numpy()
    rs = np.random.RandomState(1)
    uc = (rs.randint(0, 2 ** 23, (rays_per_step, 64)) * 2.0 ** -23).astype(np.float32)
    uf = (rs.randint(0, 2 ** 23, (rays_per_step, 128)) * 2.0 ** -23).astype(np.float32)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        params, _ = T.train_step(nerf, nerf, params, state, 1e-4, dict(eps=1e-7), [-1, -1, -1],
                                 [1, 1, 1], batch, uc, uf, 64, 128)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return rays_per_step / sec, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rays = args.cpu_rays
    value, sec = cpu_train_sample(rays, args.steps, min(args.warmup, 1), threads)
    line = {
        "impl": "reference", "metric": "rays/sec (NeRF train step fwd+bwd+Adam)", "value": value,
        "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "configs[1]: NeRF coarse+fine train step, 64+128 samples/ray, "
                               "random-init 8x256 MLP", "rays_per_step": rays,
                   "note": "reference-restatement CPU baseline (oracle torch-CPU port; JAX is "
                           "not installable here), bounded sample of the 4096-ray step"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": f"{rays} rays/step x {args.steps} steps"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)

    from learn_nerf import _native
    from learn_nerf.model import NeRFModel
    from learn_nerf.render import NeRFRenderer
    from learn_nerf.train import TrainLoop

    peaks = load_peaks()
    if args.tc_stages is not None:
        _native.set_tc_stages(args.tc_stages)
    n = args.rays
    prec = args.precision
    coarse, fine = NeRFModel(precision=prec), NeRFModel(precision=prec)
    loop = TrainLoop(coarse, fine, init_rng=2, lr=1e-4, coarse_ts=64, fine_ts=128, device=dev,
                     ray_chunk=args.ray_chunk)
    bbox = ([-1.0, -1.0, -1.0], [1.0, 1.0, 1.0])
    step = loop.step_fn(*bbox)
    host_batch = synth_batch(n, rank).pin_memory()
    batch = host_batch.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    mlp_events = []

    def timed_mlp(fn):
        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            mlp_events.append((e0, e1))
            return r
        return wrapper

    orig_fwd, orig_bwd = _native.nerf_mlp_fwd, _native.nerf_mlp_bwd

    def one_step(i, host=False):
        if args.workload == "train":
            b = host_batch.to(dev, non_blocking=True) if host else batch
            logs = step(1000 + i, b)
            if host:
                return [float(v) for v in logs.values()]  # D2H read of the step's result
            return logs
        r = NeRFRenderer(coarse=coarse, fine=fine, coarse_params=loop.state.params["coarse"],
                         fine_params=loop.state.params["fine"],
                         background=loop.state.params["background"], bbox_min=bbox[0],
                         bbox_max=bbox[1], coarse_ts=64, fine_ts=128)
        b = host_batch.to(dev, non_blocking=True) if host else batch
        out = r.render_rays(1000 + i, b[:, :2].contiguous())["fine"]["outputs"]
        return out.cpu() if host else out

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(i)
    barrier()

    # ---- timed region 1: device-resident inputs, per-step CUDA events, L2 flushed between steps
    _native.nerf_mlp_fwd, _native.nerf_mlp_bwd = timed_mlp(orig_fwd), timed_mlp(orig_bwd)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _native.launch_count()
    evs = []
    barrier()
    for i in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step(i)
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = (_native.launch_count() - launches0) // max(args.steps, 1)
    clocks = sampler.stop()
    _native.nerf_mlp_fwd, _native.nerf_mlp_bwd = orig_fwd, orig_bwd
    step_ms = [a.elapsed_time(b) for a, b in evs]
    mlp_ms = sum(a.elapsed_time(b) for a, b in mlp_events) / max(args.steps, 1)
    ms = float(np.mean(step_ms))

    # ---- timed region 2 (e2e): host pinned batch -> H2D -> step -> D2H of the logged scalars
    barrier()
    t_e2e = []
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        one_step(i, host=True)
        torch.cuda.synchronize()
        t_e2e.append(time.perf_counter() - t0)
    barrier()
    e2e_ms = float(np.mean(t_e2e)) * 1e3

    t = torch.tensor([ms, e2e_ms, mlp_ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, e2e_ms, mlp_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        total_rays = n * world
        value = total_rays / (ms * 1e-3)
        e2e_value = total_rays / (e2e_ms * 1e-3)
        flop_per_sample = FLOP_TRAIN_PER_SAMPLE if args.workload == "train" else FLOP_FWD_PER_SAMPLE
        mlp_flops = flop_per_sample * SAMPLES_PER_RAY * n  # per rank, per step
        achieved_tf = mlp_flops / (mlp_ms * 1e-3) / 1e12 if mlp_ms > 0 else 0.0
        line = {
            "metric": "rays/sec (NeRF train step fwd+bwd+Adam)" if args.workload == "train"
                      else "rays/sec (NeRF render)",
            "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": prec, "data": "synthetic",
            "config": {"workload": ("configs[1]: NeRF coarse+fine train step" if args.workload == "train"
                                    else "NeRF coarse+fine render") +
                                   ", 64+128 samples/ray, random-init 8x256 MLP, bbox [-1,1]^3",
                       "rays_per_gpu": n, "mlp_precision": prec, "ray_chunk": args.ray_chunk,
                       "l2": "256 MiB flush between timed steps; per-step working set >> 126 MB L2",
                       "parallelism": f"ray-sharded dp{world}, NCCL all-reduce of flat grads"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(host_batch.numel() * 4),
                    "d2h_bytes_per_step": 16 if args.workload == "train" else int(n * 3 * 4)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "NeRF MLP fwd+bwd (all launches of "
                         "lnrf_nerf_mlp_fwd/_bwd)" if args.workload == "train" else "NeRF MLP fwd",
                         "achieved": achieved_tf, "peak": peaks["tf"], "unit": "TFLOP/s",
                         "frac": achieved_tf / peaks["tf"], "traffic": None,
                         "peak_source": peaks["source"] + " (sustained bf16 cuBLAS)",
                         "mlp_ms_per_step": mlp_ms, "mlp_share_of_step": mlp_ms / ms,
                         "algorithmic_flop_per_sample": flop_per_sample},
        }
        if args.cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            v, sec = cpu_train_sample(args.cpu_rays, 2, 1, threads)
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                                    "sample": f"oracle torch-CPU train step, {args.cpu_rays} rays/step x 2 "
                                              f"steps ({sec:.2f} s/step)"}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "render"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--rays", type=int, default=4096, help="rays per GPU per step")
    ap.add_argument("--ray_chunk", type=int, default=None)
    ap.add_argument("--cpu_rays", type=int, default=512, help="rays per step of the CPU sample")
    ap.add_argument("--tc_stages", type=int, default=None, help="bf16 kernel tuning knob")
    ap.add_argument("--no_cpu_baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
