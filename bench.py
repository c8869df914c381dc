"""bench.py -- rays/sec of the NeRF hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

Workload at N=1: BASELINE.json configs[1] -- NeRF coarse+fine TRAINING STEP (fwd + bwd +
Adam) on synthetic rays, 4096 rays/GPU, 64 coarse + 128 fine samples, random-init
NeRFModel, bbox [-1,1]^3.  Rays shard over ranks (weak scaling: 4096 rays per GPU), one
exchange of the flat gradient per step.  One JSON line is printed by rank 0.

After the headline's timed regions the same run measures every other BASELINE.json config with
the same method (own warm-up, CUDA-event step times with an L2 flush, e2e through the public API
with host buffers, roofline of the dominant kernels, clocks) and attaches them as
`extra_configs`: configs[1] fp32 leg, configs[0] (128x128 view in 1024-ray chunks, with its own
CPU-port baseline), configs[2] (Instant-NGP, 32,768 rays/GPU = 2^18 global at N=8), configs[3]
(Ref-NeRF) and configs[4] (800x800 render, rows sharded over the ranks); plus `hbm_stages`, the
achieved GB/s of the HBM-bound kernels timed alone.  `--no_extra` skips them; `--workload`,
`--model`, `--precision` select a single config as before.
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
# Ref-NeRF: fwd 590,336 MAC + normal chain 489,472 MAC; backward adds directional
# dW/dX 71,040, trunk dW 555,008 + dX 524,288, tangent pass 489,472 + its dW 489,472 MAC
REF_FLOP_FWD_PER_SAMPLE = 2 * (590_336 + 489_472)
REF_FLOP_TRAIN_PER_SAMPLE = 2 * 3_709_088
SAMPLES_PER_RAY = 64 + 192
# fp32-accurate paths: an UNFUSED chain of split-fp16 tcgen05 GEMMs (A by TMA) whose fp32 activations cross HBM once per
# producer and once per consumer, i.e. HBM-bound.  Algorithmic fp32 values moved per sample,
# every GEMM counted as (K inputs read + N outputs written), dW GEMMs as (M + N read), bit masks as 8 words:
#   NeRF  forward 5,264 (ten Dense layers + the two head kernels), backward 10,320 (heads, 13 dW, 9 dX)
#   Ref-NeRF forward 9,750 (trunk + 7-layer normal chain + d x_emb + heads), backward 19,464 (directional block,
#   8 dX + 10 dW of the first-order chain, tangent pass: 8 T GEMMs + 9 dW, three amax passes)
FP32_BYTES_PER_SAMPLE = {("nerf", False): 4 * 5264, ("nerf", True): 4 * (5264 + 10320),
                         ("refnerf", False): 4 * 9750, ("refnerf", True): 4 * (9750 + 19464)}
# ncu (profiles/r01d_nerf_train_kernels_ncu_full.txt): fwd 4.230 + dX 4.035 + dW 8.539 GB of DRAM
# traffic for the 786,432 samples of the fine level
NCU_TRAIN_DRAM_BYTES_PER_SAMPLE = (4.230e9 + 4.035e9 + 8.539e9) / 786432


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
    """SM clock + throttle reasons sampled DURING the timed region (NVML, every 10 ms)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.smax = index, [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def _uuid_index(self):
        # CUDA_VISIBLE_DEVICES may remap ordinals; torch's index -> NVML handle by UUID
        try:
            return "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid).replace("GPU-", "")
        except Exception:
            return None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = self._uuid_index()
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else None
            except Exception:
                h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {pynvml.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                     pynvml.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     pynvml.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     pynvml.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}

            def loop():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        for bit, nm in names.items():
                            if r & bit:
                                self.reasons.add(nm)
                    except Exception as e:  # noqa: BLE001
                        self.err = repr(e)
                        return
                    time.sleep(0.01)

            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.smax, reasons=[f"nvml unavailable: {self.err}"],
                        samples=0)
        return dict(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.smax,
                    reasons=sorted(self.reasons), samples=len(self.samples))


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


def cpu_render_sample(chunk_rays, chunks, threads):
    """configs[0] on the CPU: the oracle's render_rays (numpy sampling / compositing + torch-CPU fp32
    MLP) on `chunks` chunks of `chunk_rays` rays of a 128x128 view.  Returns (rays/s, s/chunk, rays)."""
    import math
    from oracle import models_torch as M
    from oracle import render_np
    from oracle import train_torch as T
    torch.set_num_threads(threads)
    nerf = M.NeRFModel()
    params = T.init_params(nerf, nerf, 2)
    rays = render_np.bare_rays((0.0, 0.0, -1.0), (0.0, 0.0, 4.0), (1.0, 0.0, 0.0), (0.0, -1.0, 0.0),
                               math.radians(60.0), math.radians(60.0), 128, 128)
    r = render_np.NeRFRenderer(M.as_numpy_model_fn(nerf, params["coarse"]), M.as_numpy_model_fn(nerf, params["fine"]),
                               params["background"].numpy(), np.float32([-1, -1, -1]), np.float32([1, 1, 1]), 64, 128)
    rs = np.random.RandomState(1)
    times = []
    for i in range(chunks + 1):
        a = (6 + i) * chunk_rays  # chunks from the middle of the view (they hit the box)
        sub = rays[a:a + chunk_rays]
        uc = (rs.randint(0, 2 ** 23, (len(sub), 64)) * 2.0 ** -23).astype(np.float32)
        uf = (rs.randint(0, 2 ** 23, (len(sub), 128)) * 2.0 ** -23).astype(np.float32)
        t0 = time.perf_counter()
        r.render_rays(uc, uf, sub)
        if i > 0:  # first chunk = warm-up
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return chunk_rays / sec, sec, chunk_rays * chunks


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rays = args.cpu_rays
    value, sec = cpu_train_sample(rays, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "rays/sec (NeRF train step fwd+bwd+Adam)", "value": value,
        "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "configs[1]: NeRF coarse+fine train step, 64+128 samples/ray, "
                               "random-init 8x256 MLP", "rays_per_step": rays,
                   "note": "PORT, SAMPLED: the oracle's torch-CPU fp32 restatement of the reference's train step "
                           "(JAX / Flax / optax are not installable here), on a bounded sample of the workload: "
                           f"{rays} of the 4096 rays of a step, every step"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": f"{rays}-ray sample of the 4096-ray step x {args.steps} steps ({sec:.2f} s/step), "
                                   f"{args.warmup} warm-up steps"},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------ GPU arm
def build_models(args, dev):
    """(coarse, fine, train_kwargs) as scripts/train_nerf.py:141-170 create_model does."""
    if args.model == "ngp":
        from learn_nerf.instant_ngp import InstantNGPModel
        mk = lambda L: InstantNGPModel(table_sizes=[2 ** 18] * L,
                                       grid_sizes=[2 ** (4 + i // 2) for i in range(L)],
                                       bbox_min=[-1.0] * 3, bbox_max=[1.0] * 3, precision=args.precision)
        return mk(6), mk(16), dict(adam_eps=1e-15, adam_b1=0.9, adam_b2=0.99)
    if args.model == "refnerf":
        from learn_nerf.ref_nerf import RefNERFModel
        return RefNERFModel(sh_degree=4), RefNERFModel(sh_degree=4), {}
    if args.model == "ngpref":  # train_nerf.py:141-170 with --instant_ngp --ref_nerf
        from learn_nerf.instant_ngp import InstantNGPRefNERFModel
        mk = lambda L: InstantNGPRefNERFModel(table_sizes=[2 ** 18] * L,
                                              grid_sizes=[2 ** (4 + i // 2) for i in range(L)],
                                              bbox_min=[-1.0] * 3, bbox_max=[1.0] * 3)
        return mk(6), mk(16), dict(adam_eps=1e-15, adam_b1=0.9, adam_b2=0.99)
    from learn_nerf.model import NeRFModel
    return NeRFModel(precision=args.precision), NeRFModel(precision=args.precision), {}


# algorithmic HBM/L2 bytes per point of the hash-grid kernels: 8 corners x 8 B per
# level gathered (+8 B/level written, +12 B coordinate); backward counts the RMW twice + d_enc.
def ngp_grid_bytes_per_ray(train):
    fwd = 64 * (6 * 72 + 12) + 192 * (16 * 72 + 12)
    bwd = 64 * (6 * (128 + 8) + 12) + 192 * (16 * (128 + 8) + 12)
    return fwd + (bwd if train else 0)


def measure(args, rank, local_rank, world, dev, peaks):
    """One config: warm-up, timed region 1 (device-resident inputs, CUDA events, L2 flush), timed
    region 2 (e2e through the public API with host buffers).  Returns the JSON line (rank 0) or None."""
    from learn_nerf import _native
    from learn_nerf.render import NeRFRenderer
    from learn_nerf.train import TrainLoop

    n = args.rays or (32768 if args.model in ("ngp", "ngpref") else 4096)
    prec = args.precision if args.model in ("nerf", "ngp") else "fp32"
    if args.model == "refnerf" and args.ray_chunk is None and n > 2048:
        args.ray_chunk = 2048  # 24 KB of saved activations per sample: keep the workspace near 10 GB
    coarse, fine, train_kwargs = build_models(args, dev)
    loop = TrainLoop(coarse, fine, init_rng=2, lr=1e-4, coarse_ts=64, fine_ts=128, device=dev,
                     ray_chunk=args.ray_chunk, **train_kwargs)
    bbox = ([-1.0, -1.0, -1.0], [1.0, 1.0, 1.0])
    step = loop.step_fn(*bbox)
    host_batch = synth_batch(n, rank).pin_memory()
    batch = host_batch.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # the dominant kernels, timed with CUDA events on the launching (current) stream
    dom_names = {"ngp": ["hashgrid_fwd", "hashgrid_bwd"], "refnerf": ["refnerf_fwd", "refnerf_bwd"],
                 "ngpref": ["ngpref_fwd", "ngpref_bwd"],
                 "nerf": ["nerf_mlp_fwd", "nerf_mlp_bwd"]}[args.model]
    dom_events = []

    def timed(fn, name):
        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            dom_events.append((e0, e1, name))
            return r
        return wrapper

    if world > 1 and args.workload == "train":
        dom_names = dom_names + ["adam_step_peers"]  # the fused NVLink gradient exchange + Adam
    originals = {nm: getattr(_native, nm) for nm in dom_names}

    if args.workload == "image":  # configs[4] / configs[0]: one W x H view, rows sharded over ranks
        from learn_nerf.dataset import CameraView
        from learn_nerf.scripts.render_nerf import render_view
        import math
        view = CameraView(camera_direction=(0.0, 0.0, -1.0), camera_origin=(0.0, 0.0, 4.0),
                          x_axis=(1.0, 0.0, 0.0), y_axis=(0.0, -1.0, 0.0), x_fov=math.radians(60.0),
                          y_fov=math.radians(60.0))
        img_renderer = NeRFRenderer(coarse=coarse, fine=fine, coarse_params=loop.state.params["coarse"],
                                    fine_params=loop.state.params["fine"],
                                    background=loop.state.params["background"], bbox_min=bbox[0],
                                    bbox_max=bbox[1], coarse_ts=64, fine_ts=128)
        n = args.width * args.height // world  # rays per GPU (for the JSON line)

    def one_step(i, host=False):
        if args.workload == "image":
            img = render_view(img_renderer, view, args.width, args.height, batch_size=args.batch_size,
                              key=1000 + i, device=dev, cuda_graph=args.cuda_graph)
            return img.cpu() if host else img
        b = host_batch.to(dev, non_blocking=True) if host else batch
        if args.workload == "train":
            logs = step(1000 + i, b)
            if host:
                return [float(v) for v in logs.values()]  # D2H read of the step's result
            return logs
        r = NeRFRenderer(coarse=coarse, fine=fine, coarse_params=loop.state.params["coarse"],
                         fine_params=loop.state.params["fine"],
                         background=loop.state.params["background"], bbox_min=bbox[0],
                         bbox_max=bbox[1], coarse_ts=64, fine_ts=128)
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
    for nm in dom_names:
        setattr(_native, nm, timed(originals[nm], nm))
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
    for nm in dom_names:
        setattr(_native, nm, originals[nm])
    step_ms = [a.elapsed_time(b) for a, b in evs]
    exch_ms = sum(a.elapsed_time(b) for a, b, k in dom_events if k == "adam_step_peers") / max(args.steps, 1)
    dom_events = [e for e in dom_events if e[2] != "adam_step_peers"]
    dom_names = [nm for nm in dom_names if nm != "adam_step_peers"]
    dom_ms = sum(a.elapsed_time(b) for a, b, _ in dom_events) / max(args.steps, 1)
    part_ms = {nm: sum(a.elapsed_time(b) for a, b, k in dom_events if k == nm) / max(args.steps, 1) for nm in dom_names}
    ms = float(np.mean(step_ms))

    # ---- timed region 2 (e2e): host pinned batch -> H2D -> step -> D2H of the logged scalars.
    # The train step is replayed as a CUDA graph here (TrainLoop(cuda_graph=True), the setting a user
    # would train with on one GPU); region 1 stays eager because it times individual C-ABI calls.
    graph_e2e = bool(args.cuda_graph and args.workload == "train" and (world == 1 or loop._peers is not None))
    graph_note = None
    if graph_e2e:
        try:
            loop.cuda_graph = True
            one_step(0, host=True)  # captures the graph (untimed)
            graph_e2e = loop._cg is not None  # e.g. Ref-NeRF with a ray_chunk stays eager
        except Exception as e:  # noqa: BLE001  (measure the eager step rather than lose the e2e number)
            loop.cuda_graph, loop._cg, graph_e2e = False, None, False
            graph_note = f"graph capture failed, eager step timed: {e!r}"[:200]
            torch.cuda.synchronize()
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

    image_note = None
    if args.workload == "image":
        # full chunks are replayed as CUDA graphs: the per-call events only see the eager tail chunk, so the
        # roofline of an image is taken over the WHOLE render (sampling + MLP + compositing + uint8 conversion)
        dom_ms = ms
        image_note = "whole render step (chunk graphs hide the individual C-ABI calls): the MLP kernel's share is not separated"
    t = torch.tensor([ms, e2e_ms, dom_ms, exch_ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, e2e_ms, dom_ms, exch_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        total_rays = n * world if args.workload != "image" else args.width * args.height
        value = total_rays / (ms * 1e-3)
        e2e_value = total_rays / (e2e_ms * 1e-3)
        train = args.workload == "train"
        what = {"nerf": "NeRF coarse+fine",
                "ngp": "Instant-NGP coarse (L=6) + fine (L=16), " + ("bf16 tcgen05 heads" if prec == "bf16" else "fp32-accurate heads (split-fp16 tcgen05 GEMMs)"),
                "ngpref": "Instant-NGP Ref-NeRF (smooth hash grid, sh_degree 4) coarse (L=6) + fine (L=16)",
                "refnerf": "Ref-NeRF (sh_degree 4) coarse+fine"}[args.model]
        cfg_name = {("nerf", True): "configs[1]: ", ("ngp", True): "configs[2]: ", ("refnerf", True): "configs[3]: ",
                    ("nerf", False): "configs[4]-style: "}.get((args.model, train), "")
        if args.workload == "image":
            which = "configs[0]" if (args.width, args.height, args.batch_size) == (128, 128, 1024) else "configs[4]"
            cfg_name = f"{which}: {args.width}x{args.height} view in chunks of {args.batch_size} rays, device-side ray generation and uint8 conversion, "
        if args.model == "ngpref":
            roofline = {"bound": "hbm", "kernel": "lnrf_ngpref_fwd" + (" + lnrf_ngpref_bwd" if train else ""),
                        "achieved": None, "peak": peaks["hbm"], "unit": "GB/s", "frac": None, "traffic": None,
                        "kernel_ms_per_step": dom_ms, "kernel_share_of_step": dom_ms / ms,
                        "note": "hash-grid gathers + 64-wide fp32 GEMM chain; no single dominant kernel yet"}
        elif args.model in ("nerf", "refnerf"):
            if args.model == "nerf":
                flop_per_sample = FLOP_TRAIN_PER_SAMPLE if train else FLOP_FWD_PER_SAMPLE
            else:
                flop_per_sample = REF_FLOP_TRAIN_PER_SAMPLE if train else REF_FLOP_FWD_PER_SAMPLE
            flops = flop_per_sample * SAMPLES_PER_RAY * n  # per rank, per step
            achieved = flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
            if prec == "fp32":  # unfused chain of fp32-in / fp32-out tensor-core GEMMs: bounded by HBM
                nbytes = FP32_BYTES_PER_SAMPLE[(args.model, train)] * SAMPLES_PER_RAY * n
                gbs = nbytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
                roofline = {"bound": "hbm",
                            "kernel": "tcg_rows_kernel + tcg_tn_kernel chain (split-fp16 tcgen05 GEMMs, three MMAs per "
                                      "fp32 product; fp32 activations cross HBM between layers)",
                            "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                            "traffic": None, "peak_source": peaks["source"] + " (HBM copy)",
                            "kernel_ms_per_step": dom_ms, "kernel_share_of_step": dom_ms / ms,
                            "algorithmic_bytes_per_sample": FP32_BYTES_PER_SAMPLE[(args.model, train)],
                            "algorithmic_flop_per_sample": flop_per_sample,
                            "algorithmic_tflops": achieved, "frac_of_bf16_tensor_peak": achieved / peaks["tf"],
                            "note": "1e-5 path; the tensor pipe executes 3x the algorithmic FLOPs (hi/lo split)"}
            else:
              # train: the MLP kernels run inside a long step -> sustained cuBLAS figure; render / image: the forward
              # kernel runs (nearly) alone for 0.3 - 0.6 ms per call and clocks stay at the top bin -> burst figure
              tf_peak, tf_name = ((peaks["tf"], "sustained") if train else (peaks["tf_burst"], "burst"))
              roofline = {"bound": "tensor",
                        "kernel": "nerf_fwd_cta2_kernel" + (" + nerf_bwd_dx_cta2_kernel + nerf_bwd_dw_kernel"
                                                            if train else ""),
                        "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                        "frac": achieved / tf_peak,
                        # DRAM bytes of the three MLP kernels per 4096-ray step, from the ncu --set full
                        # capture profiles/r01d_nerf_train_kernels_ncu_full.txt (fine level 4.23 + 4.03 +
                        # 8.56 GB, coarse level = 1/3 of it); equals the algorithmic stash bytes
                        # every byte of the activation stash is written once (forward / dX) and read once
                        # (dX / dW): 2 x the workspace sizes the library itself reports for the two levels
                        "traffic": (2 * sum(_native.nerf_mlp_workspace_bytes(n * T, _native.PREC_BF16, True)
                                            for T in (64, 192))
                                    if (train and prec == "bf16" and args.model == "nerf") else None),
                        "traffic_unit": "bytes per step: 2 x lnrf_nerf_mlp_workspace_bytes (stash written once, "
                                        "read once); ncu dram__bytes of the same kernels: profiles/",
                        "traffic_ncu_r01d": (NCU_TRAIN_DRAM_BYTES_PER_SAMPLE * SAMPLES_PER_RAY * n
                                             if (train and prec == "bf16" and args.model == "nerf") else None),
                        "peak_source": peaks["source"] + f" ({tf_name} bf16 cuBLAS)",
                        "frac_of_sustained_peak": achieved / peaks["tf"],
                        "kernel_ms_per_step": dom_ms, "kernel_share_of_step": dom_ms / ms,
                        "algorithmic_flop_per_sample": flop_per_sample,
                        "frac_of_burst_peak": achieved / peaks["tf_burst"]}
            if image_note:
                roofline["note"] = image_note
            if args.model == "nerf" and args.workload != "image":  # per C-ABI call: forward kernel vs dX + dW kernels
                fl = {"nerf_mlp_fwd": FLOP_FWD_PER_SAMPLE, "nerf_mlp_bwd": FLOP_TRAIN_PER_SAMPLE - FLOP_FWD_PER_SAMPLE}
                names = {"nerf_mlp_fwd": "nerf_fwd_cta2_kernel",
                         "nerf_mlp_bwd": "nerf_bwd_dx_cta2_kernel + nerf_bwd_dw_kernel"}
                roofline["parts"] = [
                    {"kernel": names[nm] if prec == "bf16" else nm + " (split-fp16 tcgen05 GEMM chain)", "ms_per_step": part_ms[nm],
                     "achieved": fl[nm] * SAMPLES_PER_RAY * n / (part_ms[nm] * 1e-3) / 1e12,
                     "frac": fl[nm] * SAMPLES_PER_RAY * n / (part_ms[nm] * 1e-3) / 1e12 /
                             (peaks["tf"] if (train or prec != "bf16") else peaks["tf_burst"])}
                    for nm in dom_names if part_ms.get(nm, 0) > 0]
        else:
            nbytes = ngp_grid_bytes_per_ray(train) * n
            achieved = nbytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
            roofline = {"bound": "hbm", "kernel": "hashgrid_fwd_kernel" + (" + hashgrid_bwd_kernel" if train else ""),
                        "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                        "frac": achieved / peaks["hbm"], "traffic": None,
                        "peak_source": peaks["source"] + " (HBM copy); the 30.6 MB of tables are "
                                       "L2-resident, so gathers are served by L2",
                        "kernel_ms_per_step": dom_ms, "kernel_share_of_step": dom_ms / ms,
                        "algorithmic_bytes_per_ray": ngp_grid_bytes_per_ray(train)}
        line = {
            "metric": f"rays/sec ({dict(nerf='NeRF', ngp='Instant-NGP', refnerf='Ref-NeRF', ngpref='Instant-NGP Ref-NeRF')[args.model]} "
                      f"{'train step fwd+bwd+Adam' if train else 'render'})",
            "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            # train / render: fixed rays per GPU (weak); image: one fixed view split over the ranks (strong)
            "scaling": "strong" if args.workload == "image" else "weak",
            "vs_baseline": None, "dtype": prec, "data": "synthetic",
            "config": {"workload": cfg_name + what + (" train step" if train else " render") +
                                   ", 64+128 samples/ray, random-init weights, bbox [-1,1]^3",
                       "rays_per_gpu": n, "mlp_precision": prec, "ray_chunk": args.ray_chunk,
                       "l2": "256 MiB flush between timed steps; per-step working set >> 126 MB L2",
                       "parallelism": f"ray-sharded dp{world}" +
                                      (", one exchange of the flat gradient per step (fused NVLink peer "
                                       "all-reduce + Adam; NCCL fallback)" if train else
                                       ", no collective")},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": e2e_ms, "cuda_graph": graph_e2e, "cuda_graph_note": graph_note,
                    "h2d_bytes_per_step": 0 if args.workload == "image" else int(host_batch.numel() * 4),
                    "d2h_bytes_per_step": 16 if train else (int(n * 3) if args.workload == "image" else int(n * 3 * 4))},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if world > 1 and train and exch_ms > 0:
            nbytes = int(loop._n_params + 4) * 4
            line["gradient_exchange"] = {
                "kernel": "adam_peers_kernel (fused all-reduce over NVLink peer mappings + Adam + norms)",
                "ms_per_step": exch_ms, "share_of_step": exch_ms / ms, "buffer_bytes": nbytes,
                "nvlink_bytes_read_per_rank": nbytes * (world - 1),
                "nvlink_gbs_per_rank": nbytes * (world - 1) / (exch_ms * 1e-3) / 1e9,
                "note": "CUDA-event time of the kernel alone (max over ranks); the two cross-rank barriers around it "
                        "are in the step time, not in this number"}
        if args.cpu_baseline and world == 1 and args.model == "nerf" and train and prec == "bf16":
            threads = os.cpu_count() or 1
            cpu_steps = 16  # ~10 s of CPU work on the box's host cores
            v, sec = cpu_train_sample(args.cpu_rays, cpu_steps, 1, threads)
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                                    "sample": f"oracle torch-CPU train step (port of the reference, fp32, not JAX), "
                                              f"{args.cpu_rays} rays/step x {cpu_steps} steps ({sec:.2f} s/step)"}
        elif (args.cpu_baseline and world == 1 and args.workload == "image" and args.width * args.height <= 128 * 128
              and args.model == "nerf"):
            threads = os.cpu_count() or 1
            v, sec, rays_done = cpu_render_sample(args.batch_size, 4, threads)
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                                    "sample": f"oracle numpy/torch-CPU render_rays (port of the reference, fp32, not "
                                              f"JAX), {rays_done} of the view's {args.width * args.height} rays in "
                                              f"{args.batch_size}-ray chunks ({sec:.2f} s/chunk)"}
        else:
            line["cpu_baseline"] = None
        return line
    return None


# every other BASELINE.json config, measured after the headline in the same run
EXTRA_CONFIGS = [
    ("configs[1] fp32 leg", dict(workload="train", model="nerf", precision="fp32", steps=5, warmup=3)),
    ("configs[0]", dict(workload="image", model="nerf", precision="bf16", width=128, height=128, batch_size=1024,
                        steps=10, warmup=3)),
    ("configs[2]", dict(workload="train", model="ngp", precision="bf16", steps=5, warmup=3)),
    ("configs[2] fp32 heads", dict(workload="train", model="ngp", precision="fp32", steps=5, warmup=3)),
    ("configs[3]", dict(workload="train", model="refnerf", steps=3, warmup=3)),
    ("configs[4]", dict(workload="image", model="nerf", precision="bf16", width=800, height=800, batch_size=65536,
                        steps=5, warmup=3)),
]


def run_ours(args):
    import copy
    import gc
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    line = measure(args, rank, local_rank, world, dev, peaks)
    headline = (args.workload, args.model, args.precision, args.rays) == ("train", "nerf", "bf16", None)
    if args.extra and headline:
        extras = []
        for name, over in EXTRA_CONFIGS:
            a = copy.copy(args)
            a.ray_chunk = None
            for k, v in over.items():
                setattr(a, k, v)
            gc.collect()
            torch.cuda.empty_cache()
            try:
                sub = measure(a, rank, local_rank, world, dev, peaks)
            except Exception as e:  # noqa: BLE001  (one failing extra must not lose the headline)
                sub = {"error": repr(e)[:300]}
                torch.cuda.synchronize()
            if rank == 0 and sub is not None:
                sub["name"] = name
                extras.append(sub)
        stages = None
        if rank == 0:
            gc.collect()
            torch.cuda.empty_cache()
            try:
                sys.path.insert(0, os.path.join(ROOT, "profiles"))
                from stage_bench import run_stages
                stages = run_stages(262144, dev, peaks["hbm"])
            except Exception as e:  # noqa: BLE001
                stages = [{"error": repr(e)[:300]}]
        if world > 1:
            torch.distributed.barrier()
        if rank == 0:
            line["extra_configs"] = extras
            line["hbm_stages"] = stages
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "render", "image"])
    ap.add_argument("--width", type=int, default=800)
    ap.add_argument("--height", type=int, default=800)
    ap.add_argument("--batch_size", type=int, default=65536, help="rays per render_rays call (image workload)")
    ap.add_argument("--model", default="nerf", choices=["nerf", "ngp", "refnerf", "ngpref"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"],
                    help="MLP path: bf16 tcgen05 (2e-2) or fp32-accurate (1e-5: split-fp16 tcgen05 GEMMs)")
    ap.add_argument("--rays", type=int, default=None, help="rays per GPU per step (4096 NeRF, 32768 NGP)")
    ap.add_argument("--ray_chunk", type=int, default=None)
    ap.add_argument("--cpu_rays", type=int, default=512, help="rays per step of the CPU sample")
    ap.add_argument("--no_cpu_baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no_extra", dest="extra", action="store_false",
                    help="headline only: skip the extra_configs / hbm_stages blocks")
    ap.add_argument("--no_cuda_graph", dest="cuda_graph", action="store_false",
                    help="run the end-to-end region with the eager step instead of the CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
