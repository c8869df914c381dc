"""ctypes binding of liblnrf.so (include/lnrf.h).

PyTorch is used only for device memory and streams; every compute call below lands
in a hand-written sm_100a kernel.  There is NO fallback: if the shared library is
missing or a call fails, an exception is raised.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_void_p
from typing import Optional

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "liblnrf.so")

PREC_FP32 = 0
PREC_BF16 = 1
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16}


class LnrfError(RuntimeError):
    pass


_lib = None
_inited_devices = set()

_SIGS = {
    "lnrf_last_error": (c_char_p, []),
    "lnrf_version": (c_int32, []),
    "lnrf_init": (c_int32, [c_int32]),
    "lnrf_launch_count": (c_int64, []),
    "lnrf_sample_coarse": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_float, c_float,
                                     c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "lnrf_stratified": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "lnrf_sample_fine": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_int32, c_int32, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lnrf_ray_intervals": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "lnrf_termination_probs": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p,
                                         c_void_p]),
    "lnrf_z_depth": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int64, c_void_p, c_void_p,
                               c_void_p]),
    "lnrf_composite_fwd": (c_int32, [c_void_p] * 8 + [c_int64, c_int32] + [c_void_p] * 4),
    "lnrf_composite_bwd": (c_int32, [c_void_p] * 8 + [c_int64, c_int32] + [c_void_p] * 4),
    "lnrf_mse_loss": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p,
                                c_void_p]),
    "lnrf_nerf_param_count": (c_int64, []),
    "lnrf_nerf_param_floats": (c_int64, []),
    "lnrf_nerf_param_offsets": (c_int32, [c_void_p]),
    "lnrf_nerf_packed_bytes": (c_int64, []),
    "lnrf_nerf_pack_weights": (c_int32, [c_void_p, c_void_p, c_void_p]),
    "lnrf_nerf_mlp_workspace_bytes": (c_int32, [c_int64, c_int32, c_int32, c_void_p]),
    "lnrf_nerf_mlp_fwd": (c_int32, [c_void_p] * 6 + [c_int64, c_int32, c_int32, c_int32, c_void_p,
                                                     c_int64, c_void_p, c_void_p, c_void_p]),
    "lnrf_nerf_mlp_bwd": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int64] +
                          [c_void_p] * 6),
    "lnrf_nerf_render_workspace_bytes": (c_int32, [c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "lnrf_nerf_render_rays": (c_int32, [c_void_p, c_void_p, c_void_p, c_float] + [c_void_p] * 6 + [c_int32, c_void_p, c_int64,
                                        c_int32, c_int32, c_void_p, c_int64] + [c_void_p] * 5),
    "lnrf_nerf_train_workspace_bytes": (c_int32, [c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "lnrf_nerf_train_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_float] + [c_void_p] * 8 + [c_int32, c_int64, c_int32,
                                       c_int32, c_float, c_float, c_float, c_float, c_int32, c_void_p, c_int64, c_void_p,
                                       c_void_p]),
    "lnrf_adam_step": (c_int32, [c_void_p] * 4 + [c_int64, c_float, c_float, c_float, c_float,
                                                  c_int32, c_float, c_void_p, c_void_p]),
    "lnrf_adam_step_dk": (c_int32, [c_void_p] * 4 + [c_int64, c_float, c_float, c_float, c_float, c_void_p,
                                     c_float, c_void_p, c_void_p]),
    "lnrf_threefry_uniform_dk": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p]),
    "lnrf_adam_step_peers": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_int32,
                                       c_float, c_float, c_float, c_float, c_int32, c_float, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "lnrf_debug_umma_gemm": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "lnrf_debug_umma_gemm_tn": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "lnrf_tcgemm": (c_int32, [c_int32, c_int32, c_int64, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32,
                              c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_int32] + [c_void_p] * 9),
    "lnrf_bare_rays": (c_int32, [c_void_p] * 4 + [c_float, c_float, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                                   c_void_p]),
    "lnrf_rgb_to_u8": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p]),
    "lnrf_threefry_uniform": (c_int32, [ctypes.c_uint32, ctypes.c_uint32, c_int64, c_void_p, c_void_p]),
    "lnrf_refnerf_param_count": (c_int64, []),
    "lnrf_refnerf_param_floats": (c_int64, []),
    "lnrf_refnerf_param_offsets": (c_int32, [c_void_p]),
    "lnrf_refnerf_workspace_bytes": (c_int32, [c_int64, c_int32, c_void_p]),
    "lnrf_refnerf_fwd": (c_int32, [c_void_p] * 5 + [c_int64, c_int32, c_int32, c_void_p, c_int64] +
                         [c_void_p] * 5),
    "lnrf_refnerf_bwd": (c_int32, [c_void_p] * 5 + [c_int64, c_int32, c_void_p, c_int64] + [c_void_p] * 6),
    "lnrf_ngpref_mlp_param_floats": (c_int64, [c_int32]),
    "lnrf_ngpref_param_offsets": (c_int32, [c_int32, c_void_p]),
    "lnrf_ngpref_workspace_bytes": (c_int32, [c_int64, c_int32, c_int32, c_void_p]),
    "lnrf_ngpref_fwd": (c_int32, [c_void_p] * 4 + [c_int32, c_void_p, c_void_p] + [c_void_p] * 4 +
                        [c_int64, c_int32, c_int32, c_void_p, c_int64] + [c_void_p] * 5),
    "lnrf_ngpref_bwd": (c_int32, [c_void_p] * 4 + [c_int32, c_void_p, c_void_p] + [c_void_p] * 4 +
                        [c_int64, c_int32, c_void_p, c_int64] + [c_void_p] * 6),
    "lnrf_ngp_mlp_param_offsets": (c_int32, [c_int32, c_void_p]),
    "lnrf_hashgrid_fwd": (c_int32, [c_void_p] * 4 + [c_int32, c_void_p, c_void_p, c_int32,
                                                     c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                                     c_void_p, c_void_p]),
    "lnrf_hashgrid_bwd": (c_int32, [c_void_p] * 3 + [c_int32, c_void_p, c_void_p, c_int32,
                                                     c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                                     c_void_p, c_void_p, c_void_p]),
    "lnrf_ngp_packed_bytes": (c_int64, []),
    "lnrf_ngp_pack_weights": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p]),
    "lnrf_ngp_mlp_tc_workspace_bytes": (c_int32, [c_int64, c_void_p]),
    "lnrf_ngp_mlp_fwd_tc": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                      c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "lnrf_ngp_mlp_bwd_tc": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_int64] + [c_void_p] * 7),
    "lnrf_ngp_mlp_param_count": (c_int64, [c_int32]),
    "lnrf_ngp_mlp_workspace_bytes": (c_int32, [c_int64, c_int32, c_void_p]),
    "lnrf_ngp_mlp_fwd": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                   c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "lnrf_ngp_mlp_bwd": (c_int32, [c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_int64] +
                         [c_void_p] * 7),
}


def lib_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """dlopen liblnrf.so and declare the prototypes.  Needs no GPU."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise LnrfError(
                f"{_LIB_PATH} not found: build it with `python learn-nerf_b200/build.py` "
                "(there is no CPU or PyTorch fallback for the render/train hot path)")
        lib = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().lnrf_last_error()
        raise LnrfError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ensure_init(device: torch.device):
    if device.type != "cuda":
        raise LnrfError(f"liblnrf kernels need CUDA tensors, got device {device} (no CPU fallback)")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _inited_devices:
        with torch.cuda.device(idx):
            _check(load().lnrf_init(idx), "lnrf_init")
        _inited_devices.add(idx)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else c_void_p(t.data_ptr())


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise LnrfError(f"{name}: expected a contiguous CUDA float32 tensor, got {t.dtype} "
                        f"{t.device} contiguous={t.is_contiguous()}")
    return t


def _host3(v):
    arr = (c_float * 3)(*[float(x) for x in v])
    return arr


# --------------------------------------------------------------------------- wrappers
def sample_coarse(rays, bbox_min, bbox_max, u, min_t_range=1e-3, epsilon=1e-8):
    rays, u = _f32c(rays, "rays"), _f32c(u, "u")
    ensure_init(rays.device)
    n, T = u.shape
    t_min = torch.empty(n, device=rays.device)
    t_max = torch.empty(n, device=rays.device)
    mask = torch.empty(n, dtype=torch.uint8, device=rays.device)
    ts = torch.empty(n, T, device=rays.device)
    lo, hi = _host3(bbox_min), _host3(bbox_max)
    _check(load().lnrf_sample_coarse(_p(rays), n, lo, hi, min_t_range, epsilon, _p(u), T, _p(t_min),
                                     _p(t_max), _p(mask), _p(ts), _stream()), "lnrf_sample_coarse")
    return t_min, t_max, mask, ts


def stratified(t_min, t_max, u):
    u = _f32c(u, "u")
    ensure_init(u.device)
    n, T = u.shape
    ts = torch.empty(n, T, device=u.device)
    _check(load().lnrf_stratified(_p(_f32c(t_min, "t_min")), _p(_f32c(t_max, "t_max")), _p(u), n, T,
                                  _p(ts), _stream()), "lnrf_stratified")
    return ts


def sample_fine(ts_c, dens_c, t_min, t_max, u, eps=1e-8, debug=False):
    ts_c, dens_c, u = _f32c(ts_c, "ts"), _f32c(dens_c, "densities"), _f32c(u, "u")
    ensure_init(ts_c.device)
    n, Tc = ts_c.shape
    Tf = u.shape[1]
    out = torch.empty(n, Tc + Tf, device=ts_c.device)
    idx = torch.empty(n, Tf, dtype=torch.int32, device=ts_c.device) if debug else None
    new_ts = torch.empty(n, Tf, device=ts_c.device) if debug else None
    _check(load().lnrf_sample_fine(_p(ts_c), _p(dens_c), _p(_f32c(t_min, "t_min")),
                                   _p(_f32c(t_max, "t_max")), _p(u), n, Tc, Tf, eps, _p(out), _p(idx),
                                   _p(new_ts), _stream()), "lnrf_sample_fine")
    return (out, idx, new_ts) if debug else out


def composite_fwd(rays, ts, t_min, t_max, mask, dens, rgb, background, want_aux=True):
    ensure_init(ts.device)
    n, T = ts.shape
    dev = ts.device
    outputs = torch.empty(n, 3, device=dev)
    alphas = torch.empty(n, 1, device=dev) if want_aux else None
    coords = torch.empty(n, 3, device=dev) if want_aux else None
    _check(load().lnrf_composite_fwd(_p(_f32c(rays, "rays")), _p(_f32c(ts, "ts")), _p(t_min), _p(t_max),
                                     _p(mask), _p(_f32c(dens, "densities")), _p(_f32c(rgb, "rgbs")),
                                     _p(_f32c(background, "background")), n, T, _p(outputs),
                                     _p(alphas), _p(coords), _stream()), "lnrf_composite_fwd")
    return outputs, alphas, coords


def composite_bwd(ts, t_min, t_max, mask, dens, rgb, background, d_outputs, d_background):
    ensure_init(ts.device)
    n, T = ts.shape
    d_dens = torch.empty(n, T, device=ts.device)
    d_rgb = torch.empty(n, T, 3, device=ts.device)
    _check(load().lnrf_composite_bwd(_p(ts), _p(t_min), _p(t_max), _p(mask), _p(_f32c(dens, "dens")),
                                     _p(_f32c(rgb, "rgb")), _p(background),
                                     _p(_f32c(d_outputs, "d_outputs")), n, T, _p(d_dens), _p(d_rgb),
                                     _p(d_background), _stream()), "lnrf_composite_bwd")
    return d_dens, d_rgb


def mse_loss(outputs, targets_base, target_stride, n, inv_count, loss_sum, d_outputs):
    """targets_base: tensor whose data_ptr is element (0,0) of the strided targets view."""
    ensure_init(outputs.device)
    _check(load().lnrf_mse_loss(_p(_f32c(outputs, "outputs")), _p(targets_base), target_stride, n,
                                inv_count, _p(loss_sum), _p(d_outputs), _stream()), "lnrf_mse_loss")


def launch_count() -> int:
    return int(load().lnrf_launch_count())


def nerf_param_count() -> int:
    return int(load().lnrf_nerf_param_count())


def nerf_param_floats() -> int:
    return int(load().lnrf_nerf_param_floats())


def nerf_param_offsets():
    buf = (c_int64 * 24)()
    _check(load().lnrf_nerf_param_offsets(buf), "lnrf_nerf_param_offsets")
    return [int(v) for v in buf]


def nerf_packed_bytes() -> int:
    return int(load().lnrf_nerf_packed_bytes())


def nerf_pack_weights(flat: torch.Tensor, packed: torch.Tensor):
    ensure_init(flat.device)
    _check(load().lnrf_nerf_pack_weights(_p(_f32c(flat, "params")), _p(packed), _stream()),
           "lnrf_nerf_pack_weights")


def nerf_mlp_workspace_bytes(m: int, precision: int, save: bool) -> int:
    out = c_int64(0)
    _check(load().lnrf_nerf_mlp_workspace_bytes(m, precision, int(save), ctypes.byref(out)),
           "lnrf_nerf_mlp_workspace_bytes")
    return int(out.value)


def nerf_render_rays(rays, bbox_min, bbox_max, u_coarse, u_fine, coarse_flat, coarse_packed, fine_flat, fine_packed,
                     precision, background, min_t_range=1e-3, want_aux=True):
    """One C call for NeRFRenderer.render_rays (reference render.py:39-91) with two NeRFModels: returns
    (coarse_outputs[n,3], fine_outputs[n,3], fine_alphas[n,1] | None, fine_coords[n,3] | None)."""
    rays, u_coarse, u_fine = _f32c(rays, "rays"), _f32c(u_coarse, "u_coarse"), _f32c(u_fine, "u_fine")
    ensure_init(rays.device)
    n, Tc = u_coarse.shape
    Tf = u_fine.shape[1]
    dev = rays.device
    nbytes = c_int64(0)
    _check(load().lnrf_nerf_render_workspace_bytes(n, Tc, Tf, precision, ctypes.byref(nbytes)),
           "lnrf_nerf_render_workspace_bytes")
    raw = torch.empty(int(nbytes.value) + 1024, dtype=torch.uint8, device=dev)
    shift = (-raw.data_ptr()) % 1024
    ws = raw[shift: shift + int(nbytes.value)]
    coarse_out = torch.empty(n, 3, device=dev)
    fine_out = torch.empty(n, 3, device=dev)
    alphas = torch.empty(n, 1, device=dev) if want_aux else None
    coords = torch.empty(n, 3, device=dev) if want_aux else None
    lo, hi = _host3(bbox_min), _host3(bbox_max)
    _check(load().lnrf_nerf_render_rays(_p(rays), lo, hi, min_t_range, _p(u_coarse), _p(u_fine), _p(coarse_flat),
                                        _p(coarse_packed), _p(fine_flat), _p(fine_packed), precision,
                                        _p(_f32c(background, "background")), n, Tc, Tf, _p(ws), int(nbytes.value),
                                        _p(coarse_out), _p(fine_out), _p(alphas), _p(coords), _stream()),
           "lnrf_nerf_render_rays")
    return coarse_out, fine_out, alphas, coords


def _aligned_bytes(nbytes: int, device) -> torch.Tensor:
    raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    shift = (-raw.data_ptr()) % 1024
    return raw[shift: shift + nbytes]


def nerf_train_step(batch, bbox_min, bbox_max, u_coarse, u_fine, flat, m, v, grads, precision, lr, b1, b2, eps, step,
                    min_t_range=1e-3):
    """One C call for TrainLoop.step_fn (reference train.py:78-112) with two NeRFModels on one device.
    flat / m / v / grads: [coarse | fine | background(3) + pad].  Returns the device scalars
    [sse_coarse, sse_fine, |g|^2, |p|^2]."""
    batch, u_coarse, u_fine = _f32c(batch, "batch"), _f32c(u_coarse, "u_coarse"), _f32c(u_fine, "u_fine")
    ensure_init(batch.device)
    n, Tc = u_coarse.shape
    Tf = u_fine.shape[1]
    dev = batch.device
    nbytes = c_int64(0)
    _check(load().lnrf_nerf_train_workspace_bytes(n, Tc, Tf, precision, ctypes.byref(nbytes)),
           "lnrf_nerf_train_workspace_bytes")
    ws = _aligned_bytes(int(nbytes.value), dev)
    pc = pf = None
    if precision == PREC_BF16:
        pc, pf = _aligned_bytes(nerf_packed_bytes(), dev), _aligned_bytes(nerf_packed_bytes(), dev)
    scalars = torch.empty(4, device=dev)
    lo, hi = _host3(bbox_min), _host3(bbox_max)
    _check(load().lnrf_nerf_train_step(_p(batch), lo, hi, min_t_range, _p(u_coarse), _p(u_fine), _p(flat), _p(m), _p(v),
                                       _p(grads), _p(pc), _p(pf), precision, n, Tc, Tf, lr, b1, b2, eps, step, _p(ws),
                                       int(nbytes.value), _p(scalars), _stream()), "lnrf_nerf_train_step")
    return scalars


def nerf_mlp_fwd(flat, packed, x, d, rays, ts, n, T, precision, save, workspace, dens, rgb):
    ensure_init(flat.device)
    ws_bytes = 0 if workspace is None else workspace.numel() * workspace.element_size()
    _check(load().lnrf_nerf_mlp_fwd(_p(flat), _p(packed), _p(x), _p(d), _p(rays), _p(ts), n, T,
                                    precision, int(save), _p(workspace), ws_bytes, _p(dens), _p(rgb),
                                    _stream()), "lnrf_nerf_mlp_fwd")


def nerf_mlp_bwd(flat, packed, m, precision, workspace, dens, rgb, d_dens, d_rgb, d_flat):
    ensure_init(flat.device)
    ws_bytes = workspace.numel() * workspace.element_size()
    _check(load().lnrf_nerf_mlp_bwd(_p(flat), _p(packed), m, precision, _p(workspace), ws_bytes,
                                    _p(dens), _p(rgb), _p(_f32c(d_dens, "d_dens")),
                                    _p(_f32c(d_rgb, "d_rgb")), _p(d_flat), _stream()),
           "lnrf_nerf_mlp_bwd")


def adam_step(params, grads, m, v, lr, b1, b2, eps, step, grad_scale, norms_out):
    ensure_init(params.device)
    _check(load().lnrf_adam_step(_p(_f32c(params, "params")), _p(_f32c(grads, "grads")), _p(m), _p(v),
                                 params.numel(), lr, b1, b2, eps, step, grad_scale, _p(norms_out),
                                 _stream()), "lnrf_adam_step")


class PinnedRing:
    """Race-free asynchronous upload of a few words per step (CUDA-graph replays): a ring of pinned
    staging buffers, each guarded by the event of the copy that last read it, so the host never
    rewrites a buffer whose H2D copy has not executed yet and never waits unless it runs ``depth``
    steps ahead of the GPU."""

    def __init__(self, words: int, depth: int = 8):
        self.bufs = [torch.zeros(words, dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.views = [b.numpy() for b in self.bufs]
        self.events = [None] * depth
        self.i = 0

    def upload(self, values, dst: torch.Tensor):
        """values: int32 numpy array of ``words`` elements; dst: int32 device tensor."""
        j = self.i
        self.i = (j + 1) % len(self.bufs)
        if self.events[j] is not None:
            self.events[j].synchronize()
        self.views[j][:] = values
        dst.copy_(self.bufs[j], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dst.device))
        self.events[j] = ev


def adam_step_dk(params, grads, m, v, lr, b1, b2, eps, inv_bias_corr_dev, grad_scale, norms_out):
    """lnrf_adam_step with the bias corrections {1/(1-b1^t), 1/(1-b2^t)} in device memory."""
    ensure_init(params.device)
    _check(load().lnrf_adam_step_dk(_p(_f32c(params, "params")), _p(_f32c(grads, "grads")), _p(m), _p(v),
                                    params.numel(), lr, b1, b2, eps, _p(inv_bias_corr_dev), grad_scale,
                                    _p(norms_out), _stream()), "lnrf_adam_step_dk")


def threefry_uniform_dk(key_dev: torch.Tensor, shape) -> torch.Tensor:
    """Uniforms from a Threefry key held in device memory (int32/uint32 tensor of 2 words)."""
    ensure_init(key_dev.device)
    out = torch.empty(tuple(shape), device=key_dev.device)
    _check(load().lnrf_threefry_uniform_dk(_p(key_dev), out.numel(), _p(out), _stream()),
           "lnrf_threefry_uniform_dk")
    return out


def adam_step_peers(params, peer_ptrs, m, v, count, extra, lr, b1, b2, eps, step, grad_scale, norms_out,
                    extra_out, inv_bias_corr_dev=None):
    """Fused all-reduce + Adam over NVLink peer mappings of every rank's flat gradient buffer.
    ``peer_ptrs``: device addresses (ints), one per rank, rank order."""
    ensure_init(params.device)
    arr = (ctypes.c_uint64 * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
    _check(load().lnrf_adam_step_peers(_p(_f32c(params, "params")), ctypes.cast(arr, c_void_p), len(peer_ptrs),
                                       _p(m), _p(v), count, extra, lr, b1, b2, eps, step, grad_scale,
                                       _p(norms_out), _p(extra_out), _p(inv_bias_corr_dev), _stream()),
           "lnrf_adam_step_peers")


def debug_umma_gemm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """D[128,N] = bf16(A[128,K]) @ bf16(B[N,K])^T on tcgen05 (descriptor self-test)."""
    ensure_init(a.device)
    N, K = b.shape
    out = torch.empty(128, N, device=a.device)
    _check(load().lnrf_debug_umma_gemm(_p(_f32c(a, "a")), _p(_f32c(b, "b")), N, K, _p(out), _stream()),
           "lnrf_debug_umma_gemm")
    return out


def debug_umma_gemm_tn(at: torch.Tensor, bt: torch.Tensor) -> torch.Tensor:
    """D[M,N] = bf16(At[128,M])^T @ bf16(Bt[128,N]) with MN-major descriptors (dW shape)."""
    ensure_init(at.device)
    M, N = at.shape[1], bt.shape[1]
    out = torch.empty(M, N, device=at.device)
    _check(load().lnrf_debug_umma_gemm_tn(_p(_f32c(at, "at")), _p(_f32c(bt, "bt")), M, N, _p(out),
                                          _stream()), "lnrf_debug_umma_gemm_tn")
    return out


# --------------------------------------------------------------------------- Instant-NGP
class GridSpec:
    """Host-side description of a multiresolution table set (instant_ngp.py:92-118)."""

    def __init__(self, table_sizes, grid_sizes, bbox_min, bbox_max, smooth=False, base_offset=0):
        self.L = len(grid_sizes)
        self.table_sizes = [int(t) for t in table_sizes]
        self.grid_sizes = [int(g) for g in grid_sizes]
        self.rows = [t if g ** 3 > t else g ** 3 for t, g in zip(self.table_sizes, self.grid_sizes)]
        offs, off = [], int(base_offset)
        for r in self.rows:
            offs.append(off)
            off += (r * 2 + 3) // 4 * 4
        self.offsets, self.end = offs, off
        self.smooth = int(bool(smooth))
        self._off = (c_int64 * self.L)(*offs)
        self._grid = (c_int32 * self.L)(*self.grid_sizes)
        self._tab = (c_int32 * self.L)(*self.table_sizes)
        self._lo, self._hi = _host3(bbox_min), _host3(bbox_max)


def hashgrid_fwd(tables_flat, spec: GridSpec, x, rays, ts, n, T, enc):
    ensure_init(tables_flat.device)
    _check(load().lnrf_hashgrid_fwd(_p(tables_flat), spec._off, spec._grid, spec._tab, spec.L, spec._lo,
                                    spec._hi, spec.smooth, _p(x), _p(rays), _p(ts), n, T, _p(enc),
                                    _stream()), "lnrf_hashgrid_fwd")


def hashgrid_bwd(spec: GridSpec, x, rays, ts, n, T, d_enc, d_tables_flat):
    ensure_init(d_enc.device)
    _check(load().lnrf_hashgrid_bwd(spec._off, spec._grid, spec._tab, spec.L, spec._lo, spec._hi,
                                    spec.smooth, _p(x), _p(rays), _p(ts), n, T, _p(_f32c(d_enc, "d_enc")),
                                    _p(d_tables_flat), _stream()), "lnrf_hashgrid_bwd")


def ngpref_mlp_param_floats(L: int) -> int:
    return int(load().lnrf_ngpref_mlp_param_floats(L))


def ngpref_param_offsets(L: int):
    buf = (c_int64 * 10)()
    _check(load().lnrf_ngpref_param_offsets(L, buf), "lnrf_ngpref_param_offsets")
    return [int(v) for v in buf]


def ngpref_workspace_bytes(m: int, L: int, save: bool) -> int:
    out = c_int64(0)
    _check(load().lnrf_ngpref_workspace_bytes(m, L, int(save), ctypes.byref(out)), "lnrf_ngpref_workspace_bytes")
    return int(out.value)


def ngpref_fwd(flat, spec: GridSpec, x, d, rays, ts, n, T, save, workspace, dens, rgb, aux_mse, aux_neg):
    ensure_init(flat.device)
    _check(load().lnrf_ngpref_fwd(_p(flat), spec._off, spec._grid, spec._tab, spec.L, spec._lo, spec._hi, _p(x),
                                  _p(d), _p(rays), _p(ts), n, T, int(save), _p(workspace), workspace.numel(),
                                  _p(dens), _p(rgb), _p(aux_mse), _p(aux_neg), _stream()), "lnrf_ngpref_fwd")


def ngpref_bwd(flat, spec: GridSpec, x, d, rays, ts, n, T, workspace, d_dens, d_rgb, d_mse, d_neg, d_flat):
    ensure_init(flat.device)
    _check(load().lnrf_ngpref_bwd(_p(flat), spec._off, spec._grid, spec._tab, spec.L, spec._lo, spec._hi, _p(x),
                                  _p(d), _p(rays), _p(ts), n, T, _p(workspace), workspace.numel(),
                                  _p(_f32c(d_dens, "d_dens")), _p(_f32c(d_rgb, "d_rgb")), _p(_f32c(d_mse, "d_mse")),
                                  _p(_f32c(d_neg, "d_neg")), _p(d_flat), _stream()), "lnrf_ngpref_bwd")


def ngp_mlp_param_floats(L: int) -> int:
    return int(load().lnrf_ngp_mlp_param_count(L))


def ngp_mlp_param_offsets(L: int):
    buf = (c_int64 * 10)()
    _check(load().lnrf_ngp_mlp_param_offsets(L, buf), "lnrf_ngp_mlp_param_offsets")
    return [int(v) for v in buf]


def ngp_mlp_workspace_bytes(m: int, L: int) -> int:
    out = c_int64(0)
    _check(load().lnrf_ngp_mlp_workspace_bytes(m, L, ctypes.byref(out)), "lnrf_ngp_mlp_workspace_bytes")
    return int(out.value)


def ngp_mlp_fwd(flat, L, enc, d, rays, n, T, save, workspace, dens, rgb):
    ensure_init(flat.device)
    _check(load().lnrf_ngp_mlp_fwd(_p(flat), L, _p(enc), _p(d), _p(rays), n, T, int(save), _p(workspace),
                                   0 if workspace is None else workspace.numel(), _p(dens), _p(rgb),
                                   _stream()), "lnrf_ngp_mlp_fwd")


def ngp_mlp_bwd(flat, L, enc, m, workspace, dens, rgb, d_dens, d_rgb, d_flat, d_enc):
    ensure_init(flat.device)
    _check(load().lnrf_ngp_mlp_bwd(_p(flat), L, _p(enc), m, _p(workspace), workspace.numel(), _p(dens),
                                   _p(rgb), _p(_f32c(d_dens, "d_dens")), _p(_f32c(d_rgb, "d_rgb")),
                                   _p(d_flat), _p(d_enc), _stream()), "lnrf_ngp_mlp_bwd")


def ngp_packed_bytes() -> int:
    return int(load().lnrf_ngp_packed_bytes())


def ngp_pack_weights(flat, L, packed):
    ensure_init(flat.device)
    _check(load().lnrf_ngp_pack_weights(_p(_f32c(flat, "params")), L, _p(packed), _stream()), "lnrf_ngp_pack_weights")


def ngp_mlp_tc_workspace_bytes(m: int) -> int:
    out = c_int64(0)
    _check(load().lnrf_ngp_mlp_tc_workspace_bytes(m, ctypes.byref(out)), "lnrf_ngp_mlp_tc_workspace_bytes")
    return int(out.value)


def ngp_mlp_fwd_tc(packed, L, enc, d, rays, n, T, save, workspace, dens, rgb):
    ensure_init(enc.device)
    _check(load().lnrf_ngp_mlp_fwd_tc(_p(packed), L, _p(enc), _p(d), _p(rays), n, T, int(save), _p(workspace),
                                      0 if workspace is None else workspace.numel(), _p(dens), _p(rgb), _stream()),
           "lnrf_ngp_mlp_fwd_tc")


def ngp_mlp_bwd_tc(packed, L, m, workspace, dens, rgb, d_dens, d_rgb, d_flat, d_enc):
    ensure_init(d_enc.device)
    _check(load().lnrf_ngp_mlp_bwd_tc(_p(packed), L, m, _p(workspace), workspace.numel(), _p(dens), _p(rgb),
                                      _p(_f32c(d_dens, "d_dens")), _p(_f32c(d_rgb, "d_rgb")), _p(d_flat), _p(d_enc),
                                      _stream()), "lnrf_ngp_mlp_bwd_tc")


# --------------------------------------------------------------------------- rays / images
def bare_rays(origin, x_axis, y_axis, z, tan_half_x_fov, tan_half_y_fov, width, height, row0, rows, device):
    device = torch.device(device)
    ensure_init(device)
    with torch.cuda.device(device):
        rays = torch.empty(rows * width, 2, 3, device=device)
        _check(load().lnrf_bare_rays(_host3(origin), _host3(x_axis), _host3(y_axis), _host3(z),
                                     tan_half_x_fov, tan_half_y_fov, width, height, row0, rows, _p(rays),
                                     _stream()), "lnrf_bare_rays")
    return rays


def threefry_uniform(key0: int, key1: int, shape, device) -> torch.Tensor:
    device = torch.device(device)
    ensure_init(device)
    with torch.cuda.device(device):
        out = torch.empty(tuple(shape), device=device)
        _check(load().lnrf_threefry_uniform(key0, key1, out.numel(), _p(out), _stream()),
               "lnrf_threefry_uniform")
    return out


def rgb_to_u8(colors: torch.Tensor) -> torch.Tensor:
    colors = _f32c(colors.contiguous(), "colors")
    ensure_init(colors.device)
    out = torch.empty(colors.shape, dtype=torch.uint8, device=colors.device)
    _check(load().lnrf_rgb_to_u8(_p(colors), colors.numel(), _p(out), _stream()), "lnrf_rgb_to_u8")
    return out


# --------------------------------------------------------------------------- Ref-NeRF
def refnerf_param_count() -> int:
    return int(load().lnrf_refnerf_param_count())


def refnerf_param_floats() -> int:
    return int(load().lnrf_refnerf_param_floats())


def refnerf_param_offsets():
    buf = (c_int64 * 22)()
    _check(load().lnrf_refnerf_param_offsets(buf), "lnrf_refnerf_param_offsets")
    return [int(v) for v in buf]


def refnerf_workspace_bytes(m: int, save: bool) -> int:
    out = c_int64(0)
    _check(load().lnrf_refnerf_workspace_bytes(m, int(save), ctypes.byref(out)), "lnrf_refnerf_workspace_bytes")
    return int(out.value)


def refnerf_fwd(flat, x, d, rays, ts, n, T, save, workspace, dens, rgb, aux_mse, aux_neg):
    ensure_init(flat.device)
    _check(load().lnrf_refnerf_fwd(_p(flat), _p(x), _p(d), _p(rays), _p(ts), n, T, int(save), _p(workspace),
                                   workspace.numel(), _p(dens), _p(rgb), _p(aux_mse), _p(aux_neg), _stream()),
           "lnrf_refnerf_fwd")


def refnerf_bwd(flat, x, d, rays, ts, n, T, workspace, d_dens, d_rgb, d_aux_mse, d_aux_neg, d_flat):
    ensure_init(flat.device)
    _check(load().lnrf_refnerf_bwd(_p(flat), _p(x), _p(d), _p(rays), _p(ts), n, T, _p(workspace),
                                   workspace.numel(), _p(_f32c(d_dens, "d_dens")), _p(_f32c(d_rgb, "d_rgb")),
                                   _p(_f32c(d_aux_mse, "d_aux_normal_mse")),
                                   _p(_f32c(d_aux_neg, "d_aux_neg_normal")), _p(d_flat), _stream()),
           "lnrf_refnerf_bwd")


def ray_intervals(ts, t_min, t_max, want=("starts", "ends", "deltas")):
    """RaySamples.starts / ends / deltas (render.py:259-268) -> dict of [n,T] tensors."""
    ts = _f32c(ts, "ts")
    ensure_init(ts.device)
    n, T = ts.shape
    out = {k: torch.empty(n, T, device=ts.device) for k in want}
    _check(load().lnrf_ray_intervals(_p(ts), _p(_f32c(t_min, "t_min")), _p(_f32c(t_max, "t_max")), n, T,
                                     _p(out.get("starts")), _p(out.get("ends")), _p(out.get("deltas")), _stream()),
           "lnrf_ray_intervals")
    return out


def termination_probs(ts, t_min, t_max, dens):
    """RaySamples.termination_probs (render.py:270-287) -> [n,T+1]."""
    ts, dens = _f32c(ts, "ts"), _f32c(dens, "densities")
    ensure_init(ts.device)
    n, T = ts.shape
    probs = torch.empty(n, T + 1, device=ts.device)
    _check(load().lnrf_termination_probs(_p(ts), _p(_f32c(t_min, "t_min")), _p(_f32c(t_max, "t_max")), _p(dens), n, T,
                                         _p(probs), _stream()), "lnrf_termination_probs")
    return probs


def z_depth(coords, alphas, camera_origin, camera_direction, max_depth: float):
    """render_new_dataset.py:96-133 -> (z[n] in [0,1] fp32, depth[n] int32 holding the uint32 values)."""
    coords = _f32c(coords, "coords")
    alphas = _f32c(alphas.reshape(-1), "alphas")
    ensure_init(coords.device)
    n = coords.shape[0]
    z = torch.empty(n, device=coords.device)
    d32 = torch.empty(n, dtype=torch.int32, device=coords.device)
    _check(load().lnrf_z_depth(_p(coords), _p(alphas), _host3(camera_origin), _host3(camera_direction), float(max_depth),
                               n, _p(z), _p(d32), _stream()), "lnrf_z_depth")
    return z, d32


# --------------------------------------------------------------------------- device scoping
# The kernels launch on the CUDA runtime's CURRENT device and stream.  Every wrapper above runs inside
# the device of its first CUDA tensor argument, so TrainLoop(device="cuda:1") / render_view(device=...)
# work whatever the caller's current device is (and the stream is that device's current stream).
def _device_scoped(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapper


for _name in ("sample_coarse", "stratified", "sample_fine", "composite_fwd", "composite_bwd", "mse_loss",
              "nerf_pack_weights", "nerf_mlp_fwd", "nerf_mlp_bwd", "nerf_render_rays", "nerf_train_step", "adam_step", "adam_step_dk", "threefry_uniform_dk",
              "adam_step_peers", "debug_umma_gemm", "debug_umma_gemm_tn", "hashgrid_fwd", "hashgrid_bwd", "ngpref_fwd",
              "ngpref_bwd", "ngp_mlp_fwd", "ngp_mlp_bwd", "ngp_pack_weights", "ngp_mlp_fwd_tc", "ngp_mlp_bwd_tc", "rgb_to_u8", "refnerf_fwd", "refnerf_bwd", "ray_intervals",
              "termination_probs", "z_depth"):
    globals()[_name] = _device_scoped(globals()[_name])
