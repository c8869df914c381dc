"""Host-side mirror of learn_nerf/train.py: TrainLoop (losses, step_fn, save/load).

Same constructor arguments and logging-dict keys as the reference (train.py:22-165).
One step = K1 sampling -> K2 MLP fwd (coarse) -> K3 composite -> K4 fine sampling ->
K2 (fine) -> K3 -> MSE -> K5 composite bwd -> K6 MLP bwd (both levels) ->
[NCCL all-reduce of the flat gradient when world > 1] -> K10 fused Adam + norms.
All parameters, gradients and Adam moments live in single flat fp32 buffers
``[coarse | fine | background]`` so the optimiser and the all-reduce are one launch.
"""
import math
import os
import pickle
from typing import Any, Callable, Dict, Optional

import numpy as np
import torch

from . import _native, parallel, prng
from .model import ModelBase
from .render import NeRFRenderer, RaySamples, _pack_aux, _vec3


class TrainState:
    """Stand-in for flax TrainState (train.py:51-60): params tree + Adam moments + step."""

    def __init__(self, params: Dict[str, Any], flat: torch.Tensor):
        self.params = params
        self.flat = flat
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.step = 0


def default_loss_weights() -> Dict[str, float]:  # train.py:187-191
    return dict(normal_mse=3e-4, neg_normal=0.1)


class TrainLoop:
    """A stateful training loop (train.py:17-60)."""

    def __init__(self, coarse: ModelBase, fine: ModelBase, init_rng, lr: float, coarse_ts: int,
                 fine_ts: int, adam_b1: float = 0.9, adam_b2: float = 0.999, adam_eps: float = 1e-7,
                 loss_weights: Dict[str, float] = None, density_penalty: Optional[float] = None,
                 density_penalty_batch_size: int = 128, device=None, ray_chunk: Optional[int] = None,
                 cuda_graph: bool = False):
        self.coarse, self.fine = coarse, fine
        self.coarse_ts, self.fine_ts = coarse_ts, fine_ts
        self.lr, self.b1, self.b2, self.eps = lr, adam_b1, adam_b2, adam_eps
        self.loss_weights = loss_weights if loss_weights is not None else default_loss_weights()
        self.density_penalty = density_penalty
        self.density_penalty_batch_size = density_penalty_batch_size
        self.ray_chunk = ray_chunk
        # cuda_graph: capture the whole step (19 launches for NeRF) once per batch size and replay it;
        # the per-step PRNG keys and Adam bias corrections are uploaded to device memory before each
        # replay.  PRNG-key entry point, no density penalty, one GPU or the peer gradient exchange;
        # anything else runs eagerly.
        self.cuda_graph = cuda_graph
        self._cg = None
        device = torch.device(device or "cuda")
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device

        coarse_rng, fine_rng = prng.split(init_rng)  # :46
        nc, nf = coarse.param_floats(), fine.param_floats()
        flat = torch.zeros(nc + nf + 4, device=device)
        coarse_vars = coarse.init(dict(params=coarse_rng), device=device, flat=flat[:nc])
        fine_vars = fine.init(dict(params=fine_rng), device=device, flat=flat[nc:nc + nf])
        flat[nc + nf: nc + nf + 3] = -1.0  # background starts black (:56-57) and is trained
        self._slices = dict(coarse=(0, nc), fine=(nc, nc + nf), background=(nc + nf, nc + nf + 3))
        self.state = TrainState(
            params=dict(coarse=coarse_vars["params"], fine=fine_vars["params"],
                        background=flat[nc + nf: nc + nf + 3]),
            flat=flat)
        # gradients: same layout as the params, followed by the two per-rank loss sums so that ONE
        # exchange per step covers everything that is summed over ranks
        self._n_params = nc + nf + 4
        self._peers = parallel.peer_grads(self._n_params + 4, device)  # None: single GPU / NCCL path
        self._grads_all = self._peers.buffer if self._peers is not None else torch.zeros(
            self._n_params + 4, device=device)
        self._grads = self._grads_all[:self._n_params]
        self._loss_sums = self._grads_all[self._n_params:self._n_params + 2]
        self._scalars = torch.zeros(4, device=device)  # loss_c, loss_f, |g|^2, |p|^2
        self._zero3 = torch.zeros(3, device=device)
        self._scratch3 = torch.zeros(3, device=device)

    # ------------------------------------------------------------------ checkpoints
    def save(self, path: str):
        """train.py:62-69: pickle of the params tree (numpy leaves), atomic tmp + rename."""
        def to_np(t):
            if isinstance(t, dict):
                return {k: to_np(v) for k, v in t.items()}
            return t.detach().cpu().numpy().copy()
        tmp_path = path + ".tmp"
        with open(tmp_path, "wb") as f:
            pickle.dump(to_np(self.state.params), f)
        os.rename(tmp_path, path)

    def load(self, path: str):
        """train.py:71-76: replaces params only (Adam moments restart, as in the reference)."""
        with open(path, "rb") as f:
            tree = pickle.load(f)
        def put(dst, src):
            if isinstance(dst, dict):
                for k in dst:
                    put(dst[k], src[k])
            else:
                dst.copy_(torch.as_tensor(np.asarray(src), dtype=torch.float32))
        put(self.state.params, tree)
        for name in ("coarse", "fine"):
            self.state.params[name].mark_updated()

    # ------------------------------------------------------------------ step
    def step_fn(self, bbox_min, bbox_max) -> Callable[[Any, torch.Tensor], Dict[str, torch.Tensor]]:
        """train.py:78-112: returns ``step(key, batch[N,3,3]) -> logging dict`` (in place)."""
        bmin, bmax = _vec3(bbox_min), _vec3(bbox_max)

        def in_place_step(key, batch: torch.Tensor) -> Dict[str, torch.Tensor]:
            if self.cuda_graph and self._graphable(key, batch):
                return self._step_graphed(key, bmin, bmax, batch)
            return self._step(key, bmin, bmax, batch)

        return in_place_step

    def _renderer(self, bmin, bmax, params) -> NeRFRenderer:
        return NeRFRenderer(coarse=self.coarse, fine=self.fine, coarse_params=params["coarse"],
                            fine_params=params["fine"], background=params["background"],
                            bbox_min=bmin, bbox_max=bmax, coarse_ts=self.coarse_ts,
                            fine_ts=self.fine_ts)

    # ------------------------------------------------------------------ density penalty
    def _density_points(self, density_key, bmin, bmax):
        """train.py:174-180: coords = U[0,1)^3 * (bbox_max - bbox_min) + bbox_min and unit
        directions, both drawn from ``density_key``.  ``density_key`` may also be an explicit
        ``(coords[B,3], dirs[B,3])`` pair (the parity entry point).  The directions only reach
        the colour head, so they never influence the penalty or its gradient."""
        if isinstance(density_key, (tuple, list)) and isinstance(density_key[0], torch.Tensor):
            coords, dirs = density_key
            return (_native._f32c(coords.contiguous(), "coords"), _native._f32c(dirs.contiguous(), "dirs"))
        bs = self.density_penalty_batch_size
        u = prng.uniform(density_key, (bs, 3), self.device)
        lo = torch.tensor(bmin, device=self.device, dtype=torch.float32)
        hi = torch.tensor(bmax, device=self.device, dtype=torch.float32)
        coords = u * (hi - lo) + lo
        # jax.random.normal(key) = sqrt(2) * erfinv(uniform(key, minval=nextafter(-1, 0), maxval=1))
        lo1 = float(np.nextafter(np.float32(-1.0), np.float32(0.0)))
        z = math.sqrt(2.0) * torch.erfinv(torch.clamp(u * (1.0 - lo1) + lo1, min=lo1))
        dirs = z / z.norm(dim=-1, keepdim=True)
        return coords.contiguous(), dirs.contiguous()

    def average_density(self, key, model: ModelBase, params, bbox_min, bbox_max, _save=False):
        """train.py:166-184: mean density of ``model`` at random points of the bounding box.
        The points go through the renderer's fused seam as zero-length rays (x = o + d * 0)."""
        coords, dirs = self._density_points(key, _vec3(bbox_min), _vec3(bbox_max))
        rays = torch.stack([coords, dirs], dim=1).contiguous()
        ts = torch.zeros(coords.shape[0], 1, device=coords.device)
        dens, rgb, _, ctx = model.apply_rays(params, rays, ts, save=_save, slot="density_penalty")
        mean = dens.mean()
        return (mean, dens, rgb, ctx) if _save else mean

    @staticmethod
    def _split_key(key, n, tc, tf, a, b):
        """Per-chunk key: explicit uniforms are sliced, PRNG keys are re-split."""
        if isinstance(key, (tuple, list)) and isinstance(key[0], torch.Tensor):
            return (key[0][a:b].contiguous(), key[1][a:b].contiguous())
        return key if (a == 0 and b == n) else prng.fold_in(key, a)

    # ------------------------------------------------------------------ CUDA-graph replay of the step
    def _graphable(self, key, batch) -> bool:
        # multi-GPU: only with the peer exchange (its barriers and the fused kernel are plain launches)
        multi_ok = parallel.world()[1] == 1 or self._peers is not None
        return (multi_ok and self.density_penalty is None and not isinstance(key, (tuple, list))
                and (self.ray_chunk is None or self.ray_chunk >= batch.shape[0]) and batch.shape[0] > 0)

    def _step_graphed(self, key, bmin, bmax, batch: torch.Tensor) -> Dict[str, torch.Tensor]:
        st = self.state
        n = batch.shape[0]
        cg = self._cg
        if cg is None or cg["n"] != n or cg["bbox"] != (tuple(bmin), tuple(bmax)):
            cg = self._cg = self._capture(n, bmin, bmax)
        # host side of the step: key bookkeeping (train.py:137, render.py:55) and Adam's step count
        key, _density_key = prng.split(key)
        kc, kf = prng.split(key)
        st.step += 1
        scalars = np.empty(6, dtype=np.int32)
        scalars[:4] = np.array([kc.k0, kc.k1, kf.k0, kf.k1], dtype=np.uint32).view(np.int32)
        scalars[4:6] = np.array([1.0 / (1.0 - self.b1 ** st.step), 1.0 / (1.0 - self.b2 ** st.step)],
                                dtype=np.float32).view(np.int32)
        cg["ring"].upload(scalars, cg["dev_scalars"])  # pinned ring: asynchronous and race-free
        cg["batch"].copy_(batch, non_blocking=True)
        cg["graph"].replay()
        for name in ("coarse", "fine"):  # eager users (losses(), a renderer) must re-pack the new weights
            st.params[name].mark_updated()
        return cg["logs"]

    def _capture(self, n: int, bmin, bmax):
        st = self.state
        dev = self.device
        dev_scalars = torch.zeros(6, dtype=torch.int32, device=dev)  # 4 key words | 2 fp32 bias corrections
        cg = dict(n=n, bbox=(tuple(bmin), tuple(bmax)), dev_scalars=dev_scalars, ring=_native.PinnedRing(6),
                  batch=torch.zeros(n, 3, 3, device=dev))
        keys = (prng.DeviceKey(dev_scalars[0:2]), prng.DeviceKey(dev_scalars[2:4]))
        bc_dev = dev_scalars[4:6].view(torch.float32)
        # warm-up on a side stream (lazy allocations, workspace caches), then capture; neither may
        # leave a trace in the training state
        saved = (st.flat.clone(), st.m.clone(), st.v.clone())
        dev_scalars[4:6] = torch.tensor([1.0, 1.0]).view(torch.int32).to(dev)  # valid corrections for the warm-up
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step(keys, bmin, bmax, cg["batch"], _graph_bc=bc_dev)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            cg["logs"] = self._step(keys, bmin, bmax, cg["batch"], _graph_bc=bc_dev)
        for dst, src in zip((st.flat, st.m, st.v), saved):
            dst.copy_(src)
        for name in ("coarse", "fine"):
            st.params[name].mark_updated()
        cg["graph"] = graph
        return cg

    def _step(self, key, bmin, bmax, batch: torch.Tensor, _graph_bc=None) -> Dict[str, torch.Tensor]:
        st = self.state
        batch = _native._f32c(batch.contiguous(), "batch")
        n = batch.shape[0]
        world = parallel.world()[1]
        density_key = None
        if isinstance(key, (tuple, list)) and len(key) == 3:  # (u_coarse, u_fine, (coords, dirs))
            key, density_key = (key[0], key[1]), key[2]
        elif not isinstance(key, (tuple, list)):
            key, density_key = prng.split(key)  # train.py:137
        g = self._grads
        self._grads_all.zero_()
        self._scalars.zero_()
        aux_sums: Dict[str, torch.Tensor] = {}
        renderer = self._renderer(bmin, bmax, st.params)
        inv_count = 1.0 / (3.0 * n)  # jnp.mean over N*3 (:141-142)
        chunk = self.ray_chunk or n
        sc, sf, sb = self._slices["coarse"], self._slices["fine"], self._slices["background"]
        for a in range(0, n, chunk):
            b = min(a + chunk, n)
            sub = batch[a:b]
            rays = sub[:, :2].contiguous()
            out = renderer.render_rays(self._split_key(key, n, self.coarse_ts, self.fine_ts, a, b),
                                       rays, _save=True)
            targets = sub[:, 2]  # strided view: element (i,c) at base + 9 i + c
            for li, (level, model, sl) in enumerate((("coarse", self.coarse, sc),
                                                     ("fine", self.fine, sf))):
                lv = out[level]
                d_out = torch.empty_like(lv["outputs"])
                _native.mse_loss(lv["outputs"], targets, 9, b - a, inv_count,
                                 self._loss_sums[li:li + 1], d_out)
                ts: RaySamples = lv["_ts"]
                d_dens, d_rgb = _native.composite_bwd(ts.ts, ts.t_min, ts.t_max, ts._mask_u8(),
                                                      lv["densities"], lv["rgbs"],
                                                      st.params["background"], d_out,
                                                      g[sb[0]:sb[0] + 3])
                d_aux = None
                if lv["_aux"]:
                    # aux losses (:146-151): total += w * mean_rays(where(mask, sum_t v_t p_t, 0)).
                    # Their gradient w.r.t. the densities and the aux values is the compositing
                    # gradient with the aux values as colours, a zero background and d_out = w / N.
                    names, cols = _pack_aux(lv["_aux"])
                    d_comp = torch.zeros(b - a, 3, device=batch.device)
                    for i, name in enumerate(names):
                        d_comp[:, i] = self.loss_weights[name] / n
                        key_name = f"{level}_{name}"
                        aux_sums[key_name] = aux_sums.get(key_name, 0.0) + out[f"{level}_aux"][name] * (b - a)
                    d_dens_aux, d_cols = _native.composite_bwd(ts.ts, ts.t_min, ts.t_max, ts._mask_u8(),
                                                               lv["densities"], cols, self._zero3, d_comp,
                                                               self._scratch3)
                    d_dens = d_dens + d_dens_aux
                    d_aux = {name: d_cols[..., i].contiguous() for i, name in enumerate(names)}
                model.backward_rays(lv["_ctx"], d_dens, d_rgb, g[sl[0]:sl[1]], d_aux=d_aux)
        penalties: Dict[str, torch.Tensor] = {}
        if self.density_penalty is not None:
            # train.py:153-163: total += density_penalty * mean(density(model, random points)),
            # fine model first.  Every rank draws the same points, so after the all-reduce and
            # Adam's 1/world the gradient is that of the single-device loss.
            if density_key is None:
                raise ValueError("density_penalty needs a PRNG key or explicit (coords, dirs) points")
            for prefix, model, sl in (("fine", self.fine, sf), ("coarse", self.coarse, sc)):
                mean, dens, rgb, ctx = self.average_density(density_key, model, st.params[prefix], bmin, bmax,
                                                            _save=True)
                penalties[f"{prefix}_density"] = mean
                d_dens = torch.full_like(dens, self.density_penalty / dens.numel())
                model.backward_rays(ctx, d_dens, torch.zeros_like(rgb), g[sl[0]:sl[1]])
        if _graph_bc is None:
            st.step += 1
        if _graph_bc is not None and self._peers is None:  # graph capture / replay: the caller owns the step count
            self._scalars[0:2].copy_(self._loss_sums)
            _native.adam_step_dk(st.flat, g, st.m, st.v, self.lr, self.b1, self.b2, self.eps, _graph_bc,
                                 1.0 / world, self._scalars[2:4])
        elif self._peers is not None:
            # fused all-reduce + Adam: every rank reads all ranks' gradients over NVLink inside the
            # optimiser kernel (rank-order sum, 1/world folded in); barriers fence the peer reads
            self._peers.barrier()
            _native.adam_step_peers(st.flat, self._peers.ptrs, st.m, st.v, self._n_params, 2, self.lr,
                                    self.b1, self.b2, self.eps, max(st.step, 1), 1.0 / world, self._scalars[2:4],
                                    self._scalars[0:2], inv_bias_corr_dev=_graph_bc)
            self._peers.barrier()
        else:
            if world > 1 and os.environ.get("LNRF_ALLREDUCE") != "none":  # one NCCL sum (gradients + loss sums); 1/world is folded into Adam
                parallel.allreduce_sum_(self._grads_all)
            self._scalars[0:2].copy_(self._loss_sums)
            _native.adam_step(st.flat, g, st.m, st.v, self.lr, self.b1, self.b2, self.eps, st.step,
                              1.0 / world, self._scalars[2:4])
        for name in ("coarse", "fine"):
            st.params[name].mark_updated()
        s = self._scalars
        scale = inv_count / world
        logs = dict(coarse=s[0] * scale, fine=s[1] * scale)
        if aux_sums:
            names = sorted(aux_sums)
            vals = torch.stack([aux_sums[k] for k in names]) / n
            if world > 1:
                parallel.mean_scalars_(vals)
            for i, k in enumerate(names):
                logs[k] = vals[i]
        logs.update(penalties)
        logs["grad_norm"] = torch.sqrt(s[2])
        logs["param_norm"] = torch.sqrt(s[3])
        return logs

    def losses(self, key, bbox_min, bbox_max, batch: torch.Tensor, params):
        """train.py:114-165 (forward only) -> (total_loss, loss_dict)."""
        batch = _native._f32c(batch.contiguous(), "batch")
        n = batch.shape[0]
        density_key = None
        if isinstance(key, (tuple, list)) and len(key) == 3:
            key, density_key = (key[0], key[1]), key[2]
        elif not isinstance(key, (tuple, list)):
            key, density_key = prng.split(key)
        renderer = self._renderer(_vec3(bbox_min), _vec3(bbox_max), params)
        out = renderer.render_rays(key, batch[:, :2].contiguous())
        sums = torch.zeros(2, device=batch.device)
        for li, level in enumerate(("coarse", "fine")):
            _native.mse_loss(out[level]["outputs"], batch[:, 2], 9, n, 0.0, sums[li:li + 1], None)
        loss_dict = dict(coarse=sums[0] / (3.0 * n), fine=sums[1] / (3.0 * n))
        total = loss_dict["coarse"] + loss_dict["fine"]
        for prefix in ("coarse", "fine"):
            for name, loss in out[f"{prefix}_aux"].items():
                loss_dict[f"{prefix}_{name}"] = loss
                total = total + self.loss_weights[name] * loss
        if self.density_penalty is not None:  # :153-163
            for prefix, model in (("fine", self.fine), ("coarse", self.coarse)):
                penalty = self.average_density(density_key, model, params[prefix], bbox_min, bbox_max)
                loss_dict[f"{prefix}_density"] = penalty
                total = total + self.density_penalty * penalty
        return total, loss_dict
