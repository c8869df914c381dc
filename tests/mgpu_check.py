"""Multi-GPU check of the data-parallel train step (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 tests/mgpu_check.py

1. Kernel check on fixed buffers: lnrf_adam_step_peers (fused peer all-reduce + Adam) against
   ncclAllReduce + lnrf_adam_step on the same per-rank gradients: parameters, moments, norms and
   the reduced loss sums must agree to a few ulp.
2. Two TrainLoops from the same weights, one per exchange path, on rank-sharded rays: logged
   losses agree, and with the peer path every rank holds bit-identical parameters after the
   steps (the replicas cannot drift).  The two loops' parameters are only compared loosely: the
   backward kernels accumulate with atomics, so two runs differ in the last bits of the gradient
   and Adam's first steps (lr * g / (|g| + eps)) amplify that for near-zero entries."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "learn-nerf_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from helpers import make_rays
    from learn_nerf import parallel
    from learn_nerf.model import NeRFModel
    from learn_nerf.train import TrainLoop

    def make_loop(mode, precision):
        os.environ["LNRF_ALLREDUCE"] = mode
        loop = TrainLoop(NeRFModel(precision=precision), NeRFModel(precision=precision), init_rng=3, lr=5e-4,
                         coarse_ts=64, fine_ts=128, device=dev)
        return loop

    from learn_nerf import _native
    # ---- 1. fused kernel vs NCCL + Adam on identical buffers
    count, extra = 1_187_851 + 1, 2
    pg = parallel.PeerGrads(count + extra + 1, dev)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    pg.buffer.copy_(torch.randn(count + extra + 1, device=dev, generator=gen) * 1e-3)
    gen0 = torch.Generator(device=dev).manual_seed(7)
    p0 = torch.randn(count, device=dev, generator=gen0)
    outs = []
    for mode in ("peer", "nccl"):
        p, m, v = p0.clone(), torch.zeros(count, device=dev), torch.zeros(count, device=dev)
        norms, sums = torch.zeros(2, device=dev), torch.zeros(2, device=dev)
        for step in (1, 2):
            if mode == "peer":
                pg.barrier()
                _native.adam_step_peers(p, pg.ptrs, m, v, count, extra, 1e-3, 0.9, 0.999, 1e-7, step, 1.0 / world,
                                        norms, sums)
                pg.barrier()
            else:
                g = pg.buffer.clone()
                dist.all_reduce(g)
                sums.copy_(g[count:count + extra])
                _native.adam_step(p, g[:count], m, v, 1e-3, 0.9, 0.999, 1e-7, step, 1.0 / world, norms)
        torch.cuda.synchronize()
        outs.append((p, m, v, norms, sums))
    for name, x, y in zip(("params", "m", "v", "norms", "loss sums"), outs[0], outs[1]):
        d = float((x - y).abs().max())
        # a few ulp: the two kernels contract FMAs differently; the norms are atomic float sums
        tol = (1e-5 if name == "norms" else 1e-6) * max(1.0, float(y.abs().max()))
        assert d <= tol, f"fused peer all-reduce + Adam vs NCCL: {name} differ by {d}"
    if rank == 0:
        print(f"MGPU_KERNEL_OK world={world}")

    n = 512 * world
    batch = torch.from_numpy(make_rays(n, seed=11)).to(dev)
    a, b = parallel.shard_bounds(n, rank, world)
    for precision in ("fp32", "bf16"):
        peer, nccl = make_loop("peer", precision), make_loop("nccl", precision)
        assert peer._peers is not None, "peer mapping unavailable on this box"
        assert nccl._peers is None
        nccl.state.flat.copy_(peer.state.flat)
        for name in ("coarse", "fine"):
            nccl.state.params[name].mark_updated()
        sp, sn = peer.step_fn([-1.0] * 3, [1.0] * 3), nccl.step_fn([-1.0] * 3, [1.0] * 3)
        for i in range(3):
            lp, ln = sp(100 + i, batch[a:b]), sn(100 + i, batch[a:b])
            for k in lp:
                assert abs(float(lp[k]) - float(ln[k])) <= 1e-4 * max(1.0, abs(float(ln[k]))), (k, float(lp[k]), float(ln[k]))
        torch.cuda.synchronize()
        diff = float((peer.state.flat - nccl.state.flat).abs().max())
        assert diff <= 2 * 3 * 5e-4, f"{precision}: peer vs NCCL parameters differ by {diff} (> 2 * steps * lr)"
        # replicas identical on every rank
        ref = peer.state.flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, peer.state.flat), f"{precision}: rank {rank} drifted from rank 0"
        if rank == 0:
            print(f"MGPU_OK {precision} world={world} max|peer-nccl|={diff:.3g} fine_loss={float(lp['fine']):.6f}")
    # ---- 3. the CUDA-graph step on the peer path (cross-rank barriers and the fused kernel inside the graph)
    if True:
        os.environ["LNRF_ALLREDUCE"] = "peer"
        mk = lambda graph: TrainLoop(NeRFModel(precision="bf16"), NeRFModel(precision="bf16"), init_rng=3, lr=5e-4,
                                     coarse_ts=64, fine_ts=128, device=dev, cuda_graph=graph)
        gl, el = mk(True), mk(False)
        sg, se = gl.step_fn([-1.0] * 3, [1.0] * 3), el.step_fn([-1.0] * 3, [1.0] * 3)
        for i in range(4):
            lg, le = sg(300 + i, batch[a:b]), se(300 + i, batch[a:b])
        for k in lg:
            assert abs(float(lg[k]) - float(le[k])) <= 2e-3 * max(1.0, abs(float(le[k]))), (k, float(lg[k]), float(le[k]))
        assert gl._cg is not None, "the graph path was not taken"
        torch.cuda.synchronize()
        ref = gl.state.flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, gl.state.flat), f"graph path: rank {rank} drifted from rank 0"
        if rank == 0:
            print(f"MGPU_GRAPH_OK world={world} fine_loss={float(lg['fine']):.6f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
