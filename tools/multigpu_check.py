"""2-rank NCCL parity check of the sharded PPO path (run under torch.distributed.run, one rank per GPU):

  * each rank rolls out N envs with global ids [rank*N, (rank+1)*N) -> the concatenation equals a single-GPU
    rollout of 2N envs (trajectories do not depend on the GPU count);
  * one optimizer step with the gradient sum-all-reduce (+ global advantage normalisation) leaves every rank
    with the same parameters as the single-GPU step on the concatenated batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multigpu_check.py
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def make_trainer(P, n_envs, graph=False, n_mb=1, epochs=1, peer_reduce=True):
    cfg = P.Config()
    cfg.peer_reduce = peer_reduce
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.num_epochs, cfg.seed, cfg.use_cuda_graph = n_envs, 16, n_mb, epochs, 11, graph
    torch.manual_seed(0)
    return P.PPOTrainer(cfg)


def graph_vs_eager(P, N):
    """The epoch graph with the captured NCCL all-reduce must leave the parameters of the eager multi-GPU path
    (2 epochs x 2 minibatches; rollout buffers are made identical by construction: same seeds, same shard)."""
    flats = []
    for graph in (False, True):
        tr = make_trainer(P, N, graph=graph, n_mb=2, epochs=2)
        tr.cfg.use_cuda_graph = False           # eager rollout in both, so the two runs see identical buffers
        tr.collect_rollout()
        tr.cfg.use_cuda_graph = graph
        tr.update(None)
        if graph:
            assert tr._g_epoch is not None, "the distributed epoch graph was not captured"
        flats.append(tr.net.fp.flat.clone())
    return (flats[0] - flats[1]).abs().max().item()


def peer_vs_nccl(P, N):
    """The one-shot peer-memory reduction (csrc/comm.cu) against ncclAllReduce: same rollout, 2 epochs x 2 minibatches."""
    flats, used = [], []
    for peer in (False, True):
        tr = make_trainer(P, N, graph=False, n_mb=2, epochs=2, peer_reduce=peer)
        tr.collect_rollout()
        tr.update(None)
        used.append(tr.comm is not None)
        flats.append(tr.net.fp.flat.clone())
    return (flats[0] - flats[1]).abs().max().item(), used


def run_twice(P, N):
    """Two runs of the same seeded multi-GPU iterations (peer reduction, epoch graph) leave bit-identical parameters: the
    reduction adds the ranks in rank order and every other cross-block sum folds in a fixed order."""
    flats = []
    for _ in range(2):
        tr = make_trainer(P, N, graph=True, n_mb=2, epochs=2)
        for _ in range(2):
            tr.collect_rollout()
            tr.update(None)
        flats.append(tr.net.fp.flat.clone())
    return bool(torch.equal(flats[0], flats[1]))


def reduce_latency(n_floats=200968, iters=300):
    """Device time of one gradient reduction, NCCL vs peer-memory kernel (+ the norm pass NCCL needs afterwards)."""
    import gymrl_b200.dist as gd
    from gymrl_b200 import ops
    g = torch.randn(n_floats, device="cuda")
    red = torch.zeros_like(g)
    part = torch.zeros(4096, device="cuda", dtype=torch.float64)
    sumsq = torch.zeros(1, device="cuda", dtype=torch.float64)
    comm = gd.make_peer_reducer(n_floats)
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name in ("nccl_allreduce_plus_sumsq", "peer_reduce_sumsq"):
        if name.startswith("peer") and comm is None:
            continue
        def once():
            if name.startswith("peer"):
                comm.allreduce_sumsq(g, red, part)
            else:
                dist.all_reduce(g, op=dist.ReduceOp.SUM)
                ops.grad_sumsq(g, out=sumsq)
        for _ in range(20):
            once()
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            once()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) * 1e3 / iters
    if comm is not None:
        # correctness on random data: equals the NCCL sum (world = 2: one addition, order-free) and the partials give its norm
        a = torch.randn(n_floats, device="cuda")
        ref = a.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        n = comm.allreduce_sumsq(a, red, part)
        torch.cuda.synchronize()
        out["max_abs_diff_vs_nccl"] = (red - ref).abs().max().item()
        out["rel_norm_err"] = abs(part[:n].sum().item() - ref.double().pow(2).sum().item()) / ref.double().pow(2).sum().item()
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gymrl_b200.algorithms import ppo_lunarlander as P
    N = 256
    tr = make_trainer(P, N)
    tr.collect_rollout()
    tr.update(None)
    flat = tr.net.fp.flat.clone()
    # every rank must hold identical parameters after the step
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([float(torch.equal(ref, flat))], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    # gather shards on rank 0
    obs = [torch.zeros_like(tr.buffer.obs) for _ in range(world)]
    act = [torch.zeros_like(tr.buffer.action) for _ in range(world)]
    dist.all_gather(obs, tr.buffer.obs)
    dist.all_gather(act, tr.buffer.action)
    ok = True
    if rank == 0:
        # single-GPU run of the concatenated problem, outside the process group's view
        import gymrl_b200.dist as gd
        real_info = gd.info
        gd.info = lambda: (0, 1)
        try:
            one = make_trainer(P, N * world)
            one.collect_rollout()
            one.update(None)
        finally:
            gd.info = real_info
        cat_obs, cat_act = torch.cat(obs, dim=1), torch.cat(act, dim=1)
        e_obs = bool(torch.equal(cat_obs, one.buffer.obs))
        e_act = bool(torch.equal(cat_act, one.buffer.action))
        dpar = (one.net.fp.flat - flat).abs().max().item()
        scale = one.net.fp.flat.abs().max().item()
        print(f"ranks identical after step: {bool(same.item())}; shard rollout == single-GPU rollout: obs {e_obs}, actions {e_act}; "
              f"max |param diff| vs single GPU on the concatenated batch: {dpar:.3e} (param scale {scale:.2f})")
        ok = bool(same.item()) and e_obs and e_act and dpar < 2e-5
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL")
    dgraph = graph_vs_eager(P, N)
    if rank == 0:
        print(f"epoch graph (captured all-reduce) vs eager multi-GPU update: max |param diff| {dgraph:.3e}")
        ok = ok and dgraph < 1e-6
        print("MULTIGPU_GRAPH_CHECK", "PASS" if dgraph < 1e-6 else "FAIL")
    dpeer, used = peer_vs_nccl(P, N)
    same2 = torch.tensor([float(run_twice(P, N))], device="cuda")
    dist.all_reduce(same2, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_DETERMINISM_CHECK", "PASS" if same2.item() >= 1 else "FAIL", "(two seeded runs, bitwise equal parameters on every rank)")
        ok = ok and same2.item() >= 1
    lat = reduce_latency()
    if rank == 0:
        print(f"peer-memory reduction vs ncclAllReduce update: max |param diff| {dpeer:.3e} (peer path active: {used[1]}, nccl run: {not used[0]})")
        print("reduce latency (us per optimizer step, device events):", {k: round(v, 3) if isinstance(v, float) else v for k, v in lat.items()})
        okp = dpeer < 1e-6 and (not used[1] or (lat.get("max_abs_diff_vs_nccl", 1.0) <= (0.0 if world == 2 else 1e-5) and lat.get("rel_norm_err", 1.0) < 1e-12))
        print("MULTIGPU_PEER_CHECK", "PASS" if okp else "FAIL")
        ok = ok and okp
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
