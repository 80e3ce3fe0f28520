"""world_size = 2 on CPU (gloo): the host-side multi-GPU logic of the path (gymrl_b200/dist.py) — shard ranges,
gradient sum-all-reduce + 1/world rescale == gradient of the concatenated minibatch, global advantage moments ==
NumPy over the concatenated shards (ddof = 0, algorithms/ppo_lunarlander.py:236)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gymrl_b200 import dist as gd
    assert gd.info() == (rank, world)
    first, n = gd.shard(4096, rank)
    # (1) gradients: each rank's loss is a mean over its local half of the batch
    torch.manual_seed(0)
    w = torch.randn(16, 8, requires_grad=True)
    x_all, y_all = torch.randn(64, 8), torch.randn(64, 16)
    half = slice(rank * 32, (rank + 1) * 32)
    loss = ((x_all[half] @ w.T - y_all[half]) ** 2).mean()
    loss.backward()
    g = w.grad.clone()
    gd.allreduce_sum_(g)
    g *= gd.grad_scale()
    w2 = w.detach().clone().requires_grad_(True)
    ((x_all @ w2.T - y_all) ** 2).mean().backward()
    # (2) advantage moments
    rng = np.random.default_rng(1)
    adv_all = rng.standard_normal(2000)
    mine = adv_all[rank * 1000:(rank + 1) * 1000]
    sums = torch.tensor([mine.sum(), (mine ** 2).sum(), 0.0], dtype=torch.float64)
    count = gd.global_moments_(sums, len(mine))
    mean = sums[0].item() / count
    std = np.sqrt(max(sums[1].item() / count - mean * mean, 0.0))
    # (3) replicas
    m = torch.nn.Linear(4, 4)
    gd.broadcast_module_(m)
    chk = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    gathered = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(gathered, chk)
    # (4) rank-symmetric stop decision: rank 0 alone is "solved"; the pooled last-100 average decides for both
    avg_r, total_r = (250.0, 120) if rank == 0 else (100.0, 40)
    k = float(min(total_r, 100))
    s_, k_all, tot = gd.allreduce_scalars([avg_r * k, k, float(total_r)])
    out[rank] = dict(pooled_avg=s_ / k_all, pooled_total=int(tot), first=first, n=n, grad_err=float((g - w2.grad).abs().max()), count=count,
                     mean_err=abs(mean - adv_all.mean()), std_err=abs(std - adv_all.std()),
                     replicas_equal=bool(torch.equal(gathered[0], gathered[1])))
    dist.destroy_process_group()


def test_two_rank_host_logic_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]
    assert (out[0]["first"], out[0]["n"]) == (0, 4096) and (out[1]["first"], out[1]["n"]) == (4096, 4096)
    for r in (0, 1):
        assert out[r]["grad_err"] < 1e-6          # sum / world == gradient of the concatenated batch
        assert out[r]["count"] == 2000.0
        assert out[r]["mean_err"] < 1e-12 and out[r]["std_err"] < 1e-12
        assert out[r]["replicas_equal"]
        assert out[r]["pooled_total"] == 160 and abs(out[r]["pooled_avg"] - (250.0 * 100 + 100.0 * 40) / 140) < 1e-12
