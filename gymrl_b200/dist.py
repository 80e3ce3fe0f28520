"""Multi-GPU host logic: one process per GPU, env shards by global env id, ONE sum-all-reduce of
the flat gradient per optimizer step (+ a 3-scalar all-reduce for the global advantage normalisation).
Backend-agnostic on purpose: NCCL over NVLink on the GPU box, gloo in the CPU tests (tests/test_dist_cpu.py).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def info() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(n_local: int, rank: int) -> Tuple[int, int]:
    """Global env-id range [first, first + n_local) owned by `rank` (weak scaling: n_local envs per GPU)."""
    return rank * n_local, n_local


def broadcast_module_(module: torch.nn.Module, device=None, src: int = 0) -> None:
    """Identical replicas: every rank starts from rank `src`'s initialisation."""
    rank, world = info()
    if world == 1:
        return
    for p in module.parameters():
        t = p.data.to(device) if device is not None else p.data
        dist.broadcast(t, src)
        p.data = t


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    rank, world = info()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def global_moments_(sums: torch.Tensor, local_count: int) -> float:
    """sums = [sum, sumsq, (count)] float64.  All-reduces in place and returns the global element count, so that
    mean/std of the concatenated shards come out of gymrl_normalize_inplace (numpy ddof = 0 semantics, ref :236)."""
    rank, world = info()
    if world == 1:
        return float(local_count)
    sums[2] = float(local_count)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return float(sums[2].item())


def allreduce_scalars(values, device=None):
    """Sum a short list of host floats over all ranks (float64); identity when not distributed.  Every rank must call it the
    same number of times — used for decisions that have to be rank-symmetric (e.g. PPOTrainer.train()'s stop criterion: a rank
    that left the loop alone would leave its peers blocked in the next gradient all-reduce)."""
    rank, world = info()
    if world == 1:
        return [float(v) for v in values]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def grad_scale() -> float:
    """Gradients are summed across ranks; Adam rescales by 1/world so the step equals the single-GPU step on the
    concatenated minibatch (each rank's loss is a mean over its local minibatch)."""
    return 1.0 / info()[1]
