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


class PeerReducer:
    """One-shot NVLink peer-memory gradient reduction fused with the clip's sum of squares (csrc/comm.cu, include/gymrl.h
    gymrl_comm_*): replaces the ncclAllReduce + separate norm pass of round 1 on the per-minibatch path.  torch.distributed is
    only the plumbing that exchanges the cudaIpc handles once at construction (make_peer_reducer)."""

    def __init__(self, n_floats: int, n_blocks: int = 0):
        import ctypes as C

        from . import _ffi
        rank, world = info()
        self._lib, self._h = _ffi.load(), C.c_void_p()
        _ffi.check(self._lib.gymrl_comm_create(C.byref(self._h), rank, world, int(n_floats), int(n_blocks)))
        self.n_partials = int(self._lib.gymrl_comm_n_partials(self._h))
        self.n_floats = int(n_floats)

    def handle(self) -> bytes:
        import ctypes as C

        from . import _ffi
        buf = (C.c_ubyte * self._lib.gymrl_comm_handle_bytes())()
        _ffi.check(self._lib.gymrl_comm_get_handle(self._h, buf))
        return bytes(buf)

    def open(self, handles: bytes):
        from . import _ffi
        _ffi.check(self._lib.gymrl_comm_open(self._h, handles))

    def allreduce_sumsq(self, grad: torch.Tensor, reduced: torch.Tensor, sumsq_partials: torch.Tensor) -> int:
        """reduced = sum over ranks of grad (bit-identical on every rank); sumsq_partials[:n] = per-block sums of squares of it."""
        from . import _ffi
        assert grad.numel() == self.n_floats == reduced.numel() and sumsq_partials.numel() >= self.n_partials
        _ffi.check(self._lib.gymrl_comm_allreduce_sumsq(self._h, _ffi.ptr(grad, torch.float32), _ffi.ptr(reduced, torch.float32),
                                                        _ffi.ptr(sumsq_partials, torch.float64), _ffi.stream_ptr()))
        return self.n_partials

    def close(self):
        if self._h:
            self._lib.gymrl_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_peer_reducer(n_floats: int):
    """PeerReducer when the run is multi-GPU on NCCL and GYMRL_COMM != 'nccl'; None otherwise (single GPU, gloo tests) or when
    the peer mappings cannot be set up on EVERY rank (the caller then keeps the ncclAllReduce path — still a device path).
    Each local stage is followed by an all-reduced status so that the ranks always agree on which collectives come next."""
    import os
    rank, world = info()
    if world == 1 or os.environ.get("GYMRL_COMM", "p2p") == "nccl" or dist.get_backend() != "nccl":
        return None
    dev = torch.device("cuda", torch.cuda.current_device())

    def all_ok(flag: bool) -> bool:
        t = torch.tensor([1.0 if flag else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return t.item() >= 1.0

    red, handle, why = None, b"", ""
    try:
        red = PeerReducer(n_floats)
        handle = red.handle()
    except Exception as e:
        why = str(e)
    if not all_ok(red is not None and len(handle) > 0):
        if red is not None:
            red.close()
        if rank == 0 or why:
            print(f"[gymrl_b200.dist] rank {rank}: peer-memory reduction unavailable ({why or 'another rank failed'}); using ncclAllReduce", flush=True)
        return None
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    allh = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    try:
        red.open(b"".join(bytes(t.cpu().tolist()) for t in allh))
        opened = True
    except Exception as e:     # e.g. no peer access between the GPUs of this box
        opened, why = False, str(e)
    if not all_ok(opened):
        red.close()
        if rank == 0 or why:
            print(f"[gymrl_b200.dist] rank {rank}: peer mappings unavailable ({why or 'another rank failed'}); using ncclAllReduce", flush=True)
        return None
    return red


def grad_scale() -> float:
    """Gradients are summed across ranks; Adam rescales by 1/world so the step equals the single-GPU step on the
    concatenated minibatch (each rank's loss is a mean over its local minibatch)."""
    return 1.0 / info()[1]
