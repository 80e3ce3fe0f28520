"""CUDA-graph capture helper shared by the trainers (the launch-bound inner loops — a PPO rollout / epoch, an off-policy
lockstep — are captured once and replayed; every per-replay-varying scalar lives in device memory)."""
from __future__ import annotations

import torch

from . import _ffi


def capture(fn):
    """Run `fn` once eagerly on a side stream (first-launch module loads; it is a REAL execution of fn), then record it.
    Returns the graph; `graph.n_kernels` = launches of this library recorded in it."""
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    c0 = _ffi.launch_count()
    with torch.cuda.graph(g):
        fn()
    g.n_kernels = _ffi.launch_count() - c0
    return g
