"""CUDA-graph capture helper shared by the trainers (the launch-bound inner loops — a PPO rollout / epoch, an off-policy
lockstep — are captured once and replayed; every per-replay-varying scalar lives in device memory)."""
from __future__ import annotations

import torch

from . import _ffi


def capture(fn, warmup: bool = True):
    """Run `fn` once eagerly on a side stream (first-launch module loads; it is a REAL execution of fn) unless the caller
    already did (`warmup=False`), then record it.  Returns the graph; `graph.n_kernels` = launches of this library in it."""
    torch.cuda.synchronize()
    if warmup:
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


class Branches:
    """Independent launch sequences as parallel branches: fns[0] on the current stream, the others on side streams, joined
    before returning.  Inside a CUDA-graph capture the side streams join the capture (event fork / join), so the recorded graph
    has parallel branches; eagerly they are ordinary concurrent streams.  The off-policy updates are made of kernels of <= 128
    small CTAs that leave most of the chip idle — the twin critics, or the online / target forwards of a double-Q update, are
    independent and overlap (every branch must use only preallocated tensors of its own)."""

    def __init__(self, n_side: int = 2):
        self.side = [torch.cuda.Stream() for _ in range(n_side)]

    def run(self, *fns):
        if len(fns) - 1 > len(self.side):
            raise ValueError("more branches than side streams")
        main = torch.cuda.current_stream()
        used = self.side[:len(fns) - 1]
        for s in used:
            s.wait_stream(main)
        outs = [None] * len(fns)
        for k, fn in enumerate(fns[1:]):
            with torch.cuda.stream(used[k]):
                outs[k + 1] = fn()
        outs[0] = fns[0]()
        for s in used:
            main.wait_stream(s)
        return outs


class LockstepGraphs:
    """Mixin for the continuous-control off-policy trainers (TD3, DDPG): one captured CUDA graph per *phase* of the
    lockstep (TD3 updates the actor every `policy_freq`-th update, so its lockstep has `policy_freq` phases).

    The trainer provides: `_lockstep_body()` (capture-safe once the replay holds a batch: every per-step-varying scalar in
    device memory), `memory` (ReplayRing with `_size_host`), `B`, `N`, `total_updates`, `act_count`, `cfg.use_cuda_graph`
    and optionally `_lockstep_phases()`.
    """

    def _lockstep_phases(self) -> int:
        return 1

    def lockstep(self):
        mem = self.memory
        if not getattr(self.cfg, "use_cuda_graph", True) or len(mem) < self.B:
            return self._lockstep_body()
        graphs = self.__dict__.setdefault("_g_lockstep", {})
        phase = (self.total_updates + 1) % self._lockstep_phases()      # phase of the update this lockstep will run
        mirrors = (self.act_count, self.total_updates, mem._size_host)
        if phase not in graphs:
            self._lockstep_body()                                       # this lockstep, eagerly (also the warm-up) ...
            self.act_count, self.total_updates, mem._size_host = mirrors   # ... then record the same phase without running it
            graphs[phase] = capture(self._lockstep_body, warmup=False)
        else:
            graphs[phase].replay()
            self.graph_launches = getattr(self, "graph_launches", 0) + graphs[phase].n_kernels
        # host mirrors of the counters the device body advanced
        self.act_count, self.total_updates = mirrors[0] + 1, mirrors[1] + 1
        mem._size_host = min(mem.capacity, mirrors[2] + self.N)


class HostScheduledLockstep:
    """Mixin for the value-based off-policy trainers (DQN, NoisyNet DQN, DDQN + PER): `lockstep()` = act -> env step -> store ->
    update as ONE captured CUDA graph, with the host-side schedules around it:

        _lockstep_ready()   -> True once the body is capture-safe (replay holds a batch)
        _before_lockstep()  -> write host-scheduled scalars (epsilon, PER beta, learning rate) into their device slots
        _lockstep_body()    -> the device work (every per-step-varying value read from device memory)
        _after_lockstep()   -> host mirrors of the device counters + host-scheduled follow-ups (hard target sync)
        _host_mirrors() / _set_host_mirrors(m)  -> counters the body advances on the host while it is being recorded
    """

    def lockstep(self):
        if not getattr(self.cfg, "use_cuda_graph", True) or not self._lockstep_ready():
            self._before_lockstep()
            self._lockstep_body()
            self._after_lockstep()
            return
        self._before_lockstep()
        g = self.__dict__.get("_g_lockstep")
        if g is None:
            self._lockstep_body()                       # this lockstep, eagerly (also the warm-up) ...
            m = self._host_mirrors()
            self._g_lockstep = capture(self._lockstep_body, warmup=False)    # ... then record one without running it
            self._set_host_mirrors(m)
        else:
            g.replay()
            self.graph_launches = getattr(self, "graph_launches", 0) + g.n_kernels
            self._advance_host_mirrors()
        self._after_lockstep()
