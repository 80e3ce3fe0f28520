"""Chains of dense layers over flat parameter storage, executed by the C ABI kernels.

`Chain` is the building block of every reference network on the hot path that is a plain stack of
nn.Linear + activation (QNetwork, SAC/TD3 Actor trunk, the two halves of the twin Critic, the trunk of
the dueling net).  It owns its activation / gradient scratch for a fixed maximum batch, so forward and
backward allocate nothing and can be captured in CUDA graphs.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _ffi, ops
from .nn import FlatParams

f32 = torch.float32


class Chain:
    BRANCH_MAX_ROWS = 8192     # backward(): dW launches on a side branch up to this batch size

    def __init__(self, fp: FlatParams, layers: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, int]],
                 max_batch: int, backward: bool = True):
        """layers: (W [N,K], b [N], gW, gb, act) per layer, W/b/gW/gb being views into fp.flat / fp.grad."""
        self.fp, self.layers, self.M = fp, list(layers), int(max_batch)
        dev = fp.flat.device
        self.h: List[torch.Tensor] = [torch.empty(self.M, W.shape[0], device=dev, dtype=f32) for (W, _, _, _, _) in self.layers]
        self.d: List[Optional[torch.Tensor]] = [None] * len(self.layers)
        self.ws = None
        self.ws_l = [None] * len(self.layers)
        self._br = None
        if backward:
            # d[l] = gradient wrt layer l's (post-activation-derivative) pre-activation; d[-1] is supplied by the loss
            self.d = [torch.zeros(self.M, W.shape[0], device=dev, dtype=f32) for (W, _, _, _, _) in self.layers]
            need = max(ops.backward_weight_workspace(self.M, W.shape[0], W.shape[1]) for (W, _, _, _, _) in self.layers)
            self.ws = torch.empty(need, device=dev, dtype=torch.uint8)
            # one workspace per layer: the layers' dW launches may run on a side branch next to the next layer's
            self.ws_l = [torch.empty(ops.backward_weight_workspace(self.M, W.shape[0], W.shape[1]), device=dev, dtype=torch.uint8)
                         for (W, _, _, _, _) in self.layers]
            self.dx = torch.zeros(self.M, self.layers[0][0].shape[1], device=dev, dtype=f32)
            if len(self.layers) > 1 and self.M <= self.BRANCH_MAX_ROWS:
                from .graphs import Branches
                self._br = Branches(len(self.layers))      # created here, never inside a graph capture

    @staticmethod
    def from_names(fp: FlatParams, specs: Sequence[Tuple[str, str, int]], max_batch: int, backward: bool = True) -> "Chain":
        return Chain(fp, [(fp.p(w), fp.p(b), fp.g(w), fp.g(b), act) for (w, b, act) in specs], max_batch, backward)

    @property
    def out(self) -> torch.Tensor:
        return self.h[-1]

    @property
    def dout(self) -> torch.Tensor:
        return self.d[-1]

    def forward(self, x: torch.Tensor, M: Optional[int] = None, row_index: Optional[torch.Tensor] = None,
                weights: Optional[Sequence[Tuple[torch.Tensor, torch.Tensor]]] = None) -> torch.Tensor:
        """weights: optional replacement (W, b) per layer (same shapes) — used to run a target network's
        parameters through this chain's scratch."""
        M = self.M if M is None else M
        inp, ri = x, row_index
        for l, (W, b, _, _, act) in enumerate(self.layers):
            if weights is not None:
                W, b = weights[l]
            ops.linear_forward(inp, W, b, act, row_index=ri, out=self.h[l], M=M)
            inp, ri = self.h[l], None
        return self.h[-1]

    def backward(self, x: torch.Tensor, M: Optional[int] = None, row_index: Optional[torch.Tensor] = None, *, param_grads: bool = True,
                 input_grad: bool = False, accumulate: bool = False):
        """Given self.d[-1] (= dL/d output), fill parameter gradients (and optionally self.dx = dL/dx)."""
        M = self.M if M is None else M
        L = len(self.layers)

        def dW(l):
            W, b, gW, gb, act = self.layers[l]
            inp = self.h[l - 1] if l > 0 else x
            ops.linear_backward_weight(self.d[l], inp, gW, gb, row_index=row_index if l == 0 else None, workspace=self.ws_l[l],
                                       accumulate=accumulate, M=M)

        def dX(l):
            W = self.layers[l][0]
            if l > 0:
                ops.linear_backward_input(self.d[l][:M], W, self.h[l - 1], self.layers[l - 1][4], out=self.d[l - 1])
            elif input_grad:
                ops.linear_backward_input(self.d[0][:M], W, None, _ffi.ACT_NONE, out=self.dx)

        if param_grads and self._br is not None and M <= self.BRANCH_MAX_ROWS:
            # small batches (off-policy updates): a layer's weight gradient does not feed the chain of input gradients, so every
            # layer's dW launches (GEMM / sweep + fold) run as a side branch of their own next to the dX chain.  At PPO's minibatch sizes the
            # GEMMs fill the chip and nothing overlaps (measured slower there), hence the row limit.
            main = torch.cuda.current_stream()
            for l in range(L - 1, -1, -1):
                side = self._br.side[l]
                side.wait_stream(main)              # d[l] is final on the main branch (the loss, or dX(l + 1))
                with torch.cuda.stream(side):
                    dW(l)
                dX(l)
            for l in range(L):
                main.wait_stream(self._br.side[l])
        else:
            for l in range(L - 1, -1, -1):
                if param_grads:
                    dW(l)
                dX(l)
        return self.dx if input_grad else None
