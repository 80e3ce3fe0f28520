"""Host-side parameter plumbing: flat fp32 parameter/gradient buffers behind ordinary nn.Modules, and
the optimizer object the reference scripts expose (``trainer.optimizer.param_groups[0]['lr']``).

The networks stay ``torch.nn.Module``s with the reference's sub-module names so ``state_dict()`` /
``load_state_dict()`` and utils.model.ModelLoader keep working (SURVEY §5 checkpoint row, q18), but
their storage is ONE contiguous buffer so that gradient all-reduce, global-norm clipping, Adam and
Polyak each run as a single kernel over it.  No torch autograd / torch.optim math is used on the hot
path.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops


def flatten_module(module: nn.Module, order: Optional[Sequence[str]] = None, device=None):
    """Move all parameters of `module` into one flat fp32 buffer (in `order`, default registration order).

    Returns (flat_param, flat_grad, views) where views[name] = (offset, shape).  Each nn.Parameter's
    ``.data`` / ``.grad`` become views into the flat buffers.
    """
    named = dict(module.named_parameters())
    names = list(order) if order is not None else list(named.keys())
    assert sorted(names) == sorted(named.keys()), "order must list every parameter exactly once"
    device = device or next(iter(named.values())).device
    # matrices start 16-byte aligned so GEMM operand loads can be float4; vectors (biases) are packed back to
    # back so that sibling heads' biases stay contiguous (FlatParams.span views them as one vector)
    offsets, total = {}, 0
    for n in names:
        if named[n].dim() >= 2:
            total = (total + 3) // 4 * 4
        offsets[n] = total
        total += named[n].numel()
    total = (total + 3) // 4 * 4
    flat = torch.zeros(total, device=device, dtype=torch.float32)
    grad = torch.zeros(total, device=device, dtype=torch.float32)
    views: Dict[str, tuple] = {}
    for n in names:
        p = named[n]
        o, k = offsets[n], p.numel()
        flat[o:o + k].copy_(p.detach().reshape(-1).to(device, torch.float32))
        p.data = flat[o:o + k].view(p.shape)
        p.grad = grad[o:o + k].view(p.shape)
        views[n] = (o, tuple(p.shape))
    return flat, grad, views


class FlatParams:
    """A module's parameters as (flat, grad) + named 2-D/1-D views for the kernels."""

    def __init__(self, module: nn.Module, order: Optional[Sequence[str]] = None, device=None):
        self.module = module
        self.flat, self.grad, self.views = flatten_module(module, order, device)

    # -- pre-split tf32 weight images (TMA operand of the warp-specialised GEMM, csrc/wimages.cu) ---------------------------------
    def enable_weight_images(self, matrices):
        """matrices: [(first_name, last_name, rows, cols)] spans of consecutive parameters that are used as one [rows, cols] GEMM
        operand (a single parameter: first == last).  The library keeps the images current on every optimiser step / Polyak
        update of this buffer; refresh_views() and refresh_weight_images() cover host-side writes."""
        mats = []
        for first, last, rows, cols in matrices:
            o0 = self.views[first][0]
            o1, shape1 = self.views[last]
            n1 = 1
            for s_ in shape1:
                n1 *= s_
            assert o1 + n1 - o0 == rows * cols and o0 % 4 == 0, f"{first}..{last} is not a contiguous, 16-byte aligned [{rows}, {cols}] matrix"
            mats.append((o0, rows, cols))
        self._wimages = ops.weight_images_register(self.flat, mats)

    def refresh_weight_images(self):
        if getattr(self, "_wimages", None) is not None:
            ops.weight_images_refresh(self.flat)

    def __del__(self):
        try:
            if getattr(self, "_wimages", None) is not None:
                ops.weight_images_unregister(self.flat)
        except Exception:
            pass

    def p(self, name: str) -> torch.Tensor:
        o, shape = self.views[name]
        k = 1
        for s in shape:
            k *= s
        return self.flat[o:o + k].view(shape)

    def g(self, name: str) -> torch.Tensor:
        o, shape = self.views[name]
        k = 1
        for s in shape:
            k *= s
        return self.grad[o:o + k].view(shape)

    def span(self, first: str, last: str, rows: int, cols: int, grad: bool = False) -> torch.Tensor:
        """View consecutive parameters [first .. last] as one [rows, cols] matrix (e.g. actor.0.weight and
        critic.0.weight stacked into a single [512, 256] GEMM operand)."""
        o0 = self.views[first][0]
        o1, shape1 = self.views[last]
        n1 = 1
        for s in shape1:
            n1 *= s
        assert o1 + n1 - o0 == rows * cols, f"parameters {first}..{last} are not contiguous in the flat buffer"
        buf = self.grad if grad else self.flat
        return buf[o0:o0 + rows * cols].view(rows, cols)

    def numel(self) -> int:
        return self.flat.numel()

    def n_params(self) -> int:
        """Parameter elements without the alignment padding of the flat buffer (whose gradient stays zero)."""
        n = 0
        for (_, shape) in self.views.values():
            k = 1
            for s in shape:
                k *= s
            n += k
        return n

    def refresh_views(self):
        """Re-point nn.Parameter.data at the flat buffer (after load_state_dict replaced storages)."""
        named = dict(self.module.named_parameters())
        for n, (o, shape) in self.views.items():
            p = named[n]
            k = p.numel()
            if p.data.data_ptr() != self.flat[o:o + k].data_ptr():
                self.flat[o:o + k].copy_(p.data.reshape(-1))
                p.data = self.flat[o:o + k].view(shape)
            p.grad = self.grad[o:o + k].view(shape)
        self.refresh_weight_images()


class FusedAdam:
    """torch.optim.Adam's surface (param_groups / state_dict / zero_grad / step) over one flat buffer,
    executed by gymrl_adam_step (+ gymrl_grad_sumsq for clip_grad_norm_)."""

    def __init__(self, params: FlatParams, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        self.fp = params
        dev = params.flat.device
        self.exp_avg = torch.zeros_like(params.flat)
        self.exp_avg_sq = torch.zeros_like(params.flat)
        self.step_t = torch.zeros(1, device=dev, dtype=torch.int32)
        self.lr_t = torch.full((1,), float(lr), device=dev, dtype=torch.float64)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.sumsq_partials = torch.zeros(4096, device=dev, dtype=torch.float64)   # reduce_flush's per-block sums of squares
        self.done_ctr = torch.zeros(1, device=dev, dtype=torch.int32)
        self._lr_host = float(lr)
        self.betas, self.eps = betas, eps
        self.param_groups: List[dict] = [{"lr": float(lr), "betas": betas, "eps": eps, "params": list(params.module.parameters())}]

    # -- host side ----------------------------------------------------------------------------------
    def sync_lr(self):
        """Push param_groups[0]['lr'] to the device scalar (outside any captured graph)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self.lr_t.fill_(lr)
            self._lr_host = lr

    def zero_grad(self, set_to_none: bool = False):
        self.fp.grad.zero_()

    # -- device side (capture-safe) -------------------------------------------------------------------
    def launch(self, max_norm: float = 0.0, clamp: float = 0.0, grad_scale: float = 1.0, grad=None):
        """Enqueue (optional global-norm) + Adam on the current stream; no host sync, no allocation.  `grad` = an alternative
        gradient buffer of the flat layout (the peer-reduced gradient of a multi-GPU step)."""
        grad = self.fp.grad if grad is None else grad
        if max_norm > 0.0:
            ops.grad_sumsq(grad, out=self.sumsq)
        ops.adam_step(self.fp.flat, grad, self.exp_avg, self.exp_avg_sq, self.lr_t, self.step_t,
                      beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
                      sumsq=self.sumsq if max_norm > 0.0 else None, max_norm=max_norm, clamp=clamp, grad_scale=grad_scale)

    def launch_clipped(self, n_partials: int, max_norm: float, grad_scale: float = 1.0, grad=None):
        """Clip + Adam in one launch, the norm taken from self.sumsq_partials[:n_partials] (filled by
        ops.reduce_flush(self.sumsq_partials) over *all* of this optimiser's gradients, or by PeerReducer.allreduce_sumsq)."""
        ops.clip_adam_step(self.fp.flat, self.fp.grad if grad is None else grad, self.exp_avg, self.exp_avg_sq, self.lr_t, self.step_t,
                           sumsq_partials=self.sumsq_partials, n_partials=n_partials, done_counter=self.done_ctr, max_norm=max_norm,
                           beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, grad_scale=grad_scale)

    def step(self, max_norm: float = 0.0, clamp: float = 0.0, grad_scale: float = 1.0):
        self.sync_lr()
        self.launch(max_norm, clamp, grad_scale)

    # -- checkpoint surface: torch.optim.Adam's own layout, so utils.model.ModelLoader round-trips with reference checkpoints ----
    def _param_slices(self):
        """(offset, numel, shape) of every parameter in module.parameters() order — the order torch.optim.Adam numbers them."""
        by_id = {id(p): n for n, p in self.fp.module.named_parameters()}
        out = []
        for p in self.fp.module.parameters():
            o, shape = self.fp.views[by_id[id(p)]]
            out.append((o, p.numel(), shape))
        return out

    def state_dict(self):
        """{state: {i: {step, exp_avg, exp_avg_sq}}, param_groups: [...]} exactly as torch.optim.Adam.state_dict() (ref checkpoints
        hold `optimizer_state_dict` in this format, utils/model.py:337-366).  `lr` is the current (possibly annealed) value, as in torch."""
        step = float(self.step_t.item())
        state = {}
        for i, (o, k, shape) in enumerate(self._param_slices()):
            state[i] = {"step": torch.tensor(step), "exp_avg": self.exp_avg[o:o + k].view(shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + k].view(shape).clone()}
        if step == 0:
            state = {}   # torch creates per-parameter state lazily at the first step
        g = {"lr": float(self.param_groups[0]["lr"]), "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "decoupled_weight_decay": False, "params": list(range(len(self._param_slices())))}
        return {"state": state, "param_groups": [g]}

    def load_state_dict(self, sd):
        """Accepts torch.optim.Adam's layout (reference checkpoints) and this class's round-1 private layout."""
        if "param_groups" in sd:
            slices = self._param_slices()
            g = sd["param_groups"][0]
            assert len(g["params"]) == len(slices), "optimizer state does not match this network's parameter list"
            self.exp_avg.zero_(); self.exp_avg_sq.zero_()
            step = 0
            for j, pid in enumerate(g["params"]):
                st = sd["state"].get(pid)
                if st is None:
                    continue
                o, k, shape = slices[j]
                self.exp_avg[o:o + k].copy_(st["exp_avg"].reshape(-1).to(self.exp_avg.device, torch.float32))
                self.exp_avg_sq[o:o + k].copy_(st["exp_avg_sq"].reshape(-1).to(self.exp_avg.device, torch.float32))
                step = max(step, int(float(st["step"])))
            self.step_t.fill_(step)
            self.param_groups[0]["lr"] = float(g["lr"])
            # betas / eps stay the constructor's: they are baked into captured graphs and equal the script's own values
        else:
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            self.step_t.fill_(int(sd["step"]))
            self.param_groups[0]["lr"] = float(sd["lr"])
        self.sync_lr()


def layer_init(layer: nn.Module, std: float = 2 ** 0.5) -> nn.Module:
    """Orthogonal init, zero bias — same initialiser as the reference (algorithms/ppo_lunarlander.py:55-60).
    Runs once on the host at construction; parity tests load the reference's own state_dict."""
    if isinstance(layer, nn.Linear):
        nn.init.orthogonal_(layer.weight, gain=std)
        if layer.bias is not None:
            nn.init.constant_(layer.bias, 0)
    return layer
